// fpop_oracle.cpp -- CPU restatement of the reference PeakSegFPOP solver.
//
// TEST INFRASTRUCTURE ONLY.  Nothing in the product (peaksegdisk_b200/, include/) may include,
// link or call this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs use it, as the checker.  Parity is PINNED: tests/test_oracle.py checks
// this restatement byte-for-byte (segments.bed, loss.tsv and the per-row cost-function db) against
// the unmodified reference compiled by oracle/Makefile into oracle/_ref/, on the reference's own
// test vectors and on seeded synthetic inputs.
//
// What it restates (all file:line into /root/reference):
//   * one Poisson-loss piece g(x) = a*e^x + b*x + c on [lo,hi], x = log(mean)
//       cost/slope/argmin            src/funPieceListLog.cpp:52-65, 192-234
//       two_roots                    src/funPieceListLog.cpp:29-50
//       root_left / root_right       src/funPieceListLog.cpp:129-190 / 69-127   (Newton, <=100 steps)
//   * piecewise operators on a sorted vector of pieces (the reference uses std::list)
//       min_less                     src/funPieceListLog.cpp:236-437
//       min_more                     src/funPieceListLog.cpp:439-616
//       min_env + pair rule + append src/funPieceListLog.cpp:832-860, 870-1259, 1261-1285
//       shift/scale/stamp            src/funPieceListLog.cpp:618-641
//       best_piece / locate          src/funPieceListLog.cpp:689-712 / 643-653
//   * the solver driver              src/PeakSegFPOPLog.cpp:143-463
//   * the scratch db byte layout     src/PeakSegFPOPLog.cpp:12-34, 76-141
//
// Arithmetic: every floating-point expression keeps the reference's operand order and rounding
// (compile with -ffp-contract=off, no -mfma).  exp/log are the system libm by default, exactly as
// in the reference; oracle_set_math(1) switches to the product's psd_math.h implementations, which
// are bit-identical to glibc 2.39's FMA variants (useful on a host whose libm differs).
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <string>
#include <vector>
#include <stdexcept>
#include "psd_math.h"

namespace {

const double kEps = 1e-12;      // NEWTON_EPSILON, src/funPieceListLog.cpp:9
const int kMaxSteps = 100;      // NEWTON_STEPS,   src/funPieceListLog.cpp:10
const int kBackUnset = -3;      // PREV_NOT_SET,   src/funPieceListLog.cpp:11
const double kInf = INFINITY;

int g_math_mode = 0;
inline double xexp(double x) { return g_math_mode ? psd_exp(x, psd_exp_tab_host) : std::exp(x); }
inline double xlog(double x) { return g_math_mode ? psd_log(x, psd_log_tab_host) : std::log(x); }
inline double mag(double v) { return v < 0 ? -v : v; }  // the reference's ABS macro (:13)

struct Piece {
  double a, b, c;   // coefficients of e^x, x, 1
  double lo, hi;    // interval of log-mean
  int back_i;       // last row of the previous segment
  double back_x;    // log-mean of the previous segment, or +inf when the constraint is tight
};
typedef std::vector<Piece> Fun;

Piece mk(double a, double b, double c, double lo, double hi, int bi, double bx) {
  Piece p; p.a = a; p.b = b; p.c = c; p.lo = lo; p.hi = hi; p.back_i = bi; p.back_x = bx; return p;
}

// ---- one piece -------------------------------------------------------------------------------
double cost(const Piece& p, double x) {                       // :206-222
  double et = (x == -kInf) ? 0.0 : p.a * xexp(x);
  double lt = (p.b == 0) ? 0.0 : p.b * x;
  return et + lt + p.c;
}
double slope(const Piece& p, double x) {                      // :224-234
  double et = (x == -kInf) ? 0.0 : p.a * xexp(x);
  return et + p.b;
}
double cost_m(const Piece& p, double m) {                     // :52-61 (mean space)
  double base = p.a * m + p.c;
  if (p.b == 0) return base;
  double lm = xlog(m);
  double prod = lm * p.b;
  return base + prod;
}
double slope_m(const Piece& p, double m) { return p.a + p.b / m; }   // :63-65
double opt_m(const Piece& p) { return -p.b / p.a; }                   // :192-197
double opt_x(const Piece& p) { return xlog(opt_m(p)); }               // :199-204

bool two_roots(const Piece& p, double level) {                // :29-50
  if (p.b == 0) throw std::runtime_error("two_roots on degenerate piece");
  double m = opt_m(p);
  double x = xlog(m);
  double c1 = cost(p, x);
  double c2 = cost_m(p, m);
  if (0 < p.a) return c1 + kEps < level && c2 + kEps < level;
  return level + kEps < c1 && level + kEps < c2;
}

double root_right(const Piece& p, double level) {             // :69-127, Newton in mean space
  double m0 = opt_m(p);
  double c0 = cost_m(p, m0);
  double cr = cost(p, p.hi);
  if ((c0 < cr && cr < level) || (c0 > cr && cr > level)) return p.hi + 1;
  double m = m0 + 1;
  double f, pos_f = kInf, pos_m = kInf, neg_f = -kInf, neg_m = kInf;
  if (c0 < 0) { neg_f = c0; neg_m = m0; } else { pos_f = c0; pos_m = m0; }
  int step = 0;
  do {
    f = cost_m(p, m) - level;
    if (0 < f && f < pos_f) { pos_f = f; pos_m = m; }
    if (neg_f < f && f < 0) { neg_f = f; neg_m = m; }
    if (kMaxSteps <= ++step) {
      double mid = (pos_m + neg_m) / 2;
      double fm = cost_m(p, mid) - level;
      return (mag(fm) < mag(f)) ? xlog(mid) : xlog(m);
    }
    double d = slope_m(p, m);
    m = m - f / d;
  } while (kEps < mag(f));
  return xlog(m);
}

double root_left(const Piece& p, double level) {              // :129-190, Newton in log space
  double x0 = opt_x(p);
  double c0 = cost(p, x0);
  double cl = cost(p, p.lo);
  if ((level < cl && cl < c0) || (level > cl && cl > c0)) return p.lo - 1;
  double x = x0 - 1;
  double f, pos_f = kInf, pos_x = kInf, neg_f = -kInf, neg_x = kInf;
  if (c0 < 0) { neg_f = c0; neg_x = x0; } else { pos_f = c0; pos_x = x0; }
  int step = 0;
  do {
    f = cost(p, x) - level;
    if (0 < f && f < pos_f) { pos_f = f; pos_x = x; }
    if (neg_f < f && f < 0) { neg_f = f; neg_x = x; }
    if (kMaxSteps <= ++step) {
      double mid = (pos_x + neg_x) / 2;
      double fm = cost(p, mid) - level;
      return (mag(fm) < mag(f)) ? mid : x;
    }
    double d = slope(p, x);
    double off = f / d;
    x = x - off;
  } while (kEps < mag(f));
  return x;
}

bool same_coefs(const Piece& p, const Piece& q) {             // sameFuns :862-868
  return p.a == q.a && p.b == q.b && mag(p.c - q.c) < kEps;
}

// ---- running minimum from the left ----------------------------------
void min_less(const Fun& in, Fun& out) {
  out.clear();
  const int n = (int)in.size();
  double level = kInf;           // cost of the pending flat piece, +inf = tracking the input
  double left_edge = in[0].lo;   // where the next output piece starts
  double arg_at = kInf;          // where the flat piece's minimum was attained
  int i = 0;
  while (i < n) {
    const Piece& p = in[i];
    double cl = cost(p, p.lo), cr = cost(p, p.hi);
    if (level == kInf) {
      bool has_next = i + 1 < n;
      if (p.b == 0) {
        bool flat = (cr - cl) < kEps;
        bool next_above = true;
        if (has_next) { double nl = cost(in[i + 1], in[i + 1].lo); next_above = kEps < nl - cl; }
        if (next_above && !flat) { level = cl; arg_at = p.lo; }
        else { out.push_back(mk(p.a, p.b, p.c, left_edge, p.hi, kBackUnset, kInf)); left_edge = p.hi; }
      } else {
        double mu = opt_x(p);
        double cmu = cost(p, mu);
        bool next_ok = true;
        if (has_next) { double nl = cost(in[i + 1], in[i + 1].lo); next_ok = kEps < nl - cmu; }
        bool ok = kEps < cr - cmu && next_ok;
        if (mu <= p.lo && ok) { level = cost(p, p.lo); arg_at = p.lo; }
        else if (mu < p.hi && ok) {
          if (left_edge < mu) out.push_back(mk(p.a, p.b, p.c, left_edge, mu, kBackUnset, kInf));
          left_edge = mu; arg_at = mu; level = cmu;
        } else { out.push_back(mk(p.a, p.b, p.c, left_edge, p.hi, kBackUnset, kInf)); left_edge = p.hi; }
      }
    } else {
      if (p.b == 0) {
        if (p.a < 0) throw std::runtime_error("decreasing degenerate piece in min_less");
      } else {
        if (two_roots(p, level)) {
          double r = root_left(p, level);
          (void)cost(p, r);  // the reference evaluates it for its verbose trace only
          if (p.lo < r && r < p.hi) {
            out.push_back(mk(0, 0, level, left_edge, r, kBackUnset, arg_at));
            level = kInf; left_edge = r;
            i--;  // revisit this piece in tracking mode
          }
        }
        if (cr <= level + kEps && level < kInf) {
          out.push_back(mk(0, 0, level, left_edge, p.hi, kBackUnset, arg_at));
          level = kInf; left_edge = p.hi;
        }
      }
    }
    i++;
  }
  if (level < kInf) out.push_back(mk(0, 0, level, left_edge, in[n - 1].hi, kBackUnset, arg_at));
}

// ---- running minimum from the right (Appendix E.2); built back to front ----------------------
void min_more(const Fun& in, Fun& out) {
  std::vector<Piece> rev;  // pieces in emission order (right to left)
  const int n = (int)in.size();
  double level = kInf;
  double right_edge = in[n - 1].hi;
  double arg_at = kInf;
  int i = n;  // the loop decrements before use
  while (i != 0) {
    i--;
    const Piece& p = in[i];
    if (level == kInf) {
      if (p.b == 0) {
        rev.push_back(mk(p.a, p.b, p.c, p.lo, right_edge, kBackUnset, kInf)); right_edge = p.lo;
      } else {
        double mu = opt_x(p);
        double cmu = cost(p, mu);
        bool prev_ok = true;
        if (i != 0) { double pr = cost(in[i - 1], in[i - 1].hi); prev_ok = kEps < pr - cmu; }
        double cl = cost(p, p.lo);
        if (p.hi <= mu) {
          double cr = cost(p, p.hi);
          double drop = cl - cr;
          if (kEps < drop) { level = cr; arg_at = p.hi; }
          else { rev.push_back(mk(p.a, p.b, p.c, p.lo, right_edge, kBackUnset, kInf)); right_edge = p.lo; }
        } else if (p.lo < mu && kEps < cl - cmu && prev_ok) {
          if (mu < right_edge) rev.push_back(mk(p.a, p.b, p.c, mu, right_edge, kBackUnset, kInf));
          right_edge = mu; arg_at = mu; level = cmu;
        } else { rev.push_back(mk(p.a, p.b, p.c, p.lo, right_edge, kBackUnset, kInf)); right_edge = p.lo; }
      }
    } else {
      double cl = cost(p, p.lo);
      (void)cost(p, p.hi);
      double r = kInf;
      if (p.b == 0) r = xlog((level - p.c) / p.a);
      else if (two_roots(p, level)) r = root_right(p, level);
      if (p.lo < r && r < p.hi) {
        rev.push_back(mk(0, 0, level, r, right_edge, kBackUnset, arg_at));
        level = kInf; right_edge = r;
        i++;  // revisit this piece in tracking mode
      } else if (cl <= level + kEps) {
        rev.push_back(mk(0, 0, level, p.lo, right_edge, kBackUnset, arg_at));
        level = kInf; right_edge = p.lo;
      }
    }
  }
  if (level < kInf) rev.push_back(mk(0, 0, level, in[0].lo, right_edge, kBackUnset, arg_at));
  out.assign(rev.rbegin(), rev.rend());
}

// ---- pointwise minimum of two functions (Appendix E.3) ---------------------------------------
void append(Fun& out, const Piece& src, double lo, double hi) {       // push_piece :1261-1285
  if (hi <= lo) return;
  if (!out.empty()) {
    Piece& last = out.back();
    if (same_coefs(last, src) && src.back_x == last.back_x && src.back_i == last.back_i) { last.hi = hi; return; }
  }
  out.push_back(mk(src.a, src.b, src.c, lo, hi, src.back_i, src.back_x));
}

void min_pair(const Fun& f, const Fun& g, int i, int j, Fun& out) {   // push_min_pieces :870-1259
  const Piece& p = f[i];
  const Piece& q = g[j];
  const int nf = (int)f.size(), ng = (int)g.size();
  bool eq_left, eq_right;
  double lo, hi;
  if (p.lo < q.lo) { eq_left = same_coefs(g[j - 1], p); lo = q.lo; }
  else {
    lo = p.lo;
    if (q.lo < p.lo) eq_left = same_coefs(f[i - 1], q);
    else eq_left = (i == 0 && j == 0) ? false : same_coefs(f[i - 1], g[j - 1]);
  }
  if (p.hi < q.hi) { eq_right = same_coefs(f[i + 1], q); hi = p.hi; }
  else {
    hi = q.hi;
    if (q.hi < p.hi) eq_right = same_coefs(p, g[j + 1]);
    else eq_right = (i + 1 == nf && j + 1 == ng) ? false : same_coefs(f[i + 1], g[j + 1]);
  }
  if (lo == hi) return;
  if (same_coefs(p, q)) { append(out, p, lo, hi); return; }
  Piece d = mk(p.a - q.a, p.b - q.b, p.c - q.c, lo, hi, -5, 0.0);
  double mid_m = (xexp(hi) + xexp(lo)) / 2;
  double dmid = cost(d, xlog(mid_m));
  const Piece& by_mid = (dmid < 0) ? p : q;
  if (eq_left && eq_right) { append(out, by_mid, lo, hi); return; }
  if (d.b == 0) {
    if (d.a == 0) { append(out, d.c < 0 ? p : q, lo, hi); return; }
    if (d.c == 0) { append(out, d.a < 0 ? p : q, lo, hi); return; }
    double x = xlog(-d.c / d.a);
    if (lo < x && x < hi) {
      if (0 < d.a) { append(out, p, lo, x); append(out, q, x, hi); }
      else { append(out, q, lo, x); append(out, p, x, hi); }
      return;
    }
    append(out, by_mid, lo, hi);
    return;
  }
  double dl = cost(d, lo), dr = cost(d, hi);
  bool two = two_roots(d, 0.0);
  double rs = kInf, rl = kInf;
  if (two) { rs = root_left(d, 0.0); rl = root_right(d, 0.0); }
  if (eq_right) {
    if (two) {
      (void)cost(d, (rs + hi) / 2);
      double xo = opt_x(d);
      if (lo < rs && rs < xo && xo < hi) {
        if (dl < 0) { append(out, p, lo, rs); append(out, q, rs, hi); }
        else { append(out, q, lo, rs); append(out, p, rs, hi); }
        return;
      }
      bool p_low_at_zero = 0 < d.b;
      if (rs < lo) append(out, p_low_at_zero ? q : p, lo, hi);
      else append(out, p_low_at_zero ? p : q, lo, hi);
      return;
    }
    append(out, by_mid, lo, hi);
    return;
  }
  if (eq_left) {
    if (two) {
      double xo = opt_x(d);
      if (lo < xo && xo < rl && rl < hi) {
        if (dr < 0) { append(out, q, lo, rl); append(out, p, rl, hi); }
        else { append(out, p, lo, rl); append(out, q, rl, hi); }
        return;
      }
    }
    append(out, by_mid, lo, hi);
    return;
  }
  double x1 = kInf, x2 = kInf;
  if (two) {
    bool l_in = lo < rl && rl < hi;
    bool s_in = lo < rs && 0 < xexp(rs) && rs < hi;
    if (l_in) {
      if (s_in && rs < rl) { x1 = rs; x2 = rl; } else x1 = rl;
    } else if (s_in) x1 = rs;
  }
  if (x2 != kInf) {
    bool p_first;
    if (x2 - x1 < x1 - lo) {
      double bm = (xexp(lo) + xexp(x1)) / 2;
      p_first = cost(d, xlog(bm)) < 0;
    } else {
      p_first = !(cost(d, (x1 + x2) / 2) < 0);
    }
    if (p_first) { append(out, p, lo, x1); append(out, q, x1, x2); append(out, p, x2, hi); }
    else { append(out, q, lo, x1); append(out, p, x1, x2); append(out, q, x2, hi); }
  } else if (x1 != kInf) {
    double bm = (xexp(lo) + xexp(x1)) / 2;
    double before = cost(d, xlog(bm));
    double after = cost(d, (hi + x1) / 2);
    if (before < 0) {
      if (after < 0) append(out, p, lo, hi);
      else { append(out, p, lo, x1); append(out, q, x1, hi); }
    } else {
      if (after < 0) { append(out, q, lo, x1); append(out, p, x1, hi); }
      else append(out, q, lo, hi);
    }
  } else {
    double v = (mag(dmid) < kEps) ? dr : dmid;
    append(out, v < 0 ? p : q, lo, hi);
  }
}

void min_env(const Fun& f, const Fun& g, Fun& out) {                  // :832-860
  out.clear();
  size_t i = 0, j = 0;
  while (i < f.size() && j < g.size()) {
    min_pair(f, g, (int)i, (int)j, out);
    double reached = out.back().hi;
    bool adv_i = f[i].hi == reached, adv_j = g[j].hi == reached;
    if (adv_i) i++;
    if (adv_j) j++;
    if (!adv_i && !adv_j) throw std::runtime_error("min_env made no progress");
  }
}

void shift(Fun& f, double a, double b, double c) { for (auto& p : f) { p.a += a; p.b += b; p.c += c; } }  // :618-625
void scale(Fun& f, double s) { for (auto& p : f) { p.a *= s; p.b *= s; p.c *= s; } }                      // :627-634
void stamp(Fun& f, int row) { for (auto& p : f) p.back_i = row; }                                         // :636-641

void best_piece(const Fun& f, double* best_c, double* best_x, int* bi, double* bx) {  // Minimize :689-712
  *best_c = kInf;
  for (const auto& p : f) {
    double x = opt_x(p);
    if (x < p.lo) x = p.lo; else if (p.hi < x) x = p.hi;
    double c = cost(p, x);
    if (c < *best_c) { *best_c = c; *best_x = x; *bi = p.back_i; *bx = p.back_x; }
  }
}

// ---- scratch db, byte-compatible with the reference's DiskVector ------------------------------
// index of 2N 16-byte stream positions (offset + 8 zero bytes), then appended records
//   int32 size | int32 n_pieces | int32 chromEnd | n x { f64 hi, i32 back_i, f64 back_x }
struct Db {
  FILE* fp = nullptr;
  std::vector<int64_t> pos;   // file offset of each stored function (0 = unset)
  int64_t end = 0;
  bool open(const char* path, int n_entries) {
    fp = fopen(path, "w+b");
    if (!fp) return false;
    pos.assign(n_entries, 0);
    std::vector<char> zeros((size_t)16 * n_entries, 0);
    if (n_entries && fwrite(zeros.data(), 1, zeros.size(), fp) != zeros.size()) return false;
    end = (int64_t)16 * n_entries;
    return true;
  }
  bool put(int slot, const Fun& f, int chrom_end) {
    int n = (int)f.size();
    int size = 20 * n + 8;
    std::vector<char> buf(4 + size);
    char* w = buf.data();
    memcpy(w, &size, 4); w += 4;
    memcpy(w, &n, 4); w += 4;
    memcpy(w, &chrom_end, 4); w += 4;
    for (const auto& p : f) { memcpy(w, &p.hi, 8); w += 8; memcpy(w, &p.back_i, 4); w += 4; memcpy(w, &p.back_x, 8); w += 8; }
    if (fseeko(fp, end, SEEK_SET) != 0) return false;
    if (fwrite(buf.data(), 1, buf.size(), fp) != buf.size()) return false;
    char ent[16] = {0};
    memcpy(ent, &end, 8);
    if (fseeko(fp, (off_t)16 * slot, SEEK_SET) != 0) return false;
    if (fwrite(ent, 1, 16, fp) != 16) return false;
    pos[slot] = end;
    end += (int64_t)buf.size();
    return true;
  }
  void close() { if (fp) fclose(fp); fp = nullptr; }
};

// Stored function as the backtrack sees it (coefficients are not kept, lo is rebuilt).
struct Stored { std::vector<double> hi, back_x; std::vector<int> back_i; int chrom_end; };

void locate(const Stored& s, double x, int* bi, double* bx) {         // findMean :643-653
  double lo = -kInf;
  for (size_t k = 0; k < s.hi.size(); k++) {
    if (lo <= x && x <= s.hi[k]) { *bi = s.back_i[k]; *bx = s.back_x[k]; return; }
    lo = s.hi[k];
  }
}

struct Rows { std::vector<int> start, end, cov; std::string chrom; };

struct Solution {
  int status = 0;
  bool trivial = false;
  double penalty = 0;
  int n_rows = 0;
  double bases = 0, sum_wz = 0;
  double best_cost = 0;             // mean penalized cost
  double total_loss = 0;
  int n_segments = 0, n_peaks = 0, n_equality = 0;
  double total_intervals = 0, max_intervals = 0;
  std::vector<int> seg_start, seg_end;   // last segment first, as the reference writes them
  std::vector<int> seg_peak;             // 1 = peak, 0 = background
  std::vector<double> seg_mean;
};

typedef void (*row_hook_t)(void* user, int row, int which /*0=up,1=down*/, int n,
                           const double* a, const double* b, const double* c, const double* hi,
                           const int* back_i, const double* back_x);
row_hook_t g_hook = nullptr;
void* g_hook_user = nullptr;

void call_hook(int row, int which, const Fun& f) {
  if (!g_hook) return;
  std::vector<double> a, b, c, hi, bx; std::vector<int> bi;
  for (const auto& p : f) { a.push_back(p.a); b.push_back(p.b); c.push_back(p.c); hi.push_back(p.hi); bi.push_back(p.back_i); bx.push_back(p.back_x); }
  g_hook(g_hook_user, row, which, (int)f.size(), a.data(), b.data(), c.data(), hi.data(), bi.data(), bx.data());
}

// The DP + backtrack on in-memory rows (src/PeakSegFPOPLog.cpp:224-455 without the text I/O).
// db may be null (functions are then kept in memory only).
int solve_rows(const Rows& R, double penalty, bool penalty_is_inf, Db* db, Solution& S) {
  const int N = (int)R.cov.size();
  S.n_rows = N; S.penalty = penalty;
  double W = 0, SWZ = 0, xmin = kInf, xmax = -kInf;
  for (int t = 0; t < N; t++) {
    double w = R.end[t] - R.start[t];
    W += w; SWZ += w * R.cov[t];
    double lx = xlog((double)R.cov[t]);
    if (lx < xmin) xmin = lx;
    if (xmax < lx) xmax = lx;
  }
  S.bases = W; S.sum_wz = SWZ;
  if (penalty_is_inf || xmin == xmax) {                       // :224-243
    S.trivial = true;
    double bc = (SWZ != 0) ? SWZ * (1 - xlog(SWZ) + xlog(W)) : 0;
    S.best_cost = bc / W; S.total_loss = bc;
    S.n_segments = 1; S.n_peaks = 0;
    S.seg_start.push_back(R.start[0]); S.seg_end.push_back(R.end[N - 1]);
    S.seg_peak.push_back(0); S.seg_mean.push_back(SWZ / W);
    return 0;
  }
  std::vector<Stored> kept(db ? 0 : (size_t)2 * N);
  auto keep = [&](int slot, const Fun& f, int chrom_end) -> bool {
    if (db) return db->put(slot, f, chrom_end);
    Stored& s = kept[slot]; s.chrom_end = chrom_end;
    for (const auto& p : f) { s.hi.push_back(p.hi); s.back_i.push_back(p.back_i); s.back_x.push_back(p.back_x); }
    return true;
  };
  Fun up, down, up_prev, down_prev, tmp;
  double cw = 0, cw_prev = -1.0;
  for (int t = 0; t < N; t++) {                               // :258-397
    double w = R.end[t] - R.start[t];
    int z = R.cov[t];
    cw += w;
    if (t == 0) {
      down.push_back(mk(1.0, (double)-z, 0.0, xmin, xmax, -1, -5.0));
    } else {
      min_less(down_prev, tmp);
      stamp(tmp, t - 1);
      shift(tmp, 0.0, 0.0, penalty / cw_prev);
      if (t == 1) up = tmp; else min_env(tmp, up_prev, up);
      scale(up, cw_prev); shift(up, w, -z * w, 0.0); scale(up, 1 / cw);
      if (t == 1) down = down_prev;
      else { min_more(up_prev, tmp); stamp(tmp, t - 1); min_env(tmp, down_prev, down); }
      scale(down, cw_prev); shift(down, w, -z * w, 0.0); scale(down, 1 / cw);
    }
    cw_prev = cw;
    S.total_intervals += up.size() + down.size();
    if (S.max_intervals < up.size()) S.max_intervals = up.size();
    if (S.max_intervals < down.size()) S.max_intervals = down.size();
    up_prev = up; down_prev = down;
    call_hook(t, 0, up); call_hook(t, 1, down);
    if (!keep(t + N, down, R.end[t])) return 7;
    if (0 < t && !keep(t, up, R.end[t])) return 7;
  }
  double best_c, best_x = 0, back_x = 0; int back_i = 0;     // :400-442
  best_piece(down, &best_c, &best_x, &back_i, &back_x);
  int prev_end = R.end[N - 1];
  int offset = 0, n_eq = 0, n_seg = 1;
  auto fetch = [&](int slot) -> Stored {
    if (!db) return kept[slot];
    Stored s; int32_t hdr[3];
    fseeko(db->fp, db->pos[slot], SEEK_SET);
    if (fread(hdr, 4, 3, db->fp) != 3) throw std::runtime_error("db read");
    s.chrom_end = hdr[2];
    std::vector<char> buf((size_t)20 * hdr[1]);
    if (hdr[1] && fread(buf.data(), 1, buf.size(), db->fp) != buf.size()) throw std::runtime_error("db read");
    const char* r = buf.data();
    for (int k = 0; k < hdr[1]; k++) {
      double hi, bx; int bi;
      memcpy(&hi, r, 8); r += 8; memcpy(&bi, r, 4); r += 4; memcpy(&bx, r, 8); r += 8;
      s.hi.push_back(hi); s.back_i.push_back(bi); s.back_x.push_back(bx);
    }
    return s;
  };
  while (0 <= back_i) {
    n_seg++;
    Stored s = fetch(offset + back_i);
    S.seg_start.push_back(s.chrom_end); S.seg_end.push_back(prev_end);
    if (offset == 0) { offset = N; S.seg_peak.push_back(0); } else { offset = 0; S.seg_peak.push_back(1); }
    S.seg_mean.push_back(xexp(best_x));
    prev_end = s.chrom_end;
    if (back_x != kInf) best_x = back_x; else n_eq++;
    locate(s, best_x, &back_i, &back_x);
  }
  S.seg_start.push_back(R.start[0]); S.seg_end.push_back(prev_end);
  S.seg_peak.push_back(0); S.seg_mean.push_back(xexp(best_x));
  S.n_segments = n_seg; S.n_peaks = (n_seg - 1) / 2; S.n_equality = n_eq;
  S.best_cost = best_c; S.total_loss = best_c * cw - penalty * S.n_peaks;
  return 0;
}

int parse_penalty(const char* s, double* pen, bool* is_inf) {         // :145-159
  *is_inf = strcmp(s, "Inf") == 0;
  try { *pen = std::stod(s); }
  catch (const std::invalid_argument&) { return 10; }
  catch (const std::out_of_range&) { return 1; }   // the reference aborts here; documented deviation
  if (*is_inf) return 0;
  if (!std::isfinite(*pen)) return 1;
  if (*pen < 0) return 2;
  return 0;
}

int read_bedgraph(const char* path, Rows& R) {                        // :160-209
  std::ifstream in(path);
  if (!in.is_open()) return 3;
  std::string line; char chrom[100]; char extra[100] = "";
  int cs, ce, cov, line_i = 0, prev_end = -1;
  while (std::getline(in, line)) {
    line_i++;
    int items = sscanf(line.c_str(), "%s %d %d %d%s\n", chrom, &cs, &ce, &cov, extra);
    if (items < 4) { printf("problem: %d items on line %d\n", items, line_i); return 4; }
    if (0 < strlen(extra)) return 5;
    if (line_i > 1 && cs != prev_end) return 6;
    prev_end = ce;
    R.start.push_back(cs); R.end.push_back(ce); R.cov.push_back(cov);
  }
  if (line_i == 0) return 9;
  R.chrom = chrom;
  return 0;
}

}  // namespace

extern "C" {

void oracle_set_math(int mode) { g_math_mode = mode; }
void oracle_set_row_hook(row_hook_t h, void* user) { g_hook = h; g_hook_user = user; }

// 64-bit fingerprint of the host libm's exp/log on a fixed input set; tests compare it with the
// value recorded when the golden vectors were generated to know whether "bit-exact vs libm" applies.
uint64_t oracle_libm_fingerprint(void) {
  uint64_t h = 1469598103934665603ULL, s = 88172645463325252ULL;
  for (int i = 0; i < 200000; i++) {
    s ^= s << 13; s ^= s >> 7; s ^= s << 17;
    double u = (double)(s >> 11) * 0x1p-53;
    double e = std::exp(u * 80.0 - 50.0), l = std::log(u * 1e4 + 1e-300);
    uint64_t be, bl; memcpy(&be, &e, 8); memcpy(&bl, &l, 8);
    h = (h ^ be) * 1099511628211ULL; h = (h ^ bl) * 1099511628211ULL;
  }
  return h;
}

// Same contract as the reference's PeakSegFPOP_disk (src/PeakSegFPOPLog.cpp:143): status code,
// <bedGraph>_penalty=<pen>_segments.bed / _loss.tsv, and the scratch db at db_path.
int oracle_fpop_disk(const char* bedgraph, const char* penalty_str, const char* db_path) {
  double penalty; bool is_inf;
  int st = parse_penalty(penalty_str, &penalty, &is_inf);
  if (st) return st;
  Rows R;
  st = read_bedgraph(bedgraph, R);
  if (st) return st;
  std::string prefix = std::string(bedgraph) + "_penalty=" + penalty_str;
  std::ofstream loss_f, seg_f;
  loss_f.open((prefix + "_loss.tsv").c_str());
  seg_f.open((prefix + "_segments.bed").c_str());
  Solution S;
  // the trivial branch never touches the db
  double xmin = kInf, xmax = -kInf;
  for (int z : R.cov) { double lx = xlog((double)z); if (lx < xmin) xmin = lx; if (xmax < lx) xmax = lx; }
  bool trivial = is_inf || xmin == xmax;
  Db db;
  if (!trivial && !db.open(db_path, 2 * (int)R.cov.size())) { db.close(); return 7; }
  try { st = solve_rows(R, penalty, is_inf, trivial ? nullptr : &db, S); }
  catch (const std::exception& e) { db.close(); fprintf(stderr, "oracle: %s\n", e.what()); return 99; }
  db.close();
  if (st) return st;
  for (int k = 0; k < S.n_segments; k++)
    seg_f << R.chrom << "\t" << S.seg_start[k] << "\t" << S.seg_end[k] << "\t"
          << (S.seg_peak[k] ? "peak" : "background") << "\t" << S.seg_mean[k] << "\n";
  loss_f << std::setprecision(20);
  if (S.trivial) loss_f << penalty_str; else loss_f << penalty;
  loss_f << "\t" << S.n_segments << "\t" << S.n_peaks << "\t" << (int)S.bases << "\t" << S.n_rows
         << "\t" << S.best_cost << "\t" << S.total_loss << "\t" << S.n_equality;
  if (S.trivial) loss_f << "\t" << 0 << "\t" << 0 << "\n";
  else loss_f << "\t" << S.total_intervals / (S.n_rows * 2) << "\t" << S.max_intervals << "\n";
  if (loss_f.fail()) return 8;
  if (seg_f.fail()) return 11;
  return 0;
}

// In-memory variant for tests and the CPU baseline: rows in, summary + segments out.
// out_summary[10] = penalty, segments, peaks, bases, rows, mean_pen_cost, total_loss, equality,
//                   mean_intervals, max_intervals.   seg_* arrays must hold n_rows entries.
int oracle_fpop_rows(int n_rows, const int* chrom_start, const int* chrom_end, const int* coverage,
                     double penalty, int penalty_is_inf, double* out_summary,
                     int* seg_start, int* seg_end, int* seg_peak, double* seg_mean) {
  Rows R;
  R.start.assign(chrom_start, chrom_start + n_rows);
  R.end.assign(chrom_end, chrom_end + n_rows);
  R.cov.assign(coverage, coverage + n_rows);
  Solution S;
  int st;
  try { st = solve_rows(R, penalty, penalty_is_inf != 0, nullptr, S); }
  catch (const std::exception& e) { fprintf(stderr, "oracle: %s\n", e.what()); return 99; }
  if (st) return st;
  out_summary[0] = penalty; out_summary[1] = S.n_segments; out_summary[2] = S.n_peaks;
  out_summary[3] = S.bases; out_summary[4] = S.n_rows; out_summary[5] = S.best_cost;
  out_summary[6] = S.total_loss; out_summary[7] = S.n_equality;
  out_summary[8] = S.trivial ? 0 : S.total_intervals / (S.n_rows * 2);
  out_summary[9] = S.trivial ? 0 : S.max_intervals;
  for (int k = 0; k < S.n_segments; k++) {
    seg_start[k] = S.seg_start[k]; seg_end[k] = S.seg_end[k]; seg_peak[k] = S.seg_peak[k]; seg_mean[k] = S.seg_mean[k];
  }
  return 0;
}

}  // extern "C"

#ifdef ORACLE_FPOP_MAIN
int main(int argc, char** argv) {
  if (argc < 4) { fprintf(stderr, "usage: %s bedGraph penalty db [math_mode]\n", argv[0]); return 100; }
  if (argc > 4) oracle_set_math(atoi(argv[4]));
  return oracle_fpop_disk(argv[1], argv[2], argv[3]);
}
#endif
