// Test tool: bit-compare psd_exp/psd_log (peaksegdisk_b200/csrc/psd_math.h) with the system libm.
// usage: math_check <n_random> [seed]   -> prints mismatch counts; exit 0 iff none.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../peaksegdisk_b200/csrc/psd_math.h"

static uint64_t s[2];
static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static uint64_t next(void) {
  uint64_t s0 = s[0], s1 = s[1], r = s0 + s1;
  s1 ^= s0; s[0] = rotl(s0, 24) ^ s1 ^ (s1 << 16); s[1] = rotl(s1, 37);
  return r;
}
static long bad_exp = 0, bad_log = 0, n_exp = 0, n_log = 0;
static int same(double a, double b) {
  if (a != a && b != b) return 1;  // any NaN == any NaN
  return PSD_D2U(a) == PSD_D2U(b);
}
static void chk_exp(double x) {
  double a = exp(x), b = psd_exp(x, psd_exp_tab_host);
  n_exp++;
  if (!same(a, b)) { if (bad_exp++ < 10) printf("EXP mismatch x=%a libm=%a psd=%a\n", x, a, b); }
}
static void chk_log(double x) {
  double a = log(x), b = psd_log(x, psd_log_tab_host);
  n_log++;
  if (!same(a, b)) { if (bad_log++ < 10) printf("LOG mismatch x=%a libm=%a psd=%a\n", x, a, b); }
}
int main(int argc, char** argv) {
  long n = argc > 1 ? atol(argv[1]) : 1000000;
  s[0] = argc > 2 ? strtoull(argv[2], 0, 10) : 12345; s[1] = 0x9e3779b97f4a7c15ULL;
  // specials
  double sp[] = {0.0, -0.0, 1.0, -1.0, INFINITY, -INFINITY, NAN, 0x1p-1074, 0x1p-1022, 0x1.fffffffffffffp1023,
                 709.782712893384, 709.782712893385, -745.1332191019411, -745.1332191019412, -708.3964185322641,
                 -708.4, -744.0, 1024.0, -1024.0, 512.0, -512.0, 0x1p-54, 0x1p-55, -0x1p-54, 0.9375, 1.0647, 0.6875, 1.375};
  for (unsigned i = 0; i < sizeof sp / sizeof sp[0]; i++) { chk_exp(sp[i]); chk_log(sp[i]); chk_log(-sp[i]); chk_exp(-sp[i]); }
  // all integers (log of coverage values) and their ratios
  for (long k = 0; k <= 5000000; k++) chk_log((double)k);
  for (long i = 0; i < n; i++) {
    uint64_t r = next();
    // 1) any bit pattern
    chk_exp(PSD_U2D(r)); chk_log(PSD_U2D(r));
    // 2) uniform in [-750, 750] (covers under/overflow + subnormal results)
    double u = (double)(next() >> 11) * 0x1p-53;
    chk_exp((u - 0.5) * 1500.0);
    // 3) the solver's range: log-means in [-40, 25]
    double v = (double)(next() >> 11) * 0x1p-53;
    chk_exp(v * 65.0 - 40.0);
    // 4) log of positive numbers across all exponents, and near 1
    double w = (double)(next() >> 11) * 0x1p-53;
    chk_log(PSD_U2D(r & 0x7fffffffffffffffULL));
    chk_log(0.9 + 0.2 * w);
    chk_log(w * 1e6);
    chk_log(ldexp(1.0 + w, (int)(next() % 40) - 20));
    // 5) subnormal inputs to log, subnormal outputs of exp
    chk_log(PSD_U2D(r & 0x000fffffffffffffULL));
    chk_exp(-708.0 - 38.0 * w);
    chk_exp(700.0 + 10.0 * w);
  }
  printf("exp: %ld checked, %ld mismatches; log: %ld checked, %ld mismatches\n", n_exp, bad_exp, n_log, bad_log);
  return (bad_exp || bad_log) ? 1 : 0;
}
