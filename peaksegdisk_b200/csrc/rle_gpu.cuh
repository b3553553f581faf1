// rle_gpu.cuh -- run-length encoding of count vectors on the device.
//
// Replaces the host-side `rle(count.vec)` / cumsum / data.frame round trip of the reference's
// in-memory front end (R/PeakSegFPOP_vec.R:18-25: rows are (chromStart, chromEnd, count) with
// chromEnd = cumsum(run lengths), chromStart = previous chromEnd, first 0): count vectors are copied
// to HBM as they are (4 B per position) and become the DP kernel's (weight, coverage) rows plus a
// chromEnd array for the backtrack, without the text or row-array detour.
//
// HBM-bound integer work, no tensor cores.  A position is a "head" when it differs from its
// predecessor; its row is the number of heads before it; a head at position i closes the previous
// row: chromEnd[row-1] = i, weight[row-1] = i - (position of the previous head).  ONE kernel, one
// pass over the counts (single-pass scan with decoupled look-back, per vector):
//   * vectors are cut into tiles of 8,192 positions (8 warps x 32 stripes x 32 lanes); a tile never
//     straddles two vectors; blocks take tiles in ticket order (atomic counter), so every
//     predecessor of a running tile is running or done and the look-back cannot deadlock;
//   * a warp loads its 32 stripes as 128-byte coalesced requests (all 32 loads in flight), heads
//     come from shuffles + ballots, ranks from popcounts; the only shared memory is 8 warp totals;
//   * thread 0 publishes the tile's (head count, last head position) as one 64-bit word, looks back
//     over the vector's earlier tiles for its exclusive prefix, then publishes the inclusive one.
// Traffic = algorithmic bytes: 4 B per position read + 12 B per row written.
#pragma once
#include <cuda_runtime.h>

#define PSD_RLE_WARPS 8
#define PSD_RLE_STRIPES 32
#define PSD_RLE_TILE (PSD_RLE_WARPS * 32 * PSD_RLE_STRIPES)

struct RleVec {
  long long raw_off;   // first position in the packed count buffer
  long long row_off;   // first row in the packed row arrays
  int n_pos;
  int tile0;           // id of the vector's first tile
};

struct RleParams {
  const RleVec* vecs;
  const int* tile_vec;   // tile -> vector
  int n_tiles, n_vecs;
  const int* raw;
  int* coverage; int* chrom_end; int* weight;   // indexed by row_off + row
  unsigned long long* tile_state;   // zeroed before the launch; see rle_pack()
  unsigned int* ticket;             // zeroed before the launch
  int* n_rows;                      // per vector
  int* error;                       // set to 1 if a look-back ever gave up (never expected)
};

// tile state word: [63:62] flag (0 empty, 1 aggregate of this tile, 2 inclusive prefix),
// [61:31] head count, [30:0] last head position + 1 (0 = no head so far)
__device__ __forceinline__ unsigned long long rle_pack(unsigned flag, unsigned count, unsigned last1) {
  return ((unsigned long long)flag << 62) | ((unsigned long long)count << 31) | (unsigned long long)last1;
}

__global__ void __launch_bounds__(PSD_RLE_WARPS * 32) rle_encode_kernel(const RleParams P) {
  __shared__ int s_tile;
  __shared__ int s_wtot[PSD_RLE_WARPS];
  __shared__ int s_wlast[PSD_RLE_WARPS];   // last head position + 1 within the warp's range, 0 = none
  __shared__ unsigned s_excl_count, s_excl_last1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) s_tile = (int)atomicAdd(P.ticket, 1u);
  __syncthreads();
  const int tile = s_tile;
  if (tile >= P.n_tiles) return;
  const int vi = P.tile_vec[tile];
  const RleVec V = P.vecs[vi];
  const int* __restrict__ v = P.raw + V.raw_off;
  const int n_pos = V.n_pos;
  const int wbase = (tile - V.tile0) * PSD_RLE_TILE + warp * (32 * PSD_RLE_STRIPES);

  // ---- heads of the warp's 32 stripes --------------------------------------------------------------
  unsigned mask[PSD_RLE_STRIPES]; int val[PSD_RLE_STRIPES];
#pragma unroll
  for (int s = 0; s < PSD_RLE_STRIPES; s++) {
    const int i = wbase + 32 * s + lane;
    val[s] = (i < n_pos) ? __ldcs(v + i) : 0;
  }
  int carry = (wbase > 0 && wbase < n_pos) ? __ldg(v + wbase - 1) : 0;   // the value just before the warp's range
  int total = 0, wlast1 = 0;
#pragma unroll
  for (int s = 0; s < PSD_RLE_STRIPES; s++) {
    const int i = wbase + 32 * s + lane;
    int prev = __shfl_up_sync(0xffffffffu, val[s], 1);
    if (lane == 0) prev = carry;
    const bool head = (i < n_pos) && (i == 0 || val[s] != prev);
    mask[s] = __ballot_sync(0xffffffffu, head);
    carry = __shfl_sync(0xffffffffu, val[s], 31);
    total += __popc(mask[s]);
    if (mask[s]) wlast1 = wbase + 32 * s + (31 - __clz(mask[s])) + 1;
  }
  if (lane == 0) { s_wtot[warp] = total; s_wlast[warp] = wlast1; }
  __syncthreads();

  // ---- tile prefix by decoupled look-back (thread 0) -------------------------------------------------
  if (threadIdx.x == 0) {
    unsigned t_count = 0, t_last1 = 0;
#pragma unroll
    for (int w = 0; w < PSD_RLE_WARPS; w++) { t_count += (unsigned)s_wtot[w]; if (s_wlast[w]) t_last1 = (unsigned)s_wlast[w]; }
    unsigned e_count = 0, e_last1 = 0;
    if (tile > V.tile0) {
      atomicExch(P.tile_state + tile, rle_pack(1u, t_count, t_last1));
      for (int t = tile - 1; t >= V.tile0; t--) {
        unsigned long long w = 0;
        long long spins = 0;
        do {
          w = atomicAdd(P.tile_state + t, 0ull);
          if (++spins > (1ll << 22)) { *P.error = 1; break; }
        } while ((w >> 62) == 0);
        e_count += (unsigned)((w >> 31) & 0x7fffffffu);
        if (e_last1 == 0) e_last1 = (unsigned)(w & 0x7fffffffu);
        if ((w >> 62) != 1) break;   // an inclusive prefix ends the walk
      }
    }
    atomicExch(P.tile_state + tile, rle_pack(2u, e_count + t_count, t_last1 ? t_last1 : e_last1));
    s_excl_count = e_count; s_excl_last1 = e_last1;
    const bool last_tile = (tile - V.tile0 + 1) * PSD_RLE_TILE >= n_pos;
    if (last_tile) {   // the vector's final row ends at n_pos
      const int n_rows = (int)(e_count + t_count);
      const int last_head = (int)(t_last1 ? t_last1 : e_last1) - 1;
      P.n_rows[vi] = n_rows;
      P.chrom_end[V.row_off + n_rows - 1] = n_pos;
      P.weight[V.row_off + n_rows - 1] = n_pos - last_head;
    }
  }
  __syncthreads();

  // ---- scatter: coverage of this row, chromEnd and weight of the row it closes ---------------------
  int rank = (int)s_excl_count;
  int prev1 = (int)s_excl_last1;            // last head position + 1 before the warp's range
#pragma unroll
  for (int w = 0; w < PSD_RLE_WARPS; w++) {
    if (w < warp) { rank += s_wtot[w]; if (s_wlast[w]) prev1 = s_wlast[w]; }
  }
  int* cov = P.coverage + V.row_off;
  int* end = P.chrom_end + V.row_off;
  int* wgt = P.weight + V.row_off;
  const unsigned lt = (1u << lane) - 1u;
#pragma unroll
  for (int s = 0; s < PSD_RLE_STRIPES; s++) {
    const unsigned m = mask[s];
    if ((m >> lane) & 1u) {
      const int i = wbase + 32 * s + lane;
      const unsigned lower = m & lt;
      const int row = rank + __popc(lower);
      const int prev_head = lower ? wbase + 32 * s + (31 - __clz(lower)) : prev1 - 1;
      cov[row] = val[s];
      if (row > 0) { end[row - 1] = i; wgt[row - 1] = i - prev_head; }
    }
    rank += __popc(m);
    if (m) prev1 = wbase + 32 * s + (31 - __clz(m)) + 1;
  }
}
