// rle_gpu.cuh -- run-length encoding of count vectors on the device.
//
// Replaces the host-side `rle(count.vec)` / cumsum / data.frame round trip of the reference's
// in-memory front end (R/PeakSegFPOP_vec.R:18-25: rows are (chromStart, chromEnd, count) with
// chromEnd = cumsum(run lengths), chromStart = previous chromEnd, first 0): count vectors are copied
// to HBM as they are (4 B per position) and become the DP kernel's (weight, coverage) rows plus a
// chromEnd array for the backtrack, without the text or row-array detour.
//
// HBM-bound integer work, no tensor cores: a position is a "head" when it differs from its
// predecessor; row index = number of heads before it.  Vectors are cut into tiles of 8,192
// positions (8 warps x 32 stripes x 32 lanes); a tile never straddles two vectors.
//   pass 1  rle_count_kernel    heads per tile                               reads 4 B/position
//   pass 2  rle_scan_kernel     per vector: exclusive scan of its tile counts, n_rows, last chromEnd
//   pass 3  rle_scatter_kernel  coverage[row], chromEnd[row-1]               reads 4 B/position (L2), writes 8 B/row
//   pass 4  rle_weight_kernel   weight[row] = chromEnd[row] - chromEnd[row-1]    8 B/row
// All loads of a warp are 128-byte coalesced stripes; ranks come from ballots, no shared-memory scan
// beyond the 8 warp totals of a block.  Algorithmic bytes: 4 per position + 12 per row.
#pragma once
#include <cuda_runtime.h>

#define PSD_RLE_WARPS 8
#define PSD_RLE_STRIPES 32
#define PSD_RLE_TILE (PSD_RLE_WARPS * 32 * PSD_RLE_STRIPES)

struct RleVec {
  long long raw_off;   // first position in the packed count buffer
  long long row_off;   // first row in the packed row arrays (capacity n_pos rows)
  int n_pos;
  int tile0;           // id of the vector's first tile
};

struct RleParams {
  const RleVec* vecs;
  const int* tile_vec;   // tile -> vector
  int n_tiles, n_vecs;
  const int* raw;
  int* coverage; int* chrom_end; int* weight;   // indexed by row_off + row
  int* tile_count;       // heads per tile; exclusive prefix within the vector after pass 2
  int* n_rows;           // per vector
};

// Head masks of the 32 stripes a warp owns: stripe s covers positions wbase + 32*s + lane.
__device__ __forceinline__ int rle_warp_heads(const int* __restrict__ v, int n_pos, int wbase, int lane,
                                              unsigned (&mask)[PSD_RLE_STRIPES], int (&val)[PSD_RLE_STRIPES]) {
#pragma unroll
  for (int s = 0; s < PSD_RLE_STRIPES; s++) {
    const int i = wbase + 32 * s + lane;
    val[s] = (i < n_pos) ? __ldg(v + i) : 0;
  }
  int carry = (wbase > 0 && wbase < n_pos) ? __ldg(v + wbase - 1) : 0;   // value just before the warp's range
  int total = 0;
#pragma unroll
  for (int s = 0; s < PSD_RLE_STRIPES; s++) {
    const int i = wbase + 32 * s + lane;
    int prev = __shfl_up_sync(0xffffffffu, val[s], 1);
    if (lane == 0) prev = carry;
    const bool head = (i < n_pos) && (i == 0 || val[s] != prev);
    mask[s] = __ballot_sync(0xffffffffu, head);
    carry = __shfl_sync(0xffffffffu, val[s], 31);
    total += __popc(mask[s]);
  }
  return total;
}

__global__ void __launch_bounds__(PSD_RLE_WARPS * 32) rle_count_kernel(const RleParams P) {
  __shared__ int wtot[PSD_RLE_WARPS];
  const int tile = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int vi = P.tile_vec[tile];
  const RleVec V = P.vecs[vi];
  const int wbase = (tile - V.tile0) * PSD_RLE_TILE + warp * (32 * PSD_RLE_STRIPES);
  unsigned mask[PSD_RLE_STRIPES]; int val[PSD_RLE_STRIPES];
  const int total = rle_warp_heads(P.raw + V.raw_off, V.n_pos, wbase, lane, mask, val);
  if (lane == 0) wtot[warp] = total;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
#pragma unroll
    for (int w = 0; w < PSD_RLE_WARPS; w++) t += wtot[w];
    P.tile_count[tile] = t;
  }
}

// One warp per vector: exclusive scan of its tile counts (in place), row count, chromEnd of the last row.
__global__ void rle_scan_kernel(const RleParams P) {
  const int vi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (vi >= P.n_vecs) return;
  const RleVec V = P.vecs[vi];
  const int nt = (V.n_pos + PSD_RLE_TILE - 1) / PSD_RLE_TILE;
  int running = 0;
  for (int t0 = 0; t0 < nt; t0 += 32) {
    const int t = t0 + lane;
    const int c = (t < nt) ? P.tile_count[V.tile0 + t] : 0;
    int incl = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += o; }
    if (t < nt) P.tile_count[V.tile0 + t] = running + incl - c;
    running += __shfl_sync(0xffffffffu, incl, 31);
  }
  if (lane == 0) {
    P.n_rows[vi] = running;
    if (running > 0) P.chrom_end[V.row_off + running - 1] = V.n_pos;
  }
}

__global__ void __launch_bounds__(PSD_RLE_WARPS * 32) rle_scatter_kernel(const RleParams P) {
  __shared__ int wtot[PSD_RLE_WARPS];
  const int tile = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int vi = P.tile_vec[tile];
  const RleVec V = P.vecs[vi];
  const int wbase = (tile - V.tile0) * PSD_RLE_TILE + warp * (32 * PSD_RLE_STRIPES);
  unsigned mask[PSD_RLE_STRIPES]; int val[PSD_RLE_STRIPES];
  const int total = rle_warp_heads(P.raw + V.raw_off, V.n_pos, wbase, lane, mask, val);
  if (lane == 0) wtot[warp] = total;
  __syncthreads();
  int rank = P.tile_count[tile];
#pragma unroll
  for (int w = 0; w < PSD_RLE_WARPS; w++) rank += (w < warp) ? wtot[w] : 0;
  int* cov = P.coverage + V.row_off;
  int* end = P.chrom_end + V.row_off;
  const unsigned lt = (1u << lane) - 1u;
#pragma unroll
  for (int s = 0; s < PSD_RLE_STRIPES; s++) {
    if ((mask[s] >> lane) & 1u) {
      const int row = rank + __popc(mask[s] & lt);
      cov[row] = val[s];
      if (row > 0) end[row - 1] = wbase + 32 * s + lane;   // this run starts where the previous one ends
    }
    rank += __popc(mask[s]);
  }
}

__global__ void __launch_bounds__(PSD_RLE_WARPS * 32) rle_weight_kernel(const RleParams P) {
  const int tile = blockIdx.x;
  const int vi = P.tile_vec[tile];
  const RleVec V = P.vecs[vi];
  const int n_rows = P.n_rows[vi];
  const int base = (tile - V.tile0) * PSD_RLE_TILE;
  if (base >= n_rows) return;
  const int* end = P.chrom_end + V.row_off;
  int* w = P.weight + V.row_off;
  for (int r = base + threadIdx.x; r < base + PSD_RLE_TILE && r < n_rows; r += blockDim.x)
    w[r] = end[r] - (r ? end[r - 1] : 0);
}
