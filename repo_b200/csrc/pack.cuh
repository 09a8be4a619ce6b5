// Weight / bias packing: fp32 nn.Linear / nn.GRUCell tensors -> the shared-memory image the
// layer machine streams (vm.cuh).  One 128-feature tile of a layer becomes `ksl` slabs of 8 KB:
//   slab = [hi: 2 k-groups x 128 rows x 8 fp16][lo: same]   (tcgen05 K-major core matrices,
//   LBO = 2048 B between k-groups, SBO = 128 B between 8-row groups)
// so a slab lands in smem with a single 1-D bulk copy and needs no swizzle.
#pragma once
#include "ptx.cuh"

namespace rb {

struct PackJob {   // one 128-row tile of one weight matrix
  const float* w;  // source matrix, row-major
  int ld;          // its row stride
  int row0, nrows; // source rows [row0, row0+nrows), nrows <= 128 (rest of the tile is zero)
  int col0, ncols; // source cols [col0, col0+ncols)
  int kofs;        // destination k of source column col0 (columns outside are zero)
  int ksl;         // k16 slabs in the tile
  uint32_t w_slab; // first destination slab
  int blk0;        // first block of this job (prefix sum of ksl)
};
struct BiasJob {   // one 128-float bias tile: dst[i] = a[a_off+i] (+ b[b_off+i]) for i < n else 0
  const float* a;
  const float* b;
  int a_off, b_off, n, dst_tile;
};
constexpr int kMaxPackJobs = 48;
struct PackArgs {
  PackJob jobs[kMaxPackJobs];
  int n_jobs;
  uint8_t* wblob;
};
struct BiasArgs {
  BiasJob jobs[kMaxPackJobs];
  int n_jobs;
  float* bias;
};

__global__ void __launch_bounds__(128) pack_weights_kernel(const __grid_constant__ PackArgs a) {
  int ji = 0;
  while (ji + 1 < a.n_jobs && (int)blockIdx.x >= a.jobs[ji + 1].blk0) ++ji;
  const PackJob& j = a.jobs[ji];
  const int slab = blockIdx.x - j.blk0;
  const int m = threadIdx.x;
  uint8_t* dst = a.wblob + (size_t)(j.w_slab + slab) * 8192;
  const bool vrow = m < j.nrows;
  const float* src = j.w + (size_t)(j.row0 + (vrow ? m : 0)) * j.ld + j.col0;
#pragma unroll
  for (int kg = 0; kg < 2; ++kg) {
    __align__(16) __half hi[8];
    __align__(16) __half lo[8];
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
      const int kc = slab * 16 + kg * 8 + kk - j.kofs;
      const float v = (vrow && kc >= 0 && kc < j.ncols) ? src[kc] : 0.f;
      split_f16(v, hi[kk], lo[kk]);
    }
    *reinterpret_cast<uint4*>(dst + kg * 2048 + m * 16) = *reinterpret_cast<const uint4*>(hi);
    *reinterpret_cast<uint4*>(dst + 4096 + kg * 2048 + m * 16) = *reinterpret_cast<const uint4*>(lo);
  }
}

__global__ void __launch_bounds__(128) pack_bias_kernel(const __grid_constant__ BiasArgs a) {
  const BiasJob& j = a.jobs[blockIdx.x];
  const int i = threadIdx.x;
  float v = 0.f;
  if (i < j.n) {
    if (j.a) v = j.a[j.a_off + i];
    if (j.b) v += j.b[j.b_off + i];
  }
  a.bias[j.dst_tile * 128 + i] = v;
}


// ---------------------------------------------------------------------------------------------
// "rows on M" machine (rows.cuh): weights are the B operand.  One job = one GEMM's weight matrix
// W'(n_pad x 16*ksl), assembled from up to three row segments of the source (e.g. the r/z/n
// gate rows of one GRU unit chunk).  Slab = [hi: 2 k-groups x n_pad rows x 16 B][lo: same].
struct PackRowsJob {
  const float* w;
  int ld, col0, ncols, kofs, ksl, n_pad, nseg;
  int seg_src[3], seg_n[3], seg_dst[3];
  uint32_t dst_off16;  // destination offset / 16
  int blk0;
};
struct BiasRowsJob {   // dst[dst_off + i] = (i < n) ? a[a_off+i] (+ b[b_off+i]) : 0   for i < n_pad
  const float* a;
  const float* b;
  int a_off, b_off, n, n_pad, dst_off;
};
constexpr int kMaxRowsJobs = 64;
struct PackRowsArgs {
  PackRowsJob jobs[kMaxRowsJobs];
  int n_jobs;
  uint8_t* wblob;
  const float* scale;  // optional device scalar: weights are multiplied by it before the fp16 split (conv.cuh)
};
struct BiasRowsArgs {
  BiasRowsJob jobs[64];
  int n_jobs;
  float* bias;
};

__device__ __forceinline__ void pack_rows_weights_block(const PackRowsArgs& a, int block) {
  int ji = 0;
  while (ji + 1 < a.n_jobs && block >= a.jobs[ji + 1].blk0) ++ji;
  const PackRowsJob& j = a.jobs[ji];
  const int slab = block - j.blk0;
  const int m = threadIdx.x;
  if (m >= j.n_pad) return;
  int src_row = -1;
  for (int s = 0; s < j.nseg; ++s)
    if (m >= j.seg_dst[s] && m < j.seg_dst[s] + j.seg_n[s]) src_row = j.seg_src[s] + (m - j.seg_dst[s]);
  uint8_t* dst = a.wblob + (size_t)j.dst_off16 * 16 + (size_t)slab * j.n_pad * 64;
  const float* src = j.w + (size_t)(src_row < 0 ? 0 : src_row) * j.ld + j.col0;
  const float ws = a.scale ? *a.scale : 1.f;
#pragma unroll
  for (int kg = 0; kg < 2; ++kg) {
    __align__(16) __half hi[8];
    __align__(16) __half lo[8];
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
      const int kc = slab * 16 + kg * 8 + kk - j.kofs;
      const float v = (src_row >= 0 && kc >= 0 && kc < j.ncols) ? src[kc] * ws : 0.f;
      split_f16(v, hi[kk], lo[kk]);
    }
    *reinterpret_cast<uint4*>(dst + kg * j.n_pad * 16 + m * 16) = *reinterpret_cast<const uint4*>(hi);
    *reinterpret_cast<uint4*>(dst + j.n_pad * 32 + kg * j.n_pad * 16 + m * 16) = *reinterpret_cast<const uint4*>(lo);
  }
}

__device__ __forceinline__ void pack_rows_bias_block(const BiasRowsArgs& a, int block) {
  const BiasRowsJob& j = a.jobs[block];
  for (int i = threadIdx.x; i < j.n_pad; i += blockDim.x) {
    float v = 0.f;
    if (i < j.n) {
      if (j.a) v = j.a[j.a_off + i];
      if (j.b) v += j.b[j.b_off + i];
    }
    a.bias[j.dst_off + i] = v;
  }
}

__global__ void __launch_bounds__(256) pack_rows_weights_kernel(const __grid_constant__ PackRowsArgs a) {
  pack_rows_weights_block(a, (int)blockIdx.x);
}
__global__ void __launch_bounds__(256) pack_rows_bias_kernel(const __grid_constant__ BiasRowsArgs a) {
  pack_rows_bias_block(a, (int)blockIdx.x);
}
// weights and biases of one launch's operands in ONE kernel: blocks [0, n_weight_blocks) pack weight slabs, the rest biases
// (a dense layer / conv call is then two launches — pack + GEMM — instead of three; ~90 such calls per training iteration)
__global__ void __launch_bounds__(256) pack_rows_both_kernel(const __grid_constant__ PackRowsArgs a, const __grid_constant__ BiasRowsArgs b,
                                                             int n_weight_blocks) {
  if ((int)blockIdx.x < n_weight_blocks) pack_rows_weights_block(a, (int)blockIdx.x);
  else pack_rows_bias_block(b, (int)blockIdx.x - n_weight_blocks);
}

}  // namespace rb
