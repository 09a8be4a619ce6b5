// Optimiser tail of every update (dreamer.py:286-289, 356-359, 370-373; repo.py:87-96):
// clip_grad_norm_(params, max_norm) followed by Adam.step, over ONE flat fp32 bucket per parameter group
// (the same bucket the data-parallel all-reduce works on), with no host synchronisation: the squared norm
// stays on the device and the clip coefficient is recomputed by every thread of the update kernel.
#pragma once
#include <cuda_runtime.h>

namespace rb {

__global__ void __launch_bounds__(256) sqnorm_kernel(const float* __restrict__ g, long long n, float* __restrict__ out) {
  float acc = 0.f;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float v = g[i];
    acc = fmaf(v, v, acc);
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
  __shared__ float warp_sums[8];
  if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 8) {
    float v = warp_sums[threadIdx.x];
#pragma unroll
    for (int s = 4; s > 0; s >>= 1) v += __shfl_xor_sync(0xffu, v, s);
    if (threadIdx.x == 0) atomicAdd(out, v);
  }
}

// Bias gradients: out[c] += sum over this block's row slice of x[r * ld + c] (out zeroed by the caller).  The trainers need ~35
// of these per iteration over tall, narrow matrices ((T*B or (H-1)*N) x 12..600); ATen's generic reduce takes ~33 us for each.
// One warp reads 32 consecutive columns of a row; the 8 warps of a block stride over the rows of the slice.
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ x, long long rows, int cols, long long ld,
                                                     float* __restrict__ out) {
  const int c = blockIdx.x * 32 + (threadIdx.x & 31), w = threadIdx.x >> 5;
  const long long per = (rows + gridDim.y - 1) / gridDim.y, r0 = (long long)blockIdx.y * per, r1 = min(rows, r0 + per);
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  if (c < cols) {
    long long r = r0 + w;
    for (; r + 24 < r1; r += 32) {   // four independent loads in flight
      a0 += x[r * ld + c]; a1 += x[(r + 8) * ld + c]; a2 += x[(r + 16) * ld + c]; a3 += x[(r + 24) * ld + c];
    }
    for (; r < r1; r += 8) a0 += x[r * ld + c];
  }
  __shared__ float part[8][32];
  part[w][threadIdx.x & 31] = (a0 + a1) + (a2 + a3);
  __syncthreads();
  if (w == 0 && c < cols) {
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) v += part[i][threadIdx.x];
    atomicAdd(out + c, v);
  }
}

// torch semantics: clip_coef = min(1, max_norm / (||g|| + 1e-6)); Adam with bias correction
//   m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
__global__ void __launch_bounds__(256) adam_clip_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                                                        float* __restrict__ v, long long n, const float* __restrict__ sqnorm,
                                                        float max_norm, float lr, float b1, float b2, float eps,
                                                        float bc1, float bc2_sqrt) {
  float coef = 1.f;
  if (sqnorm && max_norm > 0.f) coef = fminf(1.f, max_norm / (sqrtf(*sqnorm) + 1e-6f));
  const float step_size = lr / bc1;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float gi = g[i] * coef;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    g[i] = gi;  // like clip_grad_norm_, the clipped gradient is left in place
    p[i] -= step_size * (mi / (sqrtf(vi) / bc2_sqrt + eps));
  }
}

// Graph-safe variant: the step count lives on the device.  `adam_step_inc_kernel` (one thread) advances it, then every
// thread of the update derives the bias corrections from it, so a captured CUDA graph replays with the right step.
__global__ void adam_step_inc_kernel(int* step) { *step += 1; }

__global__ void __launch_bounds__(256) adam_clip_dev_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                                                            float* __restrict__ v, long long n, const float* __restrict__ sqnorm,
                                                            float max_norm, float lr, float b1, float b2, float eps,
                                                            const int* __restrict__ step) {
  float coef = 1.f;
  if (sqnorm && max_norm > 0.f) coef = fminf(1.f, max_norm / (sqrtf(*sqnorm) + 1e-6f));
  const double t = (double)*step;
  const float bc1 = (float)(1.0 - pow((double)b1, t)), bc2_sqrt = (float)sqrt(1.0 - pow((double)b2, t));
  const float step_size = lr / bc1;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float gi = g[i] * coef;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    g[i] = gi;
    p[i] -= step_size * (mi / (sqrtf(vi) / bc2_sqrt + eps));
  }
}

}  // namespace rb
