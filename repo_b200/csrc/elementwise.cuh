// HBM-bound helpers around the recurrence: the Monte-Carlo tanh-Normal entropy (SampleDist.entropy,
// models/utils.py:160-163 over TanhBijector :112-134) and the replay gather + preprocess
// (SequenceReplayBuffer.sample common/buffers.py:156-166 + preprocess common/utils.py:74-80).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace rb {

// entropy[m] = -(1/K) sum_k sum_a log p(tanh(mean + std*eps[k,m,a]))
//   log p(y) = Normal(mean,std).log_prob(atanh(clamp(y))) - 2(log2 - x - softplus(-2x)),  x = atanh(clamp(y))
// One thread per (m, sample-slice); eps is read once, coalesced along m*A+a (HBM-bound: K*M*A*4 bytes).
constexpr int kEntSlices = 4;  // threads cooperating on one row (split over k), reduced with shuffles
__global__ void __launch_bounds__(256) tanh_normal_entropy_kernel(const float* __restrict__ mean,
                                                                  const float* __restrict__ std_,
                                                                  const float* __restrict__ eps, float* __restrict__ out,
                                                                  int M, int A, int K) {
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  const int m = gid / kEntSlices, sl = gid % kEntSlices;
  float acc = 0.f;
  if (m < M) {
    for (int a = 0; a < A; ++a) {
      const float mu = mean[(size_t)m * A + a], sd = std_[(size_t)m * A + a];
      const float inv2var = 1.f / (2.f * sd * sd);
      const float c0 = logf(sd) + 0.9189385332046727f;  // log(std) + log(sqrt(2*pi))
      for (int k = sl; k < K; k += kEntSlices) {
        const float e = __ldg(eps + ((size_t)k * M + m) * A + a);
        const float y = tanhf(mu + sd * e);
        const float yc = fabsf(y) <= 1.f ? fminf(fmaxf(y, -0.99999997f), 0.99999997f) : y;
        const float x = atanhf(yc);
        const float d = x - mu;
        const float t = -2.f * x;
        const float sp = t > 20.f ? t : log1pf(expf(t));  // F.softplus(-2x)
        const float ladj = 2.f * (0.6931471805599453f - x - sp);
        acc += (-(d * d) * inv2var - c0) - ladj;
      }
    }
  }
#pragma unroll
  for (int s = 1; s < kEntSlices; s <<= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
  if (m < M && sl == 0) out[m] = -acc / (float)K;
}

// Replay gather: time-major flat index i = l*B + b -> ring slot (start[b] + l (+ pos) mod length); one block
// per (l,b) copies the frame, applying x/255*2-1 in numpy's exact operation order (no FMA contraction).
__global__ void __launch_bounds__(128) replay_gather_kernel(const uint8_t* __restrict__ obs, const float* __restrict__ act,
                                                            const float* __restrict__ rew, const float* __restrict__ done,
                                                            const long long* __restrict__ starts, int B, int L, long long pos,
                                                            int full, long long length, int frame_bytes, int act_dim,
                                                            float* __restrict__ obs_out, float* __restrict__ act_out,
                                                            float* __restrict__ rew_out, float* __restrict__ nonterm_out,
                                                            long long* __restrict__ index_out) {
  const int i = blockIdx.x;  // l*B + b
  const int l = i / B, b = i - l * B;
  long long idx = starts[b] + l;
  if (full) idx = (idx + pos) % length;
  const uint8_t* src = obs + (size_t)idx * frame_bytes;
  float* dst = obs_out + (size_t)i * frame_bytes;
  if ((frame_bytes & 15) == 0 && ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0) {
    const uint4* s4 = reinterpret_cast<const uint4*>(src);
    for (int v = threadIdx.x; v < frame_bytes / 16; v += blockDim.x) {
      const uint4 q = __ldg(s4 + v);
      const uint32_t w[4] = {q.x, q.y, q.z, q.w};
      float4* d4 = reinterpret_cast<float4*>(dst + (size_t)v * 16);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float4 o;
        o.x = __fsub_rn(__fmul_rn(__fdiv_rn((float)(w[j] & 0xff), 255.f), 2.f), 1.f);
        o.y = __fsub_rn(__fmul_rn(__fdiv_rn((float)((w[j] >> 8) & 0xff), 255.f), 2.f), 1.f);
        o.z = __fsub_rn(__fmul_rn(__fdiv_rn((float)((w[j] >> 16) & 0xff), 255.f), 2.f), 1.f);
        o.w = __fsub_rn(__fmul_rn(__fdiv_rn((float)(w[j] >> 24), 255.f), 2.f), 1.f);
        d4[j] = o;
      }
    }
  } else {
    for (int v = threadIdx.x; v < frame_bytes; v += blockDim.x)
      dst[v] = __fsub_rn(__fmul_rn(__fdiv_rn((float)src[v], 255.f), 2.f), 1.f);
  }
  for (int v = threadIdx.x; v < act_dim; v += blockDim.x) act_out[(size_t)i * act_dim + v] = act[(size_t)idx * act_dim + v];
  if (threadIdx.x == 0) {
    rew_out[i] = rew[idx];
    nonterm_out[i] = 1.f - done[idx];
    if (index_out) index_out[i] = idx;
  }
}

}  // namespace rb

#include "vm.cuh"

namespace rb {

// Backward helper of the implicit-GEMM convolutions (same ConvMap as the forward gather in conv.cuh):
// im2col materialises the gathered rows, needed once per layer for the weight-gradient GEMM dW = g^T col.
__global__ void __launch_bounds__(256) im2col_kernel(const float* __restrict__ in, float* __restrict__ col, long long rows,
                                                     const __grid_constant__ ConvMap cm) {
  const int K = cm.ntaps * cm.C;
  const long long total = rows * K;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    const long long row = idx / K;
    const int kk = (int)(idx - row * K), tl = kk / cm.C, ci = kk - tl * cm.C;
    const int tap = cm.tap0 + tl, ty = tap / cm.TW, tx = tap - ty * cm.TW;
    const int per = cm.RA * cm.RB;
    const int fr = (int)(row / per), rem = (int)(row - (long long)fr * per), a = rem / cm.RB, b = rem - a * cm.RB;
    const int iy = a * cm.sy + ty * cm.dy + cm.y0, ix = b * cm.sx + tx * cm.dx + cm.x0;
    float v = 0.f;
    if (iy >= 0 && iy < cm.H && ix >= 0 && ix < cm.W)
      v = cm.in_nchw ? in[(((size_t)fr * cm.C + ci) * cm.H + iy) * cm.W + ix] : in[(((size_t)fr * cm.H + iy) * cm.W + ix) * cm.pix + ci];
    col[idx] = v;
  }
}

// Power-of-two operand scales for the gradient convolutions (conv.cuh): one read-only pass finds max|x|, the
// last block to finish turns it into s = 2^floor(log2(target / max|x|)).  scales = [s_a, s_b, 1/(s_a*s_b)] followed by the
// same triple with a and b swapped (6 floats); `which` selects the slot this tensor fills (the other one must already
// hold its value, 1 by default).
__global__ void __launch_bounds__(256) absmax_scale_kernel(const float* __restrict__ x, long long n, unsigned* scratch /* [2]: max bits, blocks done */,
                                                           float target, float* __restrict__ scales, int which) {
  float m = 0.f;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long n4 = n >> 2;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = __ldg(x4 + i);
    m = fmaxf(fmaxf(m, fabsf(v.x)), fmaxf(fmaxf(fabsf(v.y), fabsf(v.z)), fabsf(v.w)));
  }
  for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) m = fmaxf(m, fabsf(x[i]));
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, s));
  __shared__ float wm[8];
  __shared__ bool last;
  if ((threadIdx.x & 31) == 0) wm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) m = fmaxf(m, wm[i]);
    atomicMax(scratch, __float_as_uint(m));   // non-negative floats order like their bit patterns
    __threadfence();
    last = atomicAdd(scratch + 1, 1u) == gridDim.x - 1;
    if (last) {
      const float amax = fmaxf(__uint_as_float(atomicMax(scratch, 0u)), 1e-30f);
      const float s = exp2f(floorf(log2f(target / amax)));
      scales[which] = s;
      scales[2] = 1.f / (s * scales[1 - which]);
      // second triple with the operand slots swapped (the same tensor used as the other operand)
      scales[3 + (1 - which)] = s;
      scales[3 + which] = scales[1 - which];
      scales[5] = scales[2];
      scratch[0] = 0u;
      scratch[1] = 0u;   // ready for the next call on this stream
    }
  }
}

// TIA mask mixing (tia.py:72,124-127): mask = sigmoid(Conv2d(6,1,1)(cat(t_mask, d_mask))), recon = t*mask + d*(1-mask).
// t_out / d_out are the (F,6,H,W) NCHW outputs of the two TIAObservationModels: channels 0-2 reconstruction, 3-5 mask
// features.  One thread per pixel.
__global__ void __launch_bounds__(256) tia_mix_fwd_kernel(const float* __restrict__ t_out, const float* __restrict__ d_out,
                                                          const float* __restrict__ w /* 6 */, const float* __restrict__ b,
                                                          float* __restrict__ recon /* (F,3,H,W) */, float* __restrict__ mask /* (F,H,W) */,
                                                          long long frames, int hw) {
  const long long total = frames * hw;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    const long long fr = idx / hw;
    const int px = (int)(idx - fr * hw);
    const float* tp = t_out + fr * 6 * hw + px;
    const float* dp = d_out + fr * 6 * hw + px;
    float z = b[0];
#pragma unroll
    for (int c = 0; c < 3; ++c) z += w[c] * tp[(3 + c) * hw] + w[3 + c] * dp[(3 + c) * hw];
    const float m = 1.f / (1.f + expf(-z));
    mask[idx] = m;
#pragma unroll
    for (int c = 0; c < 3; ++c) recon[(fr * 3 + c) * hw + px] = tp[c * hw] * m + dp[c * hw] * (1.f - m);
  }
}

// backward: g (F,3,H,W) -> d_t_out, d_d_out (F,6,H,W) and d_w[6], d_b (atomically accumulated; zero them first)
__global__ void __launch_bounds__(256) tia_mix_bwd_kernel(const float* __restrict__ t_out, const float* __restrict__ d_out,
                                                          const float* __restrict__ w, const float* __restrict__ mask,
                                                          const float* __restrict__ g, float* __restrict__ g_t, float* __restrict__ g_d,
                                                          float* __restrict__ g_wb /* 7 */, long long frames, int hw) {
  const long long total = frames * hw;
  const long long stride = (long long)gridDim.x * blockDim.x;
  float acc[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    const long long fr = idx / hw;
    const int px = (int)(idx - fr * hw);
    const float* tp = t_out + fr * 6 * hw + px;
    const float* dp = d_out + fr * 6 * hw + px;
    float* gt = g_t + fr * 6 * hw + px;
    float* gd = g_d + fr * 6 * hw + px;
    const float m = mask[idx];
    float dm = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float gc = g[(fr * 3 + c) * hw + px];
      gt[c * hw] = gc * m;
      gd[c * hw] = gc * (1.f - m);
      dm += gc * (tp[c * hw] - dp[c * hw]);
    }
    const float dz = dm * m * (1.f - m);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      gt[(3 + c) * hw] = dz * w[c];
      gd[(3 + c) * hw] = dz * w[3 + c];
      acc[c] += dz * tp[(3 + c) * hw];
      acc[3 + c] += dz * dp[(3 + c) * hw];
    }
    acc[6] += dz;
  }
  __shared__ float red[8][7];
#pragma unroll
  for (int i = 0; i < 7; ++i) {
    float v = acc[i];
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][i] = v;
  }
  __syncthreads();
  if (threadIdx.x < 7) {
    float v = 0.f;
    for (int wv = 0; wv < 8; ++wv) v += red[wv][threadIdx.x];
    atomicAdd(g_wb + threadIdx.x, v);
  }
}

// Backward glue of the sub-pixel transposed convolution (conv.py): G[(f,a,b), (py,px,c)] = g[f, 2a+py, 2b+px, c]
// (zero outside the Ho x Wo grid and in the channel padding up to cpad), plus the bias gradient db[c] = sum g — one
// pass over the gradient instead of pad / permute / contiguous / sum.  Thread t owns column n = t % cpad (block and
// grid strides are multiples of cpad), so its partial bias sum stays in a register.
__global__ void __launch_bounds__(256) grad_unshuffle_kernel(const float* __restrict__ g, int g_nchw, float* __restrict__ G,
                                                             float* __restrict__ db /* C, zeroed */, long long rows /* F*RA*RB */,
                                                             int RA, int RB, int Ho, int Wo, int C, int cpad) {
  const long long total = rows * cpad;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const int n = threadIdx.x % cpad;   // blockDim.x % cpad == 0
  const int cls = n / C, c = n - cls * C, py = cls >> 1, px = cls & 1;
  const bool live = n < 4 * C;
  float acc = 0.f;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    const long long row = idx / cpad;
    const int b = (int)(row % RB);
    const long long t = row / RB;
    const int a = (int)(t % RA);
    const long long fr = t / RA;
    const int y = 2 * a + py, x = 2 * b + px;
    float v = 0.f;
    if (live && y < Ho && x < Wo)
      v = g_nchw ? g[((fr * C + c) * Ho + y) * Wo + x] : g[((fr * Ho + y) * Wo + x) * C + c];
    G[idx] = v;
    acc += v;
  }
  __shared__ float red[256];
  red[threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x < cpad && live) {
    float v = 0.f;
    for (int i = threadIdx.x; i < 256; i += cpad) v += red[i];
    atomicAdd(db + c, v);
  }
}

// The same pass four columns per thread (cpad is a multiple of 16): one index decomposition per quad instead of a 64-bit
// division per element, 16-byte stores, and 16-byte loads where a quad stays inside one sub-pixel class (NHWC gradient with
// C % 4 == 0).  0.75 -> see DESIGN 4d for the three decoder layers of a 2,450-frame update.
__global__ void __launch_bounds__(256) grad_unshuffle4_kernel(const float* __restrict__ g, int g_nchw, float* __restrict__ G,
                                                              float* __restrict__ db /* C, zeroed */, int rows /* F*RA*RB */,
                                                              int RA, int RB, int Ho, int Wo, int C, int cpad) {
  __shared__ float sdb[64];
  const int Q = cpad >> 2, R = 256 / Q;          // quads per row, rows per pass
  const int q = threadIdx.x % Q, rr = threadIdx.x / Q;
  const int n0 = 4 * q;
  const bool vec = !g_nchw && (C & 3) == 0;
  if (threadIdx.x < 64) sdb[threadIdx.x] = 0.f;
  __syncthreads();
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int row = blockIdx.x * R + rr; row < rows; row += gridDim.x * R) {
    const int b = row % RB, t = row / RB, a = t % RA, fr = t / RA;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n0 < 4 * C) {
      if (vec) {
        const int cls = n0 / C, c = n0 - cls * C, y = 2 * a + (cls >> 1), x = 2 * b + (cls & 1);
        if (y < Ho && x < Wo) v = __ldg(reinterpret_cast<const float4*>(g + (((size_t)fr * Ho + y) * Wo + x) * C + c));
      } else {
        float el[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int n = n0 + e, cls = n / C, c = n - cls * C, y = 2 * a + (cls >> 1), x = 2 * b + (cls & 1);
          el[e] = 0.f;
          if (n < 4 * C && y < Ho && x < Wo)
            el[e] = g_nchw ? __ldg(g + (((size_t)fr * C + c) * Ho + y) * Wo + x) : __ldg(g + (((size_t)fr * Ho + y) * Wo + x) * C + c);
        }
        v = make_float4(el[0], el[1], el[2], el[3]);
      }
    }
    *reinterpret_cast<float4*>(G + (size_t)row * cpad + n0) = v;
    acc[0] += v.x; acc[1] += v.y; acc[2] += v.z; acc[3] += v.w;
  }
#pragma unroll
  for (int e = 0; e < 4; ++e)
    if (n0 + e < 4 * C) atomicAdd(&sdb[(n0 + e) % C], acc[e]);
  __syncthreads();
  if (threadIdx.x < C) atomicAdd(db + threadIdx.x, sdb[threadIdx.x]);
}

}  // namespace rb
