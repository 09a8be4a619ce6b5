// Implicit-GEMM convolution for the pixel encoder / decoder (reference: algorithms/repo/models/encoder.py:21-41,
// decoder.py:28-48), cuDNN-free.  One persistent CTA per SM walks (row tile, feature tile) work items:
//
//   rows   = output positions (frame, a, b)            -> M = 128 TMEM lanes
//   K      = (tap_y, tap_x, channel) receptive field    -> gathered on the fly from the input tensor
//   N      = output features (<= 256 per tile)          -> weights are the B operand, streamed through a ring
//
// Warp roles:  warp 0  weight loader  (1-D bulk copies of pre-packed fp16 hi/lo slabs)
//              warp 1  MMA issuer     (tcgen05.mma kind::f16, three products per k16 slab: hi*hi + lo*hi + hi*lo)
//              warps 4-7   epilogue   (TMEM -> bias / ReLU / mask -> global; thread = output position)
//              warps 8-15  gatherers  (global fp32 -> fp16 hi/lo core matrices in the ring stage)
// The accumulator is double-buffered in TMEM (2 x 256 columns), so the epilogue of item i overlaps the
// gather + MMAs of item i+1.
//
// ConvTranspose2d(k, stride 2) runs as ONE stride-1 gather over ceil(k/2)^2 taps that produces all four output
// parity classes at once (features = (py, px, cout), weights zero where a class has no tap); the epilogue
// "pixel-shuffles" them onto the 2x finer output grid (ConvMap.shuffle).
#pragma once
#include "ptx.cuh"
#include "rows.cuh"
#include "vm.cuh"

namespace rb {

constexpr int kCvThreads = 768;    // warps 0-3 loader / MMA / spare, 4-7 epilogue, 8-23 gatherers
constexpr int kCvProducers = 512;
constexpr int kCvEpi = 128;
constexpr int kCvMaxStages = 4;
constexpr int kCvAccCols = 256;

struct ConvParams {
  const float* x;          // input activations
  const uint8_t* wblob;    // n_tiles x k16 slabs of NP*64 bytes: [hi: 2 k-groups x NP x 16 B][lo: same]
  const float* bias;       // n_tiles*NP floats (padded with zeros)
  const float* relu_mask;  // optional, indexed like `out`: out = mask > 0 ? y : 0 (backward through a ReLU)
  // optional device floats [s_x, s_w, 1/(s_x*s_w)]: powers of two that lift both operands into fp16's normal range
  // before the hi/lo split (gradients are far below 6e-5, where fp16 has no mantissa left); undone in the epilogue
  const float* scales;
  float* out;
  // Split-activation format ("HL"): a tensor stored as two fp16 planes, hi then lo (x = hi + lo), each laid out like
  // the fp32 tensor would be.  The producing epilogue splits once; gathers of HL inputs are pure 16-byte copies.
  int in_hl, out_hl, mask_hl;         // x / out / relu_mask are HL (pointers address the hi plane)
  long long in_lo_bytes, out_lo_elems, mask_lo_elems;  // distance from the hi plane to the lo plane
  uint32_t a_lbo;                     // byte stride between k-groups of the gathered operand inside a ring stage
  // dense-layer options (plain GEMM maps only): ELU instead of ReLU, the mask interpreted as a layer OUTPUT whose
  // activation derivative multiplies the result (relu' = [h > 0], elu' = h > 0 ? 1 : h + 1), row strides so that out /
  // mask can be column windows of wider row-major matrices (the activation stash)
  int act_elu, mask_elu;
  int out_ld, mask_ld;
  // plain fp32 GEMM maps only: rows are (step, row-in-step) with out_tile_rows rows per step, and the output is stored per
  // 128-row tile of a step as [chunk of 16 features][quarter][128 rows][4 floats] — the layout rssm_rows_kernel reads its
  // posterior addend in (32 lanes x 16 contiguous bytes per load; here 32 rows x 16 contiguous bytes per store)
  int out_tile_rows;
  long long* dbg;   // profiling only: CTA 0 writes cycle sums [mma wait_acc, wait_full, issue | gather g0 loop, wait_empty | epi wait, work | chunks, items]
  ConvMap cm;
  int n_rows, K, k16;      // GEMM rows, real K, k16 slabs
  int n_total, NP, n_tiles;  // output features, features per tile (multiple of 16, <= 256), tiles
  int cout;                // channels per output pixel (n_total, or n_total/4 when cm.shuffle)
  int kc16, n_stages, stage_bytes;  // ring geometry: k16 slabs per stage, stages, bytes per stage
  // biases staged in shared memory by the prologue (n_tiles * NP floats behind the tap table): the epilogue's per-group bias
  // read was a global load on the critical path of every 16-column group — 9 of the 16 us a one-tile launch takes
  int bias_smem;
};

// A operand (gathered rows) inside a ring stage: k-group g (8 fp16 of every row) starts at g * kCvALbo; the 32-byte
// pad over 128 rows x 16 B rotates the banks so the gatherers' 8-byte stores (4 groups x 2 halves x 2 rows per
// half-warp) are conflict-free.
constexpr uint32_t kCvALbo = 2048 + 32;     // fp32 inputs: 8-byte stores, 4 k-groups x 2 halves x 2 rows per half-warp
constexpr uint32_t kCvALboHL = 2048 + 16;   // HL inputs: 16-byte stores, 8 k-groups of one row per quarter-warp
__host__ __device__ inline int conv_stage_bytes(int kc16, int NP) { return kc16 * (4 * (int)kCvALbo + NP * 64); }

// Gather table, built once per CTA: the input address of (row, k) separates into a per-row base plus a per-k
// offset, so the index divisions happen once per k instead of once per (row, k).  NHWC inputs with C % 4 == 0 use
// one entry per channel quad (16-byte loads), anything else one entry per k (scalar loads).
struct ConvRowInfo {   // per GEMM row of the current item: input pointer of its origin pixel and the origin itself
  uint64_t base;
  int ay, bx;
};
struct ConvCol {       // per output feature (single feature tile): offset inside the frame, sub-pixel class, bias
  int off;
  short py, px;
  float bias;
  int pad;
};
constexpr int kCvColEntries = 256;
struct ConvTap {
  int off;         // BYTE offset from the row base
  short tdy, tdx;  // tap displacement in pixels (sentinel -30000 beyond K: fails the bounds test)
};
__host__ __device__ inline bool conv_quads(const ConvMap& cm) { return !cm.in_nchw && (cm.C & 3) == 0 && (cm.pix & 3) == 0; }
// entries: one per k (scalar path), per channel quad (fp32 NHWC), or per channel oct (HL input, C % 8 == 0)
__host__ __device__ inline int conv_table_entries(const ConvMap& cm, int k16, int in_hl = 0) {
  return in_hl ? k16 * 2 : (conv_quads(cm) ? k16 * 4 : k16 * 16);
}
__device__ __forceinline__ void conv_build_table(ConvTap* table, const ConvMap& cm, int k16, int in_hl, int tid, int nthreads) {
  const bool quads = conv_quads(cm);
  const int n_ent = conv_table_entries(cm, k16, in_hl);
  const int kstep = in_hl ? 8 : (quads ? 4 : 1), esz = in_hl ? 2 : 4;
  for (int i = tid; i < n_ent; i += nthreads) {
    const int k = kstep * i;
    const int tap = k / cm.C, ci = k - tap * cm.C;
    ConvTap e{0, -30000, -30000};
    if (tap < cm.ntaps) {
      const int ty = tap / cm.TW, tx = tap - ty * cm.TW;
      e.tdy = (short)(ty * cm.dy);
      e.tdx = (short)(tx * cm.dx);
      e.off = esz * (cm.in_nchw ? (ci * cm.H + e.tdy) * cm.W + e.tdx : (e.tdy * cm.W + e.tdx) * cm.pix + ci);
    }
    table[i] = e;
  }
}

// opaque 64-bit value: keeps a precomputed row pointer in registers instead of letting the compiler re-derive it
// from the kernel parameter on every use
__device__ __forceinline__ uint64_t opaque(uint64_t v) {
  asm volatile("" : "+l"(v));
  return v;
}

template <bool SCALED>
__global__ void __launch_bounds__(kCvThreads, 1) conv_rows_kernel(const __grid_constant__ ConvParams P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* ring = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + (size_t)P.n_stages * P.stage_bytes);
  // bars: [0,4) full, [4,8) empty, [8,10) acc_full, [10,12) acc_empty
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
  ConvRowInfo* row_info = reinterpret_cast<ConvRowInfo*>(bars + 16);   // 4 gather groups x 128 rows
  ConvCol* cols = reinterpret_cast<ConvCol*>(row_info + 4 * 128);      // scalar-store epilogue: one entry per feature
  ConvTap* table = reinterpret_cast<ConvTap*>(cols + kCvColEntries);
  float* bias_s = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(table + conv_table_entries(P.cm, P.k16, P.in_hl)) + 15) & ~uintptr_t(15));
  const float* bias_src = P.bias_smem ? bias_s : P.bias;   // what the epilogue reads

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + 4);
  const uint32_t bar_accf = smem_u32(bars + 8), bar_acce = smem_u32(bars + 10);
  const ConvMap& cm = P.cm;
  const uint64_t t_entry = global_timer_ns();   // (profiling: CTA 0 reports entry / prologue / MMA / epilogue / exit times)

  if (threadIdx.x == 0) {
    for (int i = 0; i < kCvMaxStages; ++i) {
      mbar_init(bar_full + 8 * i, 4 + 1);   // the four warps of the gather group that owns the chunk + the weight loader
      mbar_init(bar_empty + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_accf + 8 * i, 1);
      mbar_init(bar_acce + 8 * i, kCvEpi);
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), 512);
    tmem_relinquish();
  }
  conv_build_table(table, cm, P.k16, P.in_hl, threadIdx.x, kCvThreads);
  if (P.bias_smem)
    for (int n = threadIdx.x; n < P.n_tiles * P.NP; n += kCvThreads) bias_s[n] = P.bias[n];
  const bool have_cols = P.n_tiles == 1 && P.n_total <= kCvColEntries;
  if (have_cols) {
    for (int n = threadIdx.x; n < P.n_total; n += kCvThreads) {
      int py = 0, px = 0, co = n;
      if (cm.shuffle) {
        const int cls = n / P.cout;
        co = n - cls * P.cout; py = cls >> 1; px = cls & 1;
      }
      ConvCol c;
      c.off = cm.out_nchw ? (co * cm.Ho + py) * cm.Wo + px : (py * cm.Wo + px) * P.cout + co;
      c.py = (short)py; c.px = (short)px;
      c.bias = P.bias[n];
      c.pad = co;
      cols[n] = c;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (P.dbg && blockIdx.x == 0 && threadIdx.x == 0) { P.dbg[10] = (long long)t_entry; P.dbg[11] = (long long)global_timer_ns(); }

  const int row_tiles = (P.n_rows + 127) / 128;
  const int n_items = row_tiles * P.n_tiles;
  const uint32_t slab_bytes = (uint32_t)P.NP * 64u;
  const uint32_t a_lbo = P.a_lbo;
  const uint32_t a_half = (uint32_t)P.kc16 * 2u * kCvALbo;  // A hi region (lo follows), then the weight slabs (sized for the larger stride)
  const int n_stages = P.n_stages;

  // register budget per role (setmaxnreg inside each branch, see rows.cuh): launch pool 768*80: 128*40 + 128*88 + 512*88
  if (warp == 0) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    // ================================ weight loader ================================
    uint32_t slot = 0, phase = 0;
    for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
      const int n_tile = w % P.n_tiles;
      const uint8_t* src = P.wblob + (size_t)n_tile * P.k16 * slab_bytes;
      for (int c0 = 0; c0 < P.k16; c0 += P.kc16) {
        const int nsl = min(P.kc16, P.k16 - c0);
        mbar_wait(bar_empty + 8 * slot, phase ^ 1);
        if (elect_one()) {
          const uint32_t bytes = (uint32_t)nsl * slab_bytes;
          mbar_arrive_expect_tx(bar_full + 8 * slot, bytes);
          bulk_g2s(smem_u32(ring + (size_t)slot * P.stage_bytes + 2 * a_half), src + (size_t)c0 * slab_bytes, bytes,
                   bar_full + 8 * slot);
        }
        __syncwarp();
        if (++slot == (uint32_t)n_stages) { slot = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    // ================================ MMA issuer ================================
    uint32_t slot = 0, phase = 0, item = 0;
    long long m_acc = 0, m_full = 0, m_issue = 0, n_chunks = 0;
    const uint32_t idesc = make_idesc_f16(128, P.NP);
    const uint32_t w_lbo = (uint32_t)P.NP * 16u, w_lo = w_lbo * 2u;
    const uint32_t ring_a = smem_u32(ring);
    for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++item) {
      const uint32_t buf = item & 1u, use = item >> 1;
      const long long A0 = clock64();
      mbar_wait(bar_acce + 8 * buf, (use & 1u) ^ 1u);  // epilogue has drained this accumulator
      m_acc += clock64() - A0;
      tc_fence_after();
      const uint32_t d = tmem_base + buf * kCvAccCols;
      uint32_t acc = 0;
      for (int c0 = 0; c0 < P.k16; c0 += P.kc16) {
        const int nsl = min(P.kc16, P.k16 - c0);
        const long long F0 = clock64();
        mbar_wait(bar_full + 8 * slot, phase);
        const long long F1 = clock64();
        m_full += F1 - F0; ++n_chunks;
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = ring_a + slot * (uint32_t)P.stage_bytes;
          uint32_t wa = sa + 2 * a_half;
          for (int j = 0; j < nsl; ++j) {
            const uint64_t a_hi = make_smem_desc(sa + (uint32_t)j * 2u * a_lbo, a_lbo, 128);
            const uint64_t a_lo = make_smem_desc(sa + a_half + (uint32_t)j * 2u * a_lbo, a_lbo, 128);
            const uint64_t b_hi = make_smem_desc(wa, w_lbo, 128);
            const uint64_t b_lo = make_smem_desc(wa + w_lo, w_lbo, 128);
            umma_f16(d, a_hi, b_hi, idesc, (j == 0) ? acc : 1u);
            umma_f16(d, a_lo, b_hi, idesc, 1u);
            umma_f16(d, a_hi, b_lo, idesc, 1u);
            wa += slab_bytes;
          }
          umma_commit(bar_empty + 8 * slot);
        }
        __syncwarp();
        m_issue += clock64() - F1;
        acc = 1u;
        if (++slot == (uint32_t)n_stages) { slot = 0; phase ^= 1; }
      }
      if (elect_one()) umma_commit(bar_accf + 8 * buf);
      __syncwarp();
    }
    if (P.dbg && blockIdx.x == 0 && lane == 0) { P.dbg[0] = m_acc; P.dbg[1] = m_full; P.dbg[2] = m_issue; P.dbg[7] = n_chunks; P.dbg[8] = item; P.dbg[12] = (long long)global_timer_ns(); }
  } else if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");  // spare warps of warpgroup 0: the dec is warpgroup-wide
  } else if (warp >= 8) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 88;");
    // ================================ gatherers ================================
    // Groups of four warps, one per ring slot (n_stages = 2..4 of the four groups are active); group g builds ring
    // chunks g, g+G, g+2G, ... of this CTA's flattened (item, chunk) sequence, so G chunks are in flight at once and
    // every slot has exactly one producer group (a second group on the same slot could run a whole barrier phase ahead
    // of the consumer, which the one-bit phase parity cannot express).  (With all 16 warps on the same chunk every chunk paid one full
    // load round trip plus the whole per-chunk instruction stream before the next one could start.)  Inside a group
    // thread pg owns channel quad / oct q = pg % 8 of rows rs + 16 j (j < 8), rs = pg / 8: a warp instruction reads
    // 4 rows x 128 contiguous bytes.  Per-row base pointer and origin live in a per-group shared-memory block that is
    // rebuilt when the group moves to a new item; per-k offsets come from the tap table.
    const int p = threadIdx.x - 256;
    const int grp = p >> 7, pg = p & 127;
    const int q = pg & 7, rs = pg >> 3;
    ConvRowInfo* info = row_info + grp * 128;
    const bool quads = conv_quads(cm);
    const int per_frame = cm.RA * cm.RB;
    const int chunks_per_item = (P.k16 + P.kc16 - 1) / P.kc16;
    const int my_items = n_items > (int)blockIdx.x ? (n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    auto group_sync = [&] { asm volatile("bar.sync %0, 128;" ::"r"(2 + grp) : "memory"); };
    // this group's position: item index, chunk inside the item; its ring slot is always `grp`
    const int G = n_stages;
    int it = grp < G ? 0 : my_items, ch = grp;
    while (it < my_items && ch >= chunks_per_item) { ch -= chunks_per_item; ++it; }
    const uint32_t slot = (uint32_t)grp;
    uint32_t phase = 0;
    int cur_item = -1;
    const long long lo_delta = P.in_lo_bytes;
    long long g_loop = 0, g_wait = 0;
    const long long G0 = clock64();
    while (it < my_items) {
      if (it != cur_item) {
        group_sync();   // everyone is done reading the previous item's rows
        const int w = blockIdx.x + it * gridDim.x;
        const int row = (w / P.n_tiles) * 128 + pg;
        const int fr = row / per_frame, rem = row - fr * per_frame, a = rem / cm.RB, b = rem - a * cm.RB;
        ConvRowInfo ri;
        ri.ay = row < P.n_rows ? a * cm.sy + cm.y0 : -30000;   // rows past the end fail every bounds test
        ri.bx = b * cm.sx + cm.x0;
        const long long e = cm.in_nchw ? ((long long)fr * cm.C * cm.H + ri.ay) * cm.W + ri.bx
                                       : (((long long)fr * cm.H + ri.ay) * cm.W + ri.bx) * cm.pix;
        ri.base = reinterpret_cast<uint64_t>(P.x) + (row < P.n_rows ? e * (P.in_hl ? 2 : 4) : 0);
        info[pg] = ri;
        group_sync();
        cur_item = it;
      }
      const int c0 = ch * P.kc16;
      const int kchunk = 16 * min(P.kc16, P.k16 - c0);
      uint8_t* stage = ring + (size_t)slot * P.stage_bytes;
      bool waited = false;
      if (P.in_hl) {
        const bool mine = q * 8 < kchunk;
        ConvTap e{0, -30000, -30000};
        if (mine) e = table[(c0 * 16 + q * 8) >> 3];
        const long long ob = e.off;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint4 vh[4], vl[4];
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const ConvRowInfo ri = info[rs + 16 * (half * 4 + jj)];
            const bool ok = (unsigned)(ri.ay + e.tdy) < (unsigned)cm.H && (unsigned)(ri.bx + e.tdx) < (unsigned)cm.W;
            vh[jj] = make_uint4(0, 0, 0, 0);
            vl[jj] = make_uint4(0, 0, 0, 0);
            if (ok) {
              vh[jj] = __ldg(reinterpret_cast<const uint4*>(ri.base + ob));
              vl[jj] = __ldg(reinterpret_cast<const uint4*>(ri.base + ob + lo_delta));
            }
          }
          if (!waited) { const long long W0 = clock64(); mbar_wait(bar_empty + 8 * slot, phase ^ 1); g_wait += clock64() - W0; waited = true; }
          if (mine) {
            uint8_t* dst = stage + (uint32_t)q * a_lbo + (uint32_t)(rs + 64 * half) * 16u;
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              *reinterpret_cast<uint4*>(dst + jj * 256) = vh[jj];
              *reinterpret_cast<uint4*>(dst + a_half + jj * 256) = vl[jj];
            }
          }
        }
      } else {
        const float xs = SCALED ? __ldg(P.scales) : 1.f;
#pragma unroll 1
        for (int hh = 0; hh < 2; ++hh) {
          const int kl = hh * 32 + q * 4;
          if (kl >= kchunk) break;   // uniform per q; the barrier wait below happens in pass 0
          ConvTap e4[4];
          if (quads) e4[0] = table[(c0 * 16 + kl) >> 2];
          else {
#pragma unroll
            for (int i = 0; i < 4; ++i) e4[i] = table[c0 * 16 + kl + i];
          }
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            float4 v[4];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              const ConvRowInfo ri = info[rs + 16 * (half * 4 + jj)];
              if (quads) {
                const bool ok = (unsigned)(ri.ay + e4[0].tdy) < (unsigned)cm.H && (unsigned)(ri.bx + e4[0].tdx) < (unsigned)cm.W;
                v[jj] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (ok) v[jj] = __ldg(reinterpret_cast<const float4*>(ri.base + (long long)e4[0].off));
              } else {
                float el[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const bool ok = (unsigned)(ri.ay + e4[i].tdy) < (unsigned)cm.H && (unsigned)(ri.bx + e4[i].tdx) < (unsigned)cm.W;
                  el[i] = 0.f;
                  if (ok) el[i] = __ldg(reinterpret_cast<const float*>(ri.base + (long long)e4[i].off));
                }
                v[jj] = make_float4(el[0], el[1], el[2], el[3]);
              }
            }
            if (!waited) { const long long W0 = clock64(); mbar_wait(bar_empty + 8 * slot, phase ^ 1); g_wait += clock64() - W0; waited = true; }
            uint8_t* dst = stage + (uint32_t)(hh * 4 + (q >> 1)) * a_lbo + (uint32_t)(q & 1) * 8u + (uint32_t)(rs + 64 * half) * 16u;
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              float4 x4 = v[jj];
              if (SCALED) { x4.x *= xs; x4.y *= xs; x4.z *= xs; x4.w *= xs; }
              uint2 h, l;
              split2_f16(x4.x, x4.y, h.x, l.x);
              split2_f16(x4.z, x4.w, h.y, l.y);
              *reinterpret_cast<uint2*>(dst + jj * 256) = h;
              *reinterpret_cast<uint2*>(dst + a_half + jj * 256) = l;
            }
          }
        }
      }
      if (!waited) mbar_wait(bar_empty + 8 * slot, phase ^ 1);   // a thread without work in this chunk still hands it over
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_full + 8 * slot);
      phase ^= 1;   // G chunks ahead = the same slot, next phase
      ch += G;
      while (ch >= chunks_per_item) { ch -= chunks_per_item; ++it; }
    }
    g_loop = clock64() - G0;
    if (P.dbg && blockIdx.x == 0 && p == 0) { P.dbg[3] = g_loop; P.dbg[4] = g_wait; }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 88;");
    // ================================ epilogue ================================
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
    const int per_frame = cm.RA * cm.RB;
    const int cout = P.cout;
    const float unscale = SCALED ? __ldg(P.scales + 2) : 1.f;
    uint32_t item = 0;
    long long e_wait = 0, e_work = 0;
    for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++item) {
      const uint32_t buf = item & 1u, use = item >> 1;
      const int n_tile = w % P.n_tiles;
      const int row = (w / P.n_tiles) * 128 + r;
      const bool valid = row < P.n_rows;
      const int fr = row / per_frame, rem = row - fr * per_frame, a = rem / cm.RB, b = rem - a * cm.RB;
      const int oy = a * cm.osy + cm.oy0, ox = b * cm.osx + cm.ox0;
      const int n0 = n_tile * P.NP, ncols = min(P.NP, P.n_total - n0);
      // The backward epilogues multiply by the activation derivative of the layer below, read from global memory (`relu_mask`):
      // like the biases it sat on the critical path of every 16-feature group, so the line of group c + 16 is prefetched
      // into L1 while group c is processed (group 0: before the accumulator is even awaited).  No registers involved.
      auto prefetch_mask = [&](int c) {
        // (dense layers only — `mask_ld` set: in the conv data gradients the gather warps need the L1 more than the epilogue
        // does, and the same prefetch made a world-model update 0.15 ms slower)
        if (!valid || !P.relu_mask || !P.mask_ld || c >= ncols) return;
        const int nb = n0 + c;
        int py = 0, px = 0, co = nb;
        if (cm.shuffle) {
          const int cls = nb / cout;
          co = nb - cls * cout; py = cls >> 1; px = cls & 1;
        }
        const int yy = oy + py, xx = ox + px;
        if (yy >= cm.Ho || xx >= cm.Wo || cm.out_nchw) return;
        const size_t od = (((size_t)fr * cm.Ho + yy) * cm.Wo + xx) * cout + co;
        const size_t om = P.mask_ld ? (size_t)fr * P.mask_ld + co : od;
        const char* mp = P.mask_hl ? reinterpret_cast<const char*>(reinterpret_cast<const __half*>(P.relu_mask) + om)
                                   : reinterpret_cast<const char*>(P.relu_mask + om);
        asm volatile("prefetch.global.L1 [%0];" ::"l"(mp));
      };
      prefetch_mask(0);
      const long long E0 = clock64();
      mbar_wait(bar_accf + 8 * buf, use & 1u);
      const long long E1 = clock64();
      e_wait += E1 - E0;
      tc_fence_after();
      for (int c = 0; c < ncols; c += 16) {
        float v[16];
        prefetch_mask(c + 16);
        tmem_ld16(tlane + buf * kCvAccCols + (uint32_t)c, v);
        tmem_ld_wait();
        if (!valid) continue;
        const int nb = n0 + c, cnt = min(16, ncols - c);
        // features that are contiguous in an NHWC / row-major output: vector stores, cnt/4 quads per 16-column group
        const bool vec = !cm.out_nchw && (cnt & 3) == 0 && (cout & 3) == 0 &&
                         ((cnt == 16 && (cout & 15) == 0) || !(cm.shuffle || P.out_hl || P.mask_hl));
        const int nq = cnt >> 2;
        if (vec) {
          int py = 0, px = 0, co = nb;
          if (cm.shuffle) {
            const int cls = nb / cout;
            co = nb - cls * cout; py = cls >> 1; px = cls & 1;
          }
          const int yy = oy + py, xx = ox + px;
          if (yy < cm.Ho && xx < cm.Wo) {
            const size_t od = (((size_t)fr * cm.Ho + yy) * cm.Wo + xx) * cout + co;
            const size_t o = P.out_ld ? (size_t)fr * P.out_ld + co : od;          // plain GEMM into a column window
            const size_t om = P.mask_ld ? (size_t)fr * P.mask_ld + co : od;
            const float4* bp = reinterpret_cast<const float4*>(bias_src + nb);
            float y[16];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 bb = bp[j];
              y[4 * j] = fmaf(v[4 * j], unscale, bb.x); y[4 * j + 1] = fmaf(v[4 * j + 1], unscale, bb.y);
              y[4 * j + 2] = fmaf(v[4 * j + 2], unscale, bb.z); y[4 * j + 3] = fmaf(v[4 * j + 3], unscale, bb.w);
            }
            if (P.act_elu) {
#pragma unroll
              for (int i = 0; i < 16; ++i) y[i] = act_t<ACT_ELU>(y[i]);
            } else if (cm.relu) {
#pragma unroll
              for (int i = 0; i < 16; ++i) y[i] = fmaxf(y[i], 0.f);
            }
            if (P.relu_mask) {
              if (P.mask_hl) {  // sign of the hi plane = sign of the activation
                const uint4* mp = reinterpret_cast<const uint4*>(reinterpret_cast<const __half*>(P.relu_mask) + om);
                const uint4 m0 = __ldg(mp), m1 = __ldg(mp + 1);
                const uint32_t mw[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&mw[i]));
                  y[2 * i] = f.x > 0.f ? y[2 * i] : 0.f;
                  y[2 * i + 1] = f.y > 0.f ? y[2 * i + 1] : 0.f;
                }
              } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  if (j >= nq) break;
                  const float4 m = __ldg(reinterpret_cast<const float4*>(P.relu_mask + om) + j);
                  const float mm[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    const float d = mm[i] > 0.f ? 1.f : (P.mask_elu ? mm[i] + 1.f : 0.f);
                    y[4 * j + i] *= d;
                  }
                }
              }
            }
            if (P.out_hl) {
              uint32_t h[8], l[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                split2_f16(y[2 * i], y[2 * i + 1], h[i], l[i]);
                // the hi plane doubles as the ReLU mask of the backward pass (sign of the activation): a positive value
                // below fp16's smallest subnormal must not vanish, so it is stored as that subnormal (6e-8 absolute)
                if (y[2 * i] > 0.f && (h[i] & 0x7fffu) == 0u) h[i] |= 1u;
                if (y[2 * i + 1] > 0.f && (h[i] & 0x7fff0000u) == 0u) h[i] |= 0x10000u;
              }
              uint4* hp = reinterpret_cast<uint4*>(reinterpret_cast<__half*>(P.out) + o);
              uint4* lp = reinterpret_cast<uint4*>(reinterpret_cast<__half*>(P.out) + P.out_lo_elems + o);
              hp[0] = make_uint4(h[0], h[1], h[2], h[3]); hp[1] = make_uint4(h[4], h[5], h[6], h[7]);
              lp[0] = make_uint4(l[0], l[1], l[2], l[3]); lp[1] = make_uint4(l[4], l[5], l[6], l[7]);
            } else {
              float4* op = reinterpret_cast<float4*>(P.out + o);
              int qstride = 1;   // float4s between the quarters of a 16-feature group
              if (P.out_tile_rows) {
                const int stp = fr / P.out_tile_rows, rr = fr - stp * P.out_tile_rows;
                const size_t tile = (size_t)stp * ((P.out_tile_rows + 127) >> 7) + (rr >> 7);
                op = reinterpret_cast<float4*>(P.out + (tile * ((P.n_total + 15) >> 4) + (nb >> 4)) * 2048u) + (rr & 127);
                qstride = 128;
              }
#pragma unroll
              for (int j = 0; j < 4; ++j)
                if (j < nq) op[j * qstride] = make_float4(y[4 * j], y[4 * j + 1], y[4 * j + 2], y[4 * j + 3]);
            }
          }
        } else if (have_cols) {
          // scalar stores with the per-feature table: one shared-memory read replaces the class / offset arithmetic
          const long long rb_d = cm.out_nchw ? ((long long)fr * cout * cm.Ho + oy) * cm.Wo + ox
                                             : (((long long)fr * cm.Ho + oy) * cm.Wo + ox) * cout;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            if (i < cnt) {
              const ConvCol cc = cols[nb + i];
              if (oy + cc.py < cm.Ho && ox + cc.px < cm.Wo) {
                const size_t od = (size_t)(rb_d + cc.off);
                const size_t o = P.out_ld ? (size_t)fr * P.out_ld + cc.pad : od;
                float y = fmaf(v[i], unscale, cc.bias);
                if (P.act_elu) y = act_t<ACT_ELU>(y);
                else if (cm.relu) y = fmaxf(y, 0.f);
                if (P.relu_mask) {
                  const float m = __ldg(P.relu_mask + (P.mask_ld ? (size_t)fr * P.mask_ld + cc.pad : od));
                  y *= m > 0.f ? 1.f : (P.mask_elu ? m + 1.f : 0.f);
                }
                P.out[o] = y;
              }
            }
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            if (i < cnt) {
              const int n = nb + i;
              int py = 0, px = 0, co = n;
              if (cm.shuffle) {
                const int cls = n / cout;
                co = n - cls * cout; py = cls >> 1; px = cls & 1;
              }
              const int yy = oy + py, xx = ox + px;
              if (yy < cm.Ho && xx < cm.Wo) {
                const size_t od = cm.out_nchw ? (((size_t)fr * cout + co) * cm.Ho + yy) * cm.Wo + xx
                                              : (((size_t)fr * cm.Ho + yy) * cm.Wo + xx) * cout + co;
                const size_t o = P.out_ld ? (size_t)fr * P.out_ld + co : od;
                float y = fmaf(v[i], unscale, bias_src[n]);
                if (P.act_elu) y = act_t<ACT_ELU>(y);
                else if (cm.relu) y = fmaxf(y, 0.f);
                if (P.relu_mask) {
                  const float m = __ldg(P.relu_mask + (P.mask_ld ? (size_t)fr * P.mask_ld + co : od));
                  y *= m > 0.f ? 1.f : (P.mask_elu ? m + 1.f : 0.f);
                }
                P.out[o] = y;
              }
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(bar_acce + 8 * buf);
    }
    if (P.dbg && blockIdx.x == 0 && r == 0) P.dbg[13] = (long long)global_timer_ns();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
  if (P.dbg && blockIdx.x == 0 && threadIdx.x == 0) P.dbg[14] = (long long)global_timer_ns();
}


// =================================================================================================
// Weight gradient of the same implicit GEMM:  dW[n, k] = sum_rows G[row, n] * X[row, k]
// (G = output gradient rows, X = the gathered receptive fields, never materialised).
//
// The contraction runs over ROWS, so both operands are "MN-major" for tcgen05 (the M / N index is the
// contiguous one in memory): A = X^T with M = k (tiles of 128), B = G^T with N = n (<= 256), 16 rows per MMA.
// smem core matrix = 8 rows x 16 bytes (8 consecutive k, or n, of one row); SBO = stride between 8-k groups
// (padded to 144 B so the gather threads' 8-byte stores rotate over all banks), LBO = stride between 8-row
// groups.  One CTA owns a "super tile" of mt k-tiles x all n in TMEM (mt * NP <= 512 columns) and a slice of
// the rows; slices are summed with fp32 atomics into the zero-initialised dW.
// =================================================================================================
constexpr int kWgRows = 32;           // rows per ring stage (two MMA K-slabs)
constexpr uint32_t kWgSbo = 144;
constexpr int kWgMaxSuper = 16;

struct WgradParams {
  const float* x;        // layer input (gather source)
  const float* g;        // output-gradient rows (n_rows, g_ld)
  float* dw;             // (n_total, K) fp32, zero-initialised
  const float* scales;   // optional device floats [s_x, s_g, 1/(s_x*s_g)]
  long long* dbg;        // profiling only: CTA 0 writes cycle sums [issue, wait_empty, data+store, arrive | mma_wait_full, mma_issue, stages]
  int x_hl;              // x is in split-activation format (fp16 hi plane at x, lo plane x_lo_bytes further)
  long long x_lo_bytes;
  ConvMap cm;
  int n_rows, K, k16, n_total, g_ld, NP;
  // blockIdx.y = slice of the output features: slice s takes columns [s * NP, s * NP + n) of g and rows [s * NP, ..) of dw,
  // n = NP except n_last in the last slice (dW of a wide Linear layer in ONE launch instead of one per 256 features)
  int n_last;
  int n_super;                       // super tiles
  int m0[kWgMaxSuper], mt[kWgMaxSuper];   // first 128-k tile and tile count of each super tile
  int cta0[kWgMaxSuper + 1];         // first CTA of each super tile (prefix sums of the row splits)
  int kt_max;                        // 128 * max(mt): k extent the stage is laid out for
  int n_stages, stage_bytes;
};

__host__ __device__ inline int wgrad_stage_bytes(int kt, int NP) { return (int)kWgSbo * (kt + NP); }

__global__ void __launch_bounds__(kCvThreads, 1) conv_wgrad_kernel(const __grid_constant__ WgradParams P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* ring = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + (size_t)P.n_stages * P.stage_bytes);
  // bars: [0,4) full, [4,8) empty, [8] acc_full
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
  ConvTap* table = reinterpret_cast<ConvTap*>(bars + 14);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + 4), bar_accf = smem_u32(bars + 8);
  const ConvMap& cm = P.cm;
  const int n_slice = (blockIdx.y + 1 == gridDim.y) ? P.n_last : P.n_total;   // output features of this CTA's slice

  if (threadIdx.x == 0) {
    for (int i = 0; i < kCvMaxStages; ++i) {
      mbar_init(bar_full + 8 * i, kCvProducers / 32);   // one arrival per gather warp
      mbar_init(bar_empty + 8 * i, 1);
    }
    mbar_init(bar_accf, 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), 512);
    tmem_relinquish();
  }
  const bool quads = conv_quads(cm);
  conv_build_table(table, cm, P.k16, P.x_hl, threadIdx.x, kCvThreads);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // which super tile / row slice is this CTA
  int s = 0;
  while (s + 1 < P.n_super && (int)blockIdx.x >= P.cta0[s + 1]) ++s;
  const int splits = P.cta0[s + 1] - P.cta0[s], split = blockIdx.x - P.cta0[s];
  const int m0 = P.m0[s], mt = P.mt[s];
  const int stages_total = (P.n_rows + kWgRows - 1) / kWgRows;
  const int per = (stages_total + splits - 1) / splits;
  const int st0 = split * per, st1 = min(stages_total, st0 + per);
  const int kt = mt * 128;                                   // k extent of this CTA
  const uint32_t lbo_a = (uint32_t)(P.kt_max / 8) * kWgSbo, lbo_b = (uint32_t)(P.NP / 8) * kWgSbo;
  const uint32_t a_half = 4u * lbo_a, b_half = 4u * lbo_b;   // hi region bytes (lo follows)
  const int n_stages = P.n_stages;

  if (warp == 1) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    // ================================ MMA issuer ================================
    uint32_t slot = 0, phase = 0;
    const uint32_t idesc = make_idesc_f16(128, P.NP) | (1u << 15) | (1u << 16);   // A and B MN-major
    const uint32_t ring_a = smem_u32(ring);
    uint32_t first = 1;
    long long m_wait = 0, m_issue = 0;
    for (int st = st0; st < st1; ++st) {
      const long long M0 = clock64();
      mbar_wait(bar_full + 8 * slot, phase);
      const long long M1 = clock64();
      m_wait += M1 - M0;
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sa = ring_a + slot * (uint32_t)P.stage_bytes;
        const uint32_t sb = sa + 2 * a_half;
        for (int slab = 0; slab < kWgRows / 16; ++slab) {
          const uint64_t b_hi = make_smem_desc(sb + (uint32_t)slab * 2u * lbo_b, lbo_b, kWgSbo);
          const uint64_t b_lo = make_smem_desc(sb + b_half + (uint32_t)slab * 2u * lbo_b, lbo_b, kWgSbo);
          for (int t = 0; t < mt; ++t) {
            const uint32_t ao = (uint32_t)slab * 2u * lbo_a + (uint32_t)t * 16u * kWgSbo;
            const uint64_t a_hi = make_smem_desc(sa + ao, lbo_a, kWgSbo);
            const uint64_t a_lo = make_smem_desc(sa + a_half + ao, lbo_a, kWgSbo);
            const uint32_t d = tmem_base + (uint32_t)(t * P.NP);
            umma_f16(d, a_hi, b_hi, idesc, (first && slab == 0) ? 0u : 1u);
            umma_f16(d, a_lo, b_hi, idesc, 1u);
            umma_f16(d, a_hi, b_lo, idesc, 1u);
          }
        }
        umma_commit(bar_empty + 8 * slot);
      }
      __syncwarp();
      m_issue += clock64() - M1;
      first = 0;
      if (++slot == (uint32_t)n_stages) { slot = 0; phase ^= 1; }
    }
    if (P.dbg && blockIdx.x == 0 && lane == 0) { P.dbg[4] = m_wait; P.dbg[5] = m_issue; }
    if (elect_one()) umma_commit(bar_accf);
    __syncwarp();
  } else if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
  } else if (warp >= 8) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 88;");
    // ================================ gatherers ================================
    // thread p owns quad q = p % 16 of every 64-wide k (or n) block of row rs = p / 16 of the stage
    const int p = threadIdx.x - 256;
    const int q = p & 15, rs = p >> 4;
    uint32_t slot = 0, phase = 0;
    const int per_frame = cm.RA * cm.RB;
    const float xs = P.scales ? __ldg(P.scales) : 1.f, gs = P.scales ? __ldg(P.scales + 1) : 1.f;
    const uint32_t row_off = (uint32_t)(rs & 7) * 16u + (uint32_t)(q & 1) * 8u + (uint32_t)(q >> 1) * kWgSbo;
    const uint32_t a_off = (uint32_t)(rs >> 3) * lbo_a + row_off, b_off = (uint32_t)(rs >> 3) * lbo_b + row_off;
    const int kblocks = kt / 64, nblocks = (P.NP + 63) / 64;
    long long c_issue = 0, c_wait = 0, c_store = 0, c_arrive = 0;
    // (frame, a, b) of this thread's row: divided out once, then advanced by kWgRows per stage (the gather warps are
    // instruction-issue bound: ~6 warps per scheduler, so two integer divisions per stage are a visible cost)
    int fr, ra, rb;
    {
      const int row = st0 * kWgRows + rs;
      fr = row / per_frame;
      const int rem = row - fr * per_frame;
      ra = rem / cm.RB;
      rb = rem - ra * cm.RB;
    }
    for (int st = st0; st < st1; ++st) {
      const long long T0 = clock64();
      const int row = st * kWgRows + rs;
      const bool valid = row < P.n_rows;
      const int a = ra, b = rb;
      const int ay = valid ? a * cm.sy + cm.y0 : -30000, bx = b * cm.sx + cm.x0;
      const long long e0 = cm.in_nchw ? ((long long)fr * cm.C * cm.H + ay) * cm.W + bx
                                      : (((long long)fr * cm.H + ay) * cm.W + bx) * cm.pix;
      const uint64_t base = opaque(reinterpret_cast<uint64_t>(P.x + (valid ? e0 : 0)));
      uint8_t* sa = ring + (size_t)slot * P.stage_bytes;
      uint8_t* sb = sa + 2 * a_half;
      // every global load of the stage (gathered rows AND gradient rows) is issued before the ring slot is awaited:
      // one L2 round trip per stage, overlapped with the wait for the MMAs that still read the slot
      float4 gv[4];
      {
        const float* grow = P.g + (size_t)blockIdx.y * P.NP + (size_t)(valid ? row : 0) * P.g_ld;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int n = i * 64 + q * 4;
          gv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (i < nblocks && valid && n < n_slice) gv[i] = __ldg(reinterpret_cast<const float4*>(grow + n));
        }
      }
      if (P.x_hl) {
        // HL input: 16 lanes x one channel oct each = 128 k per pass, pure 16-byte copies of both planes
        const uint64_t bh = opaque(reinterpret_cast<uint64_t>(P.x) + (valid ? e0 * 2 : 0));
        uint4 vh[4], vl[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          vh[i] = make_uint4(0, 0, 0, 0);
          vl[i] = make_uint4(0, 0, 0, 0);
          const int k = m0 * 128 + i * 128 + q * 8;
          if (i * 128 < kt && k < P.k16 * 16) {
            const ConvTap e = table[k >> 3];
            const bool ok = (unsigned)(ay + e.tdy) < (unsigned)cm.H && (unsigned)(bx + e.tdx) < (unsigned)cm.W;
            if (ok) {
              vh[i] = __ldg(reinterpret_cast<const uint4*>(bh + (long long)e.off));
              vl[i] = __ldg(reinterpret_cast<const uint4*>(bh + (long long)e.off + P.x_lo_bytes));
            }
          }
        }
        const long long T1 = clock64();
        mbar_wait(bar_empty + 8 * slot, phase ^ 1);
        const long long T2 = clock64();
        c_issue += T1 - T0; c_wait += T2 - T1; c_store -= T2;
        const uint32_t o0 = (uint32_t)(rs >> 3) * lbo_a + (uint32_t)(rs & 7) * 16u + (uint32_t)q * kWgSbo;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (i * 128 < kt) {
            const uint32_t o = o0 + (uint32_t)i * 16u * kWgSbo;
            *reinterpret_cast<uint4*>(sa + o) = vh[i];
            *reinterpret_cast<uint4*>(sa + a_half + o) = vl[i];
          }
        }
      } else {
        bool waited = false;
        for (int kb0 = 0; kb0 < kblocks; kb0 += 4) {
          float4 v[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int kb = kb0 + i;
            v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (kb < kblocks) {
              const int k = m0 * 128 + kb * 64 + q * 4;
              if (k < P.k16 * 16) {
                if (quads) {
                  const ConvTap e = table[k >> 2];
                  const bool ok = (unsigned)(ay + e.tdy) < (unsigned)cm.H && (unsigned)(bx + e.tdx) < (unsigned)cm.W;
                  if (ok) v[i] = __ldg(reinterpret_cast<const float4*>(base + (long long)e.off));
                } else {
                  float el[4];
#pragma unroll
                  for (int c = 0; c < 4; ++c) {
                    const ConvTap e = table[k + c];
                    const bool ok = (unsigned)(ay + e.tdy) < (unsigned)cm.H && (unsigned)(bx + e.tdx) < (unsigned)cm.W;
                    el[c] = 0.f;
                    if (ok) el[c] = __ldg(reinterpret_cast<const float*>(base + (long long)e.off));
                  }
                  v[i] = make_float4(el[0], el[1], el[2], el[3]);
                }
              }
            }
          }
          if (!waited) {
            const long long T1 = clock64();
            mbar_wait(bar_empty + 8 * slot, phase ^ 1);
            const long long T2 = clock64();
            c_issue += T1 - T0; c_wait += T2 - T1; c_store -= T2;
            waited = true;
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int kb = kb0 + i;
            if (kb < kblocks) {
              uint2 h, l;
              split2_f16(v[i].x * xs, v[i].y * xs, h.x, l.x);
              split2_f16(v[i].z * xs, v[i].w * xs, h.y, l.y);
              const uint32_t o = a_off + (uint32_t)kb * 8u * kWgSbo;
              *reinterpret_cast<uint2*>(sa + o) = h;
              *reinterpret_cast<uint2*>(sa + a_half + o) = l;
            }
          }
        }
      }
      // output-gradient rows
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int n = i * 64 + q * 4;
        if (i < nblocks && n < P.NP) {
          uint2 h, l;
          split2_f16(gv[i].x * gs, gv[i].y * gs, h.x, l.x);
          split2_f16(gv[i].z * gs, gv[i].w * gs, h.y, l.y);
          const uint32_t o = b_off + (uint32_t)i * 8u * kWgSbo;
          *reinterpret_cast<uint2*>(sb + o) = h;
          *reinterpret_cast<uint2*>(sb + b_half + o) = l;
        }
      }
      const long long T3 = clock64();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_full + 8 * slot);
      const long long T4 = clock64();
      c_store += T3; c_arrive += T4 - T3;
      if (++slot == (uint32_t)n_stages) { slot = 0; phase ^= 1; }
      rb += kWgRows;
      while (rb >= cm.RB) { rb -= cm.RB; ++ra; }
      while (ra >= cm.RA) { ra -= cm.RA; ++fr; }
    }
    if (P.dbg && blockIdx.x == 0 && p == 0) {
      P.dbg[0] = c_issue; P.dbg[1] = c_wait; P.dbg[2] = c_store; P.dbg[3] = c_arrive; P.dbg[6] = st1 - st0;
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 88;");
    // ================================ epilogue ================================
    if (st1 > st0) {
      const int qd = warp & 3;
      const uint32_t tlane = tmem_base + ((uint32_t)(qd * 32) << 16);
      const float unscale = P.scales ? __ldg(P.scales + 2) : 1.f;
      mbar_wait(bar_accf, 0);
      tc_fence_after();
      for (int t = 0; t < mt; ++t) {
        const int k = (m0 + t) * 128 + qd * 32 + lane;
        float* dw = P.dw + (size_t)blockIdx.y * P.NP * P.K;
        for (int c = 0; c < n_slice; c += 16) {
          float v[16];
          tmem_ld16(tlane + (uint32_t)(t * P.NP + c), v);
          tmem_ld_wait();
          if (k < P.K) {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (c + i < n_slice) atomicAdd(dw + (size_t)(c + i) * P.K + k, v[i] * unscale);
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace rb
