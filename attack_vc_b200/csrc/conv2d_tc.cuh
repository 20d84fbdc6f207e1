// conv2d_tc.cuh -- the PredictiveModel's 3x3 convolutions (forward, transposed forward, both dgrads) as TMA-fed tcgen05
// implicit GEMMs:   Y[pixel][c_out] = sum over (tap, c_in block)  X[src(pixel, tap)][c_in] * W[tap][c_out][c_in]
//
//   M = 128 output pixels of one TMA box {32 channels x bw x bh x bb} over the NHWC activation tensor: the conv stride is the
//       tensor map's element stride, the tap an offset of the box origin, everything outside the tensor is zero-filled by
//       the TMA unit (the zero padding of the transposed forms; ReflectionPad2d(1) is materialised by the pass that splits
//       the activations for precision, wgrad_tc.cuh);
//   N <= 128 output channels, K = 32 input channels per pipeline stage (4 UMMA k-steps of 8), both operands K-major
//       SWIZZLE_128B exactly as cp.async.bulk.tensor writes them -- no thread touches an operand;
//   transposed convolutions (ConvTranspose2d forward, Conv2d dgrad) run one launch per output residue class
//       (oh mod s, ow mod s): a class meets only its own taps, so nothing is multiplied through structural zeros.
// Precision: 3xTF32 (hi / lo planes of activations and weights), or with C2Args.pair the hi planes plus bf16 pair planes
// (two MMAs per k-step).  The tensor core accumulates fp32 with TRUNCATION (about
// -8e-8 relative per accumulate, scripts/tc_probe.py): left alone over the 864 MMAs of a 256-channel 3x3 conv that is a
// systematic 4e-5, enough to move PReLU units across their kink and the gradients by 5e-3.  So accumulation is chunked as
// in conv_tc.cuh: one pipeline stage (12 MMAs) per TMEM buffer, two buffers; the epilogue warps add every finished chunk
// into fp32 registers with round-to-nearest while the next chunk is being multiplied.
// Warp roles: 0 TMA producer, 1 MMA issuer (+ TMEM allocation), 2-9 epilogue (bias, folded BatchNorm, act' mask,
// activation, tanh; one pixel row per lane, two warps per TMEM lane quarter with 64 columns each -- with four epilogue warps
// of 128 columns the chunk adds and the row stores of a tile took longer than its MMAs).
#pragma once
#include "wgrad_tc.cuh"

namespace avc {

constexpr int kC2Stages = 3;
constexpr int kC2Threads = 320;                  // warp 0 TMA producer, warp 1 MMA issuer, warps 2-9 epilogue (two per TMEM lane quarter, 64 columns each)
constexpr int kC2Plane = 128 * 128;              // one operand plane of a stage: 128 rows x 128 B
constexpr int kC2StageBytes = 4 * kC2Plane;      // X_hi | X_lo | W_hi | W_lo
constexpr int kC2ScrPitch = 20;                  // floats per row of an epilogue warp's 32 x 16 transpose slab (16 + 4: float4-aligned rows)
constexpr int kC2ScrBytes = 8 * 32 * kC2ScrPitch * 4;
inline size_t c2_smem_bytes() { return (size_t)kC2Stages * kC2StageBytes + kC2ScrBytes + 1024 + 128; }

struct C2Args {
  // base pixel grid of this launch (one residue class) and its boxes
  int Hb, Wb, B;
  int bw, bh, bb, nw, nh, nb;
  int a_wmul, a_hmul;                    // tensor coordinate of base pixel (w, h) for tap t: w * a_wmul + a_woff[t]
  int n_taps;
  int tap[9], a_woff[9], a_hoff[9];      // weight tap index and box offsets
  int Ci, Cop;                           // contraction channels; output channels incl. padding (multiple of 4)
  int box_n;                             // rows of the weight box: min(128, Cop rounded up to 32) -- a 32-channel output loads 32 weight rows, not 128
  // output: y[b][(h * oh_mul + oh_off)][(w * ow_mul + ow_off)][Co]
  float* y; int Ho, Wo, Co;
  int oh_mul, oh_off, ow_mul, ow_off;
  const float* bias; const float* scale; const float* shift;   // v = (acc + bias) * scale + shift   (each may be null)
  const float* dmask; float mslope;      // v *= dmask > 0 ? 1 : mslope   (tensor shaped like y)
  const float* slope_ptr; float slope;   // act: v > 0 ? v : v * slope   (slope_ptr overrides: PReLU on the device)
  int act;                               // 0 none, 1 leaky, 2 leaky then tanh
  // K split for launches with fewer tiles than SMs (the 2x1 .. 5x4-pixel layers: 16 tiles of 144 dependent stages each):
  // ksplit CTAs share a tile, each takes a contiguous range of the (tap, K block) stages and stores its raw sums to
  // part + ks * part_stride (indexed like y); c2_finish_kernel adds them in order and applies the epilogue.
  int ksplit; float* part; long long part_stride;
  // pair = 1: the "lo" tensor maps address bf16 PAIR planes (wgrad_tc.cuh st_pair4) and a k-step is TWO MMAs,
  // kind::tf32 X_hi W_hi + kind::f16 (K = 16) [x_lo | x_hi] . [w_hi | w_lo], instead of three TF32 MMAs.  At M = N = 128 every
  // MMA reads 8 KB of operands from shared memory (64 clk at 128 B/clk, 87 measured next to the TMA writes): the MMA count is
  // the kernel's clock.  Error of the bf16 cross terms: 2^-9 of a 2^-11 term, unbiased (the same split conv_tc.cuh uses).
  int pair;
};

// K-major SWIZZLE_128B operand: rows of 128 B (32 tf32), 8-row groups SBO = 1024 B apart
__device__ __forceinline__ uint64_t c2_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
               ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}

// role-level clock stamps of CTA 0 (build with -DAVC_C2_PROFILE): where each role of the pipeline waits
#ifdef AVC_C2_PROFILE
#define C2P(...) __VA_ARGS__
#else
#define C2P(...)
#endif

static __global__ void __launch_bounds__(kC2Threads, 1)
conv2d_tc_kernel(const __grid_constant__ CUtensorMap tmXh, const __grid_constant__ CUtensorMap tmXl,
                 const __grid_constant__ CUtensorMap tmWh, const __grid_constant__ CUtensorMap tmWl, const C2Args p) {
  extern __shared__ unsigned char c2_smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(c2_smem_raw) + 1023) & ~(uintptr_t)1023);
  float* scratch = reinterpret_cast<float*>(smem + (size_t)kC2Stages * kC2StageBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)kC2Stages * kC2StageBytes + kC2ScrBytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);      // after the 10 barriers
  const uint32_t bar0 = smem_u32(bars);
  auto full = [&](int s) { return bar0 + 8 * s; };
  auto empty = [&](int s) { return bar0 + 8 * (kC2Stages + s); };
  const uint32_t acc_full0 = bar0 + 8 * (2 * kC2Stages), acc_empty0 = acc_full0 + 16;    // two TMEM buffers
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // persistent: this CTA walks work items (pixel box, 128-channel output tile); the stage ring and the accumulator ring run
  // on across items, so the next item's operands are in flight while this one's epilogue stores (a launch of many small
  // tiles used to be bound by per-CTA start-up: barrier init, TMEM allocation, a cold pipeline)
  const int n_nt = (p.Cop + 127) / 128;
  const int ksplit = p.ksplit;
  const int n_work = p.nw * p.nh * p.nb * n_nt * ksplit;
  const int rows = p.bw * p.bh * p.bb;
  const int nkb = (p.Ci + 31) >> 5;
  const int n_stage = p.n_taps * nkb;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kC2Stages; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(acc_full0 + 8 * b, 1); mbar_init(acc_empty0 + 8 * b, 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (elect_one()) {
      const uint32_t bytes = (uint32_t)(2 * (rows + p.box_n) * 128);   // weight rows past c_out inside the box are zero-filled by TMA
      int s = 0; uint32_t ph = 0;
      C2P(long long pw = 0, pt0 = clock64(), pq;)
      for (int wk = blockIdx.x; wk < n_work; wk += gridDim.x) {
        const int ks = wk % ksplit, wq = wk / ksplit;
        const int q = wq / n_nt, n0 = (wq % n_nt) * 128;
        const int wi = q % p.nw, tq = q / p.nw;
        const int w0 = wi * p.bw, h0 = (tq % p.nh) * p.bh, b0 = (tq / p.nh) * p.bb;
        const int it_lo = ks * n_stage / ksplit, it_hi = (ks + 1) * n_stage / ksplit;
        for (int it = it_lo; it < it_hi; ++it) {
          const int t = it / nkb, kb = it - t * nkb;
          const int aw = w0 * p.a_wmul + p.a_woff[t], ah = h0 * p.a_hmul + p.a_hoff[t], wt = p.tap[t];
          C2P(pq = clock64();)
          mbar_wait(empty(s), ph ^ 1);
          C2P(pw += clock64() - pq;)
          mbar_expect_tx(full(s), bytes);
          const uint32_t base = smem_u32(smem) + (uint32_t)s * kC2StageBytes;
          tma_load_4d(base, &tmXh, kb * 32, aw, ah, b0, full(s));
          tma_load_4d(base + kC2Plane, &tmXl, kb * 32, aw, ah, b0, full(s));
          tma_load_3d(base + 2 * kC2Plane, &tmWh, kb * 32, n0, wt, full(s));
          tma_load_3d(base + 3 * kC2Plane, &tmWl, kb * 32, n0, wt, full(s));
          if (++s == kC2Stages) { s = 0; ph ^= 1; }
        }
      }
      C2P(if (blockIdx.x == 0) printf("[conv2d_tc cta0] producer: total %lld clk, wait empty %lld (stages/item %d, items %d, N tiles %d)\n", clock64() - pt0, pw, n_stage / ksplit, (n_work + (int)gridDim.x - 1) / (int)gridDim.x, n_nt);)
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    const bool leader = elect_one();
    int s = 0; uint32_t ph = 0;
    uint32_t chunk = 0;                                             // chunks issued so far: buffer chunk & 1
    C2P(long long iw_acc = 0, iw_full = 0, i_iss = 0, it0 = clock64(), iq;)
    for (int wk = blockIdx.x; wk < n_work; wk += gridDim.x) {
      const int ks = wk % ksplit, wq = wk / ksplit;
      const int n0 = (wq % n_nt) * 128;
      const int N = min(128, (p.Cop - n0 + 15) & ~15);              // MMA N: multiple of 16
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t idesc16 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);   // BF16 x BF16 -> fp32, K = 16
      const int it_lo = ks * n_stage / ksplit, it_hi = (ks + 1) * n_stage / ksplit;
      for (int it = it_lo; it < it_hi; ++it, ++chunk) {
        const uint32_t buf = chunk & 1;
        C2P(iq = clock64();)
        mbar_wait(acc_empty0 + 8 * buf, ((chunk >> 1) & 1) ^ 1);    // the drain warps have taken this buffer's previous chunk
        C2P(iw_acc += clock64() - iq; iq = clock64();)
        mbar_wait(full(s), ph);
        C2P(iw_full += clock64() - iq; iq = clock64();)
        tc_fence_after();
        if (leader) {
          const uint32_t d_tmem = tmem_base + buf * 128;
          uint32_t acc = 0;
          const uint32_t base = smem_u32(smem) + (uint32_t)s * kC2StageBytes;
          const int kleft = p.Ci - (it % nkb) * 32;                 // channels of this K block that exist (the rest is zero-filled)
          const int ksteps = min(4, (kleft + 7) >> 3);
          // separate loops: a branch on p.pair inside the k-step loop cost the three-MMA path 9 % (2016 -> 2194 us per
          // PredictiveModel step) -- the single issuing thread is the kernel's clock
          // descriptors of the stage's four planes once; a k-step is +2 in the start-address field (32 B >> 4; shared memory
          // addresses stay below 2^18, so the 14-bit field cannot carry)
          const uint64_t dXh = c2_desc(base), dXl = c2_desc(base + kC2Plane), dWh = c2_desc(base + 2 * kC2Plane), dWl = c2_desc(base + 3 * kC2Plane);
          if (!p.pair) {
            if (ksteps == 4) {
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                tc_mma_tf32(d_tmem, dXh + 2 * ks, dWh + 2 * ks, idesc, ks ? 1u : 0u);
                tc_mma_tf32(d_tmem, dXl + 2 * ks, dWh + 2 * ks, idesc, 1);
                tc_mma_tf32(d_tmem, dXh + 2 * ks, dWl + 2 * ks, idesc, 1);
              }
            } else {
              for (int ks = 0; ks < ksteps; ++ks) {
                tc_mma_tf32(d_tmem, dXh + 2 * ks, dWh + 2 * ks, idesc, acc);
                acc = 1;
                tc_mma_tf32(d_tmem, dXl + 2 * ks, dWh + 2 * ks, idesc, 1);
                tc_mma_tf32(d_tmem, dXh + 2 * ks, dWl + 2 * ks, idesc, 1);
              }
            }
          } else {
            // all TF32 MMAs of the stage, then all BF16 ones (same accumulator, the order is free): a change of MMA kind costs
            // ~64 clk at N = 128 (scripts/ubench/mma_shapes.cu), alternating per k-step ate the whole gain of the shorter stage
            for (int ks = 0; ks < ksteps; ++ks) tc_mma_tf32(d_tmem, dXh + 2 * ks, dWh + 2 * ks, idesc, ks ? 1u : 0u);
            for (int ks = 0; ks < ksteps; ++ks) tc_mma_bf16(d_tmem, dXl + 2 * ks, dWl + 2 * ks, idesc16);   // x_lo w_hi + x_hi w_lo
          }
          tc_commit(empty(s));
          tc_commit(acc_full0 + 8 * buf);
        }
        __syncwarp();
        C2P(i_iss += clock64() - iq;)
        if (++s == kC2Stages) { s = 0; ph ^= 1; }
      }
    }
    C2P(if (blockIdx.x == 0 && lane == 0) printf("[conv2d_tc cta0] issuer: total %lld clk, wait acc_empty %lld, wait full %lld, issue %lld\n", clock64() - it0, iw_acc, iw_full, i_iss);)
  } else {
    // ===== epilogue: TMEM lane = pixel of the box, columns = output channels =====
    const int quarter = warp & 3, half = (warp - 2) >> 2;          // TMEM lanes 32*quarter.., columns 64*half..
    const int cbase = half * 64;
    const int r = quarter * 32 + lane;
    const int w = r % p.bw, hh = (r / p.bw) % p.bh, bi = r / (p.bw * p.bh);
    const float slope = p.slope_ptr ? *p.slope_ptr : p.slope;
    uint32_t chunk = 0;
    C2P(long long ew = 0, ed = 0, es = 0, et0 = clock64(), eq;)
    for (int wk = blockIdx.x; wk < n_work; wk += gridDim.x) {
      const int ks = wk % ksplit, wq = wk / ksplit;
      const int q = wq / n_nt, n0 = (wq % n_nt) * 128;
      const int N = min(128, (p.Cop - n0 + 15) & ~15);
      const int it_lo = ks * n_stage / ksplit, it_hi = (ks + 1) * n_stage / ksplit;
      const int wi = q % p.nw, tq = q / p.nw;
      const int w0 = wi * p.bw, h0 = (tq % p.nh) * p.bh, b0 = (tq / p.nh) * p.bb;
      const bool ok = r < rows && w0 + w < p.Wb && h0 + hh < p.Hb && b0 + bi < p.B;
      const long long o = ok ? ((((long long)(b0 + bi) * p.Ho + (h0 + hh) * p.oh_mul + p.oh_off) * p.Wo + (w0 + w) * p.ow_mul + p.ow_off) * p.Co + n0) : 0;
      float acc[64];
#pragma unroll
      for (int i = 0; i < 64; ++i) acc[i] = 0.f;
#pragma unroll 1
      for (int it = it_lo; it < it_hi; ++it, ++chunk) {
        const uint32_t buf = chunk & 1;
        C2P(eq = clock64();)
        mbar_wait(acc_full0 + 8 * buf, (chunk >> 1) & 1);
        C2P(ew += clock64() - eq; eq = clock64();)
        tc_fence_after();
        const uint32_t t0 = tmem_base + ((uint32_t)(quarter * 32) << 16) + buf * 128 + (uint32_t)cbase;
#pragma unroll
        for (int c0 = 0; c0 < 64; c0 += 32) {
          if (cbase + c0 < N) {
            uint32_t v0[16], v1[16];
            tmem_ld16(t0 + c0, v0);
            tmem_ld16(t0 + c0 + 16, v1);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) { acc[c0 + i] += __uint_as_float(v0[i]); acc[c0 + 16 + i] += __uint_as_float(v1[i]); }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty0 + 8 * buf);
        C2P(ed += clock64() - eq;)
      }
      C2P(eq = clock64();)
      // Output through a per-warp 32 x 16 transpose slab: a lane owns a pixel ROW of the tile, so storing from registers made
      // every warp store hit 32 different rows with 16 bytes each (6.5 k clk per tile, more than its MMAs --
      // profiles/r02i_conv2d_tc_roles.txt).  After the transpose a warp store covers 8 rows x 64 contiguous bytes, and the
      // epilogue arithmetic (per-COLUMN constants, the act' mask) is read along channels too.
      {
        float* scr = scratch + (warp - 2) * (32 * kC2ScrPitch);
        const int o32 = ok ? (int)(o - n0) : -1;                     // element offset of this lane's row (fits: tensors < 2^31 elements)
        int orow[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) orow[i] = __shfl_sync(0xffffffffu, o32, (lane >> 2) + 8 * i);
        const bool fin = ksplit == 1;
        float* const dst_base = fin ? p.y : p.part + (long long)ks * p.part_stride;
        const int cq = (lane & 3) * 4;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int col0 = cbase + 16 * g;
          if (col0 < N) {
#pragma unroll
            for (int j = 0; j < 16; j += 4) st4(scr + lane * kC2ScrPitch + j, make_float4(acc[16 * g + j], acc[16 * g + j + 1], acc[16 * g + j + 2], acc[16 * g + j + 3]));
            __syncwarp();
            const int c = n0 + col0 + cq;
            const bool c_ok = col0 + cq < N && c < p.Co;
            float4 bq = f4zero(), sc = make_float4(1.f, 1.f, 1.f, 1.f), sf = f4zero();
            if (fin && c_ok) {
              if (p.bias) bq = ld4(p.bias + c);
              if (p.scale) { sc = ld4(p.scale + c); sf = ld4(p.shift + c); }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              if (orow[i] >= 0 && c_ok) {
                const float4 v = ld4(scr + ((lane >> 2) + 8 * i) * kC2ScrPitch + cq);
                float x[4] = {v.x, v.y, v.z, v.w};
                const long long e = (long long)orow[i] + c;
                if (fin) {
                  x[0] = fmaf(x[0] + bq.x, sc.x, sf.x); x[1] = fmaf(x[1] + bq.y, sc.y, sf.y);
                  x[2] = fmaf(x[2] + bq.z, sc.z, sf.z); x[3] = fmaf(x[3] + bq.w, sc.w, sf.w);
                  if (p.dmask) {
                    const float4 m = ld4(p.dmask + e);
                    x[0] *= m.x > 0.f ? 1.f : p.mslope; x[1] *= m.y > 0.f ? 1.f : p.mslope; x[2] *= m.z > 0.f ? 1.f : p.mslope; x[3] *= m.w > 0.f ? 1.f : p.mslope;
                  }
                  if (p.act) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) x[j] = x[j] > 0.f ? x[j] : x[j] * slope;
                    if (p.act == 2) {
#pragma unroll
                      for (int j = 0; j < 4; ++j) x[j] = tanhf(x[j]);
                    }
                  }
                }
                st4(dst_base + e, make_float4(x[0], x[1], x[2], x[3]));
              }
            }
            __syncwarp();
          }
        }
      }
      C2P(es += clock64() - eq;)
    }
    C2P(if (blockIdx.x == 0 && warp == 2 && lane == 0) printf("[conv2d_tc cta0] epilogue: total %lld clk, wait acc_full %lld, drain %lld, store %lld\n", clock64() - et0, ew, ed, es);)
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
  }
}

// y = epilogue(sum over the K splits, in order) for a K-split launch (every residue class of a layer writes the same buffers)
static __global__ void __launch_bounds__(256) c2_finish_kernel(const C2Args p, long long n4) {
  const float slope = p.slope_ptr ? *p.slope_ptr : p.slope;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)((i * 4) % p.Co);
    float4 a = ld4(p.part + i * 4);
    for (int k = 1; k < p.ksplit; ++k) a = f4add(a, ld4(p.part + (long long)k * p.part_stride + i * 4));
    float x[4] = {a.x, a.y, a.z, a.w};
    if (p.bias) { const float4 bq = ld4(p.bias + c); x[0] += bq.x; x[1] += bq.y; x[2] += bq.z; x[3] += bq.w; }
    if (p.scale) {
      const float4 sc = ld4(p.scale + c), sf = ld4(p.shift + c);
      x[0] = fmaf(x[0], sc.x, sf.x); x[1] = fmaf(x[1], sc.y, sf.y); x[2] = fmaf(x[2], sc.z, sf.z); x[3] = fmaf(x[3], sc.w, sf.w);
    }
    if (p.dmask) {
      const float4 m = ld4(p.dmask + i * 4);
      x[0] *= m.x > 0.f ? 1.f : p.mslope; x[1] *= m.y > 0.f ? 1.f : p.mslope; x[2] *= m.z > 0.f ? 1.f : p.mslope; x[3] *= m.w > 0.f ? 1.f : p.mslope;
    }
    if (p.act) {
#pragma unroll
      for (int j = 0; j < 4; ++j) x[j] = x[j] > 0.f ? x[j] : x[j] * slope;
      if (p.act == 2) {
#pragma unroll
        for (int j = 0; j < 4; ++j) x[j] = tanhf(x[j]);
      }
    }
    st4(p.y + i * 4, make_float4(x[0], x[1], x[2], x[3]));
  }
}

// ---- host side ----------------------------------------------------------------------------------------------------------
// activation tensor map: [B][H][W][C] NHWC, K-major SWIZZLE_128B boxes {32 channels, bw, bh, bb}
inline CUtensorMap c2_act_map(const float* base, const WtOperand& o, int bw, int bh, int bb) {
  if (o.C % 4) fail(AVC_ERR_INVALID, "tensor-core conv: channel count %d is not a multiple of 4", o.C);
  CUtensorMap tm;
  const cuuint64_t dims[4] = {(cuuint64_t)o.C, (cuuint64_t)o.W, (cuuint64_t)o.H, (cuuint64_t)o.B};
  const cuuint64_t strides[3] = {(cuuint64_t)o.C * 4, (cuuint64_t)o.W * o.C * 4, (cuuint64_t)o.H * o.W * o.C * 4};
  const cuuint32_t box[4] = {32u, (cuuint32_t)(bw * o.es_w), (cuuint32_t)(bh * o.es_h), (cuuint32_t)bb};
  const cuuint32_t es[4] = {1u, (cuuint32_t)o.es_w, (cuuint32_t)o.es_h, 1u};
  const CUresult r = wt_encode_fn()(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), dims, strides, box, es,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) fail(AVC_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for a [%d,%d,%d,%d] activation tensor", (int)r, o.B, o.H, o.W, o.C);
  return tm;
}
// weight tensor map: [9][rows][K] (tap, output channel, contraction channel), boxes {32, 128, 1}
inline CUtensorMap c2_weight_map(const float* base, int K, int rows, int box_n, int taps = 9) {
  CUtensorMap tm;
  const cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)taps};
  const cuuint64_t strides[2] = {(cuuint64_t)K * 4, (cuuint64_t)rows * K * 4};
  const cuuint32_t box[3] = {32u, (cuuint32_t)box_n, 1u};
  const cuuint32_t es[3] = {1u, 1u, 1u};
  const CUresult r = wt_encode_fn()(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, es,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) fail(AVC_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for a [9,%d,%d] weight tensor", (int)r, rows, K);
  return tm;
}

// box shape for a base grid [B][Hb][Wb] with up to 128 pixels per box; rows are split evenly so that the last box of an
// image is not mostly empty
inline void c2_pick_boxes(C2Args& p) {
  p.bw = std::min(p.Wb, 128);
  if (p.bw < p.Wb) { p.bw = (p.Wb + (p.Wb + 127) / 128 - 1) / ((p.Wb + 127) / 128); p.bh = 1; p.bb = 1; }
  else {
    const int maxh = std::max(1, 128 / p.bw);
    if (maxh >= p.Hb) { p.bh = p.Hb; p.bb = std::max(1, std::min(p.B, 128 / (p.bw * p.bh))); }
    else { const int nh = (p.Hb + maxh - 1) / maxh; p.bh = (p.Hb + nh - 1) / nh; p.bb = 1; }
  }
  p.nw = (p.Wb + p.bw - 1) / p.bw; p.nh = (p.Hb + p.bh - 1) / p.bh; p.nb = (p.B + p.bb - 1) / p.bb;
}

// static, not inline: see wt_init_attributes (wgrad_tc.cuh)
static void c2_init_attributes() {
  static bool done[64] = {};
  int dev = 0;
  CK(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && done[dev]) return;
  CK(cudaFuncSetAttribute(conv2d_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c2_smem_bytes()));
  if (dev >= 0 && dev < 64) done[dev] = true;
}

// X: activation planes (es_w / es_h = the gather stride); Wh / Wl: weight planes [9][rows >= Cop][Kp]
static void launch_conv2d_tc(const WtOperand& X, const float* Wh, const float* Wl, int Kp, int w_rows, C2Args p, int sm_count, cudaStream_t st, int w_taps = 9) {
  c2_init_attributes();
  p.box_n = std::min(128, (p.Cop + 31) / 32 * 32);
  const CUtensorMap tXh = c2_act_map(X.hi, X, p.bw, p.bh, p.bb), tXl = c2_act_map(X.lo, X, p.bw, p.bh, p.bb);
  const CUtensorMap tWh = c2_weight_map(Wh, Kp, w_rows, p.box_n, w_taps), tWl = c2_weight_map(Wl, Kp, w_rows, p.box_n, w_taps);
  if (p.ksplit < 1) p.ksplit = 1;
  const int n_work = p.nw * p.nh * p.nb * ((p.Cop + 127) / 128) * p.ksplit;
  conv2d_tc_kernel<<<std::min(n_work, sm_count), kC2Threads, c2_smem_bytes(), st>>>(tXh, tXl, tWh, tWl, p);
  CK(cudaGetLastError());
}

inline int c2_tiles(const C2Args& p) { return p.nw * p.nh * p.nb * ((p.Cop + 127) / 128); }
inline int c2_stages(const C2Args& p) { return p.n_taps * ((p.Ci + 31) / 32); }

static void launch_c2_finish(const C2Args& p, int sm_count, cudaStream_t st) {
  const long long n4 = (long long)p.B * p.Ho * p.Wo * p.Co / 4;
  const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>((n4 + 255) / 256, (long long)sm_count * 8));
  c2_finish_kernel<<<grid, 256, 0, st>>>(p, n4);
  CK(cudaGetLastError());
}

}  // namespace avc
