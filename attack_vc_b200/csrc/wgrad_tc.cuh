// wgrad_tc.cuh -- convolution weight gradients on the tensor cores: TMA tensor loads + tcgen05.mma, accumulators in TMEM.
//
//   dW[tap][ci][co] = sum over base pixels p of  A[srcA(p, tap)][ci] * G[srcG(p, tap)][co]
//
// Both operands are activation-shaped NHWC tensors (a 1-D time-major tensor is H = 1) whose contraction axis -- the
// pixel -- is the SLOW axis and whose channels are contiguous: exactly the "MN-major" operand form of tcgen05
// (instruction-descriptor bits 15/16).  So nothing is transposed or gathered by threads:
//   * one elected thread issues `cp.async.bulk.tensor.4d` (TMA, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) for boxes of
//     {32 channels x bw x bh x bb pixels}; conv stride / the 2x up-sampling of a transposed conv are the tensor map's
//     elementStrides, the tap is an offset of the box origin, rows and images outside the tensor are zero-filled by the
//     TMA unit.  The box lands as [pixel][128 B] rows = the canonical MN-major SWIZZLE_128B_BASE32B operand (4-row atoms
//     of 512 B along K, channel blocks LBO apart), the only MN-major form 32-bit operands have;
//   * one elected thread issues UTCHMMA kind::tf32, M = 128 (c_in) x N <= 128 (c_out) x K = 8 pixels;
//   * four warps read the finished tile from TMEM and store this CTA's partial.
// Reflect padding cannot be expressed by a tensor map, so the padded operand is materialised by the same pass that
// splits the operands for precision (next paragraph) -- one elementwise kernel per operand.
//
// Precision: 3xTF32.  a = a_hi + a_lo with a_hi = tf32(a) (round to nearest, stored with a zero low mantissa so the
// tensor core's truncation is the identity) and a_lo = a - a_hi (exact in fp32; its own truncation to tf32 costs 2^-23);
// D += a_hi*g_hi + a_lo*g_hi + a_hi*g_lo, the dropped a_lo*g_lo term is 2^-24.  Every operand plane is fed by TMA, so the
// kernel has no loader warps at all.  TMEM accumulates with truncation (~ -8e-8 relative per accumulate, scripts/tc_probe.py);
// the pixel axis is split over CTAs so that one accumulator sees at most a few hundred MMAs, and the partials are summed
// in fp64 in a fixed order.  Stated tolerance: 2e-5 relative per tensor against an fp64 evaluation (tests/test_kernels_gpu.py).
#pragma once
#include <cuda.h>

#include <algorithm>

#include "common.cuh"
#include "host_util.h"
#include "tc_common.cuh"

namespace avc {

// A pipeline stage holds KP pixels of every operand plane: A_hi | A_lo | G_hi | G_lo, each [32-channel block][KP rows][128 B].
// KP follows the tile: ~64 KB per stage, so a 32 x 64 channel tile (3 blocks) gets 80 pixels per stage and a 128 x 128 tile
// 32 -- with a fixed 32 the small-channel layers, whose pixel axis is the longest, ran on 3 KB boxes and TMA latency.
constexpr int kWtRingBytes = 192 * 1024;
constexpr int kWtStageTarget = 64 * 1024;
constexpr int kWtMaxStages = 6;
constexpr int kWtThreads = 192;                  // warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 epilogue
constexpr int kWtMaxTaps = 9;
inline size_t wt_smem_bytes() { return (size_t)kWtRingBytes + 1024 + 128; }

struct WtArgs {
  int nw, nh, nb;                        // boxes along w, h, b of the BASE pixel grid
  int bw, bh, bb;                        // base pixels per box along each axis (bw*bh*bb <= KP)
  int KP, n_stages, mbm, nbm;            // rows per stage (multiple of 8), ring depth, 32-channel blocks of a full tile (c_in, c_out)
  // A CTA can compute `tg` consecutive taps (own TMEM accumulator each) from ONE pass over the pixels: the operand whose box
  // does not move with the tap (the gradient of a Conv2d / Conv1d, the input of a transposed conv) is then loaded once per stage
  // instead of once per tap.  Measured slower than one tap per CTA (see wt_pick_boxes): tg = 1 by default.
  int tg, a_same, g_same;                // taps per CTA (<= 3); 1: that operand's box offsets are identical for every tap
  int a_wmul, a_hmul, g_wmul, g_hmul;    // tensor coordinate of base pixel (w, h): w * wmul + woff[tap], h * hmul + hoff[tap]
  int a_woff[kWtMaxTaps], a_hoff[kWtMaxTaps], g_woff[kWtMaxTaps], g_hoff[kWtMaxTaps];
  int n_taps;
  int Ci, Co, Cop;
  float* partial;                        // [gridDim.x splits][n_taps][Ci][Cop]
};

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
               ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar) : "memory");
}
// MN-major TF32 operand.  32-bit MN-major operands have exactly one legal shared-memory form: SWIZZLE_128B_BASE32B (layout
// type 1; 32-byte chunks of a 128-byte row XOR-ed with the row index mod 4 -- what CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B
// writes): atoms of 4 rows (K) x 128 B, K groups SBO = 512 B apart (consecutive rows), 32-channel blocks `lbo` bytes apart.
__device__ __forceinline__ uint64_t wt_desc(uint32_t saddr, uint32_t lbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)(512 >> 4) << 32) |
         (1ull << 46) | (1ull << 61);
}

static __global__ void __launch_bounds__(kWtThreads, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
                const __grid_constant__ CUtensorMap tmGh, const __grid_constant__ CUtensorMap tmGl, const WtArgs p) {
  extern __shared__ unsigned char wt_smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(wt_smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)kWtRingBytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);
  const uint32_t bar0 = smem_u32(bars);
  const int n_stages = p.n_stages;
  const uint32_t blk = (uint32_t)p.KP * 128;                                    // one 32-channel block of a plane
  // stage = [A copy 0: hi | lo] .. [A copy na-1] [G copy 0: hi | lo] .. [G copy ng-1]
  const int na = p.a_same ? 1 : p.tg, ng = p.g_same ? 1 : p.tg;
  const uint32_t a_copy = 2 * (uint32_t)p.mbm * blk, g_copy = 2 * (uint32_t)p.nbm * blk;
  const uint32_t offAl = (uint32_t)p.mbm * blk, offGl = (uint32_t)p.nbm * blk, offG = (uint32_t)na * a_copy;
  const uint32_t stage_bytes = (uint32_t)na * a_copy + (uint32_t)ng * g_copy;
  auto full = [&](int s) { return bar0 + 8 * s; };
  auto empty = [&](int s) { return bar0 + 8 * (kWtMaxStages + s); };
  const uint32_t acc_full = bar0 + 8 * (2 * kWtMaxStages);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int split = blockIdx.x, tap0 = (int)blockIdx.z * p.tg;
  const int ntap = min(p.tg, p.n_taps - tap0);                                  // taps of this CTA
  const int n_nt = (p.Co + 127) / 128;
  const int ci0 = ((int)blockIdx.y / n_nt) * 128, co0 = ((int)blockIdx.y % n_nt) * 128;
  const int mb = min(4, (p.Ci - ci0 + 31) / 32), nbk = min(4, (p.Co - co0 + 31) / 32);   // 32-channel blocks that exist
  const int N = nbk * 32;
  const int n_boxes = p.nw * p.nh * p.nb;
  const int per = (n_boxes + (int)gridDim.x - 1) / (int)gridDim.x;
  const int q_lo = split * per, q_hi = min(n_boxes, q_lo + per);
  const int rows = p.bw * p.bh * p.bb, ksteps = (rows + 7) >> 3;

  // unloaded channel blocks and the K tail of every stage must read as zeros
  for (int i = threadIdx.x; i < (int)(n_stages * stage_bytes / 16); i += kWtThreads) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (threadIdx.x == 0) {
    for (int s = 0; s < n_stages; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
    mbar_init(acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  fence_proxy_async();
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (elect_one()) {
      const int na_l = p.a_same ? 1 : ntap, ng_l = p.g_same ? 1 : ntap;         // copies actually loaded
      const uint32_t bytes = (uint32_t)((na_l * mb + ng_l * nbk) * 2 * rows * 128);
      int s = 0; uint32_t ph = 0;
      for (int q = q_lo; q < q_hi; ++q) {
        const int wi = q % p.nw, t = q / p.nw;
        const int hi = t % p.nh, bi = t / p.nh;
        const int w0 = wi * p.bw, h0 = hi * p.bh, b0 = bi * p.bb;
        mbar_wait(empty(s), ph ^ 1);
        mbar_expect_tx(full(s), bytes);
        const uint32_t base = smem_u32(smem) + (uint32_t)s * stage_bytes;
        for (int c = 0; c < na_l; ++c) {
          const int aw = w0 * p.a_wmul + p.a_woff[tap0 + c], ah = h0 * p.a_hmul + p.a_hoff[tap0 + c];
          const uint32_t cb = base + (uint32_t)c * a_copy;
          for (int j = 0; j < mb; ++j) {
            tma_load_4d(cb + (uint32_t)j * blk, &tmAh, ci0 + 32 * j, aw, ah, b0, full(s));
            tma_load_4d(cb + offAl + (uint32_t)j * blk, &tmAl, ci0 + 32 * j, aw, ah, b0, full(s));
          }
        }
        for (int c = 0; c < ng_l; ++c) {
          const int gw = w0 * p.g_wmul + p.g_woff[tap0 + c], gh = h0 * p.g_hmul + p.g_hoff[tap0 + c];
          const uint32_t cb = base + offG + (uint32_t)c * g_copy;
          for (int j = 0; j < nbk; ++j) {
            tma_load_4d(cb + (uint32_t)j * blk, &tmGh, co0 + 32 * j, gw, gh, b0, full(s));
            tma_load_4d(cb + offGl + (uint32_t)j * blk, &tmGl, co0 + 32 * j, gw, gh, b0, full(s));
          }
        }
        if (++s == n_stages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: D[ci][co] += A^T G, both operands MN-major =====
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    const bool leader = elect_one();
    int s = 0; uint32_t ph = 0; uint32_t acc = 0;
    for (int q = q_lo; q < q_hi; ++q) {
      mbar_wait(full(s), ph);
      tc_fence_after();
      if (leader) {
        const uint32_t base = smem_u32(smem) + (uint32_t)s * stage_bytes;
        for (int c = 0; c < ntap; ++c) {
          const uint32_t ab = base + (p.a_same ? 0u : (uint32_t)c * a_copy), gb = base + offG + (p.g_same ? 0u : (uint32_t)c * g_copy);
          const uint32_t d_tmem = tmem_base + (uint32_t)(c * 128);
          // the four descriptors once per tap; a k-step (8 pixel rows = 1024 B) is +64 in the start-address field (shared-memory
          // addresses stay below 2^18: no carry out of the 14 bits) -- the issuing thread is the kernel's clock (conv2d_tc.cuh)
          const uint64_t dAh = wt_desc(ab, blk), dAl = wt_desc(ab + offAl, blk), dGh = wt_desc(gb, blk), dGl = wt_desc(gb + offGl, blk);
          for (int ks = 0; ks < ksteps; ++ks) {
            const uint64_t k = (uint64_t)ks * 64;
            tc_mma_tf32(d_tmem, dAh + k, dGh + k, idesc, (acc || ks) ? 1u : 0u);
            tc_mma_tf32(d_tmem, dAl + k, dGh + k, idesc, 1);
            tc_mma_tf32(d_tmem, dAh + k, dGl + k, idesc, 1);
          }
        }
        acc = 1;
        tc_commit(empty(s));
      }
      __syncwarp();
      if (++s == n_stages) { s = 0; ph ^= 1; }
    }
    if (leader && q_lo < q_hi) tc_commit(acc_full);
    __syncwarp();
  } else {
    // ===== epilogue: TMEM lane = c_in row, columns = c_out =====
    const int quarter = warp & 3;
    const bool any = q_lo < q_hi;
    if (any) { mbar_wait(acc_full, 0); tc_fence_after(); }
    const int ci = ci0 + quarter * 32 + lane;
    for (int c = 0; c < ntap; ++c) {
      float* out = p.partial + (((size_t)split * p.n_taps + tap0 + c) * p.Ci + ci) * p.Cop + co0;
      for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t r[16];
        if (any) {
          tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(c * 128 + c0), r);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) r[i] = 0u;
        }
        if (ci < p.Ci) {
#pragma unroll
          for (int i = 0; i < 16; i += 4)
            if (co0 + c0 + i < p.Cop)
              st4(out + c0 + i, make_float4(__uint_as_float(r[i]), __uint_as_float(r[i + 1]), __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3])));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---- operand preparation: hi / lo planes, optionally reflect-padded -------------------------------------------------------
// The bf16 PAIR plane (conv2d_tc.cuh, C2Args.pair): the two cross terms of the split product, a_lo*b_hi + a_hi*b_lo, as ONE
// kind::f16 MMA with K = 16 per 8 channels.  Per group of 8 channels the plane holds 16 bf16 in the 32 bytes the 8 floats of
// an fp32 plane occupy: an ACTIVATION stores [lo(c0..c7) | hi(c0..c7)], a WEIGHT [hi(c0..c7) | lo(c0..c7)], so that the
// K = 16 dot product pairs lo with hi and hi with lo.  Same tensor shape, same TMA boxes, same swizzle as the fp32 planes.
// A thread owns 4 channels (float index e of the plane, e % 4 == 0, C % 8 == 0): two 8-byte stores.
__device__ __forceinline__ void st_pair4(float* pair, long long e, float4 hi, float4 lo, bool weight_order) {
  const long long u = (e >> 2) & 1;                  // first or second half of the 8-channel group
  float* first = pair + e - 2 * u;                   // group base + 2u floats (8 bytes per 4 bf16)
  float* second = pair + e + 4 - 2 * u;
  st_bf16x4(first, weight_order ? hi : lo);
  st_bf16x4(second, weight_order ? lo : hi);
}
// dst_hi/dst_lo [B][H+2ph][W+pl+pr][C] from src [B][H][W][C]; reflect padding (edge not repeated), C % 4 == 0;
// pair (optional, C % 8 == 0): the bf16 pair plane, in weight order when pair_w
static __global__ void __launch_bounds__(256) wt_split_pad_kernel(const float* __restrict__ src, float* __restrict__ hi, float* __restrict__ lo,
                                                           int B, int H, int W, int C, int ph, int pl, int pr,
                                                           float* __restrict__ pair = nullptr, int pair_w = 0) {
  const int C4 = C >> 2, Wp = W + pl + pr, Hp = H + 2 * ph;
  const long long n4 = (long long)B * Hp * Wp * C4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4) << 2;
    long long t = i / C4;
    const int wp = (int)(t % Wp); t /= Wp;
    const int hp = (int)(t % Hp), b = (int)(t / Hp);
    int w = wp - pl, h = hp - ph;
    w = w < 0 ? -w : w; if (w >= W) w = 2 * (W - 1) - w;
    h = h < 0 ? -h : h; if (h >= H) h = 2 * (H - 1) - h;
    const float4 v = ld4(src + (((long long)b * H + h) * W + w) * C + c);
    const float4 a = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
    const float4 l = f4sub(v, a);
    st4(hi + i * 4, a);
    if (lo) st4(lo + i * 4, l);
    if (pair) st_pair4(pair, i * 4, a, l, pair_w != 0);
  }
}

// hi / lo planes of g * act'(mask): the operand of a dgrad whose input passes backwards through an activation first
static __global__ void __launch_bounds__(256) wt_split_mask_kernel(const float* __restrict__ g, const float* __restrict__ mask, float slope,
                                                                   float* __restrict__ hi, float* __restrict__ lo, long long n4) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 v = ld4(g + i * 4);
    const float4 m = ld4(mask + i * 4);
    v.x *= m.x > 0.f ? 1.f : slope; v.y *= m.y > 0.f ? 1.f : slope; v.z *= m.z > 0.f ? 1.f : slope; v.w *= m.w > 0.f ? 1.f : slope;
    const float4 a = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
    st4(hi + i * 4, a);
    st4(lo + i * 4, f4sub(v, a));
  }
}

// ---- host side ----------------------------------------------------------------------------------------------------------
typedef CUresult (*WtEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                               const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline WtEncodeFn wt_encode_fn() {
  static WtEncodeFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !f)
      fail(AVC_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    return reinterpret_cast<WtEncodeFn>(f);
  }();
  return fn;
}

// One operand: an NHWC fp32 tensor [B][H][W][C] (hi and lo planes of identical shape) read in boxes of
// {32 channels, bw, bh, bb} with element strides (es_w, es_h) along w / h.
struct WtOperand { const float* hi; const float* lo; int C, W, H, B; int es_w, es_h; };

inline CUtensorMap wt_tensor_map(const float* base, const WtOperand& o, int bw, int bh, int bb) {
  if (o.C % 4) fail(AVC_ERR_INVALID, "tensor-core wgrad: channel count %d is not a multiple of 4", o.C);
  if (bw * o.es_w > 256 || bh * o.es_h > 256 || bb > 256) fail(AVC_ERR_INVALID, "tensor-core wgrad: box too large");
  CUtensorMap tm;
  const cuuint64_t dims[4] = {(cuuint64_t)o.C, (cuuint64_t)o.W, (cuuint64_t)o.H, (cuuint64_t)o.B};
  const cuuint64_t strides[3] = {(cuuint64_t)o.C * 4, (cuuint64_t)o.W * o.C * 4, (cuuint64_t)o.H * o.W * o.C * 4};
  const cuuint32_t box[4] = {32u, (cuuint32_t)(bw * o.es_w), (cuuint32_t)(bh * o.es_h), (cuuint32_t)bb};
  const cuuint32_t es[4] = {1u, (cuuint32_t)o.es_w, (cuuint32_t)o.es_h, 1u};
  const CUresult r = wt_encode_fn()(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), dims, strides, box, es,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) fail(AVC_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for a [%d,%d,%d,%d] tensor, box {32,%d,%d,%d}", (int)r, o.B, o.H, o.W, o.C, bw, bh, bb);
  return tm;
}

// Chooses the stage geometry for c_in x c_out channels and the box shape for a base pixel grid [Bn][Hb][Wb]: whole rows
// when they fit, then several rows, then several images.  (Set p.Ci / p.Co first.)
inline void wt_pick_boxes(WtArgs& p, int Wb, int Hb, int Bn) {
  p.mbm = std::min(4, (p.Ci + 31) / 32); p.nbm = std::min(4, (p.Co + 31) / 32);
  // tap grouping: possible when one operand's box does not move with the tap; three taps per CTA (one kernel row of a 3x3 conv)
  p.a_same = p.g_same = 1;
  for (int t = 1; t < p.n_taps; ++t) {
    if (p.a_woff[t] != p.a_woff[0] || p.a_hoff[t] != p.a_hoff[0]) p.a_same = 0;
    if (p.g_woff[t] != p.g_woff[0] || p.g_hoff[t] != p.g_hoff[0]) p.g_same = 0;
  }
  static const bool no_group = getenv("AVC_WT_NO_TAP_GROUP") != nullptr;
  // Measured (PredictiveModel step 256 windows, Conv1d 128 -> 128 k5 at 64 x 512): grouping is parity-green and never faster --
  // everywhere 5.76 -> 6.01 ms per step and 139 -> 180 us, on the small-channel layers only (c_in + c_out <= 96) 5.73 -> 5.77 ms:
  // a third of the CTAs each issuing three times the MMAs loses more than the saved operand reads gain.  Off unless
  // AVC_WT_TAP_GROUP=1 (kept for A/B runs and for shapes with more taps per operand byte).
  static const bool group = getenv("AVC_WT_TAP_GROUP") && atoi(getenv("AVC_WT_TAP_GROUP")) == 1;
  p.tg = (group && !no_group && p.n_taps >= 2 && (p.a_same || p.g_same)) ? std::min(3, p.n_taps) : 1;
  const int na = p.a_same ? 1 : p.tg, ng = p.g_same ? 1 : p.tg;
  const int per_row = 2 * (na * p.mbm + ng * p.nbm) * 128;
  const int cap = std::max(8, std::min(128, kWtStageTarget / per_row / 8 * 8));       // pixels a stage can hold
  p.bw = std::min(Wb, cap);
  p.bh = p.bw == Wb ? std::max(1, std::min(Hb, cap / p.bw)) : 1;
  p.bb = (p.bw == Wb && p.bh == Hb) ? std::max(1, std::min(Bn, cap / (p.bw * p.bh))) : 1;
  p.nw = (Wb + p.bw - 1) / p.bw; p.nh = (Hb + p.bh - 1) / p.bh; p.nb = (Bn + p.bb - 1) / p.bb;
  p.KP = (p.bw * p.bh * p.bb + 7) / 8 * 8;                                             // no larger than the boxes need
  p.n_stages = std::max(2, std::min(kWtMaxStages, kWtRingBytes / (per_row * p.KP)));
}

// The kernels of this header are `static __global__`: every translation unit that includes it owns a copy.  The functions
// that name a kernel are therefore `static` too (an `inline` function is merged across translation units by the linker: the
// surviving definition would launch ITS unit's copy while another unit's call had raised the shared-memory limit of a
// different copy -- "invalid argument" at launch, depending on which definitions the linker kept), and the launch functions
// raise the limit themselves, once per device and translation unit.
static void wt_init_attributes() {
  static bool done[64] = {};
  int dev = 0;
  CK(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && done[dev]) return;
  CK(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wt_smem_bytes()));
  if (dev >= 0 && dev < 64) done[dev] = true;
}

// partial must hold wt_splits(...) * n_taps * Ci * Cop floats; returns the number of splits used
inline int wt_splits(const WtArgs& p, int sm_count) {
  const int tiles = ((p.Ci + 127) / 128) * ((p.Co + 127) / 128) * ((p.n_taps + p.tg - 1) / p.tg);
  const int n_boxes = p.nw * p.nh * p.nb;
  int S = std::max(1, std::min(n_boxes, (2 * sm_count + tiles - 1) / tiles));
  S = std::max(S, (n_boxes * (p.KP / 8) * 3 + 767) / 768);   // at most 768 MMAs into one TMEM accumulator (truncating adds)
  S = std::min(S, n_boxes);
  const int per = (n_boxes + S - 1) / S;
  return (n_boxes + per - 1) / per;
}

static void launch_wgrad_tc(const WtOperand& A, const WtOperand& G, const WtArgs& p, int S, cudaStream_t st) {
  wt_init_attributes();
  const CUtensorMap tAh = wt_tensor_map(A.hi, A, p.bw, p.bh, p.bb), tAl = wt_tensor_map(A.lo, A, p.bw, p.bh, p.bb);
  const CUtensorMap tGh = wt_tensor_map(G.hi, G, p.bw, p.bh, p.bb), tGl = wt_tensor_map(G.lo, G, p.bw, p.bh, p.bb);
  dim3 grid((unsigned)S, (unsigned)(((p.Ci + 127) / 128) * ((p.Co + 127) / 128)), (unsigned)((p.n_taps + p.tg - 1) / p.tg));
  wgrad_tc_kernel<<<grid, kWtThreads, wt_smem_bytes(), st>>>(tAh, tAl, tGh, tGl, p);
  CK(cudaGetLastError());
}

}  // namespace avc
