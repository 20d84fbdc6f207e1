// conv_tc2.cuh -- the tcgen05 implicit-GEMM Conv1d with the operand roles SWAPPED: weights are the M = 128 operand, 256 window
// rows are the N operand, the accumulator in TMEM is the TRANSPOSED output tile D^T[c_out][row].
//
// Why (scripts/ubench/mma_shapes.cu, profiles/r02a_mma_shapes.txt, measured on B200):
//   * one thread issues tcgen05.mma and the tensor pipe keeps only about two instructions in flight: every cycle the issuing
//     thread spends between two MMAs on barrier waits, fences and descriptor arithmetic beyond the length of ONE MMA is a cycle
//     the pipe idles.  With N = 128 an MMA is 64 clk and the per-stage protocol of conv_tc.cuh (two mbarrier waits, a commit)
//     costs 119 clk per MMA in isolation (112-131 in the kernel) against a floor of 64.7; switching kind::tf32 <-> kind::f16
//     costs another ~64 clk at N = 128 and nothing at N = 256.
//   * operand fetch is NOT the limit: N = 128 and N = 256 both run at the 64.7 clk per 128x128x8 floor back to back, with 100
//     B/clk of concurrent STS, 160 B/clk of LDS and bulk copies in the background; 128-byte swizzle and A-from-TMEM change nothing.
//   So: N = 256 MMAs (128 clk each: twice the cover for the same protocol) -- c_out is 128 in this model, hence the swap --
//   and fewer waits per MMA (the weight ring is handed over per 4-MMA plane, the window per K block).
//
// What else the transposed tile buys: a drain lane owns one output CHANNEL and 128 consecutive rows, so the epilogue stores
// straight from registers with lanes along the channel axis (128 B per warp and row, coalesced) -- no shared-memory transpose.
//
// Same numerics as conv_tc.cuh (same split, same packed weight images, same chunked accumulation, same stage table).
//   warp 0      weight producer: one bulk copy (TMA engine, mbarrier complete_tx) per 16 KB plane of a stage, ring of 5
//   warp 1      MMA issuer
//   warps 2-7   window loaders (192 threads): (row, k-step) units of 32 B, row pointers in a shared-memory table
//   warps 8-15  drain (TMEM -> fp32 chunk sums in registers) + epilogue; setmaxnreg gives them 168 registers, the rest 88
#pragma once
#include "conv_tc.cuh"

namespace avc {

constexpr int kT2N = 256;                         // window rows (GEMM N) of one tile
constexpr int kT2Rows = kT2N + kMaxTaps;          // rows of one window stage
constexpr int kT2XStages = 2;
constexpr int kT2WSlots = 5;                      // weight ring, in planes (half stages) of 16 KB
constexpr int kT2Threads = 512;
constexpr int kT2Loaders = 192;
constexpr int kT2XPlane = (kTcKB / 4) * kT2Rows * 4;              // floats of one plane of a window stage
constexpr int kT2WPlaneBytes = (kTcKB / 4) * kTcNMax * 16;        // 16 KB
constexpr int kT2AccBufs = 2;                                     // 2 x 256 TMEM columns

inline size_t tc2_smem_bytes() {
  return (size_t)kT2XStages * 2 * kT2XPlane * 4 + (size_t)kT2WSlots * kT2WPlaneBytes + 32 * 8 + (size_t)2 * kT2Rows * 8 + (size_t)kT2N * (16 + 4);
}

// ---- epilogue ------------------------------------------------------------------------------------
// Row descriptors: everything that depends on the virtual row (utterance, output row, skip rows and their weight) is worked out
// ONCE per tile, one row per drain thread, and kept in shared memory as element offsets; the epilogue reads a row's descriptor
// with one broadcast LDS.128.  (Every lane of every drain warp needs every row's addresses: computing them per lane would cost
// a 64-bit division per row and lane.)
struct T2Desc { int yoff, y2off, g0, g1; };   // yoff: -1 dead row, <= -2: dgrad halo row at side offset -2 - yoff; g0/g1: -1 = none

__device__ __forceinline__ void t2_make_desc(const TcArgs& p, const TcPass& ps, long long u, T2Desc& d, float& rs) {
  const TcRow r = tc_row_info(p, ps, u);
  d.yoff = -1; d.y2off = 0; d.g0 = -1; d.g1 = -1; rs = r.rs;
  if (r.kind == 2) d.yoff = -2 - (int)(((long long)r.b * (p.halo_l + p.halo_r) + r.o) * p.side_n);
  if (r.kind != 1) return;
  d.yoff = (int)((long long)r.b * p.y_bs + (long long)r.o * p.y_rs);
  if (p.Y2) d.y2off = (int)((long long)r.b * p.y2_bs + (long long)r.o * p.y2_rs);
  if (p.Om) d.g0 = (int)((long long)r.b * p.om_bs + (long long)r.o * p.om_rs);
  else if (p.res.mode != RES_NONE) {
    d.g0 = (int)((long long)r.b * p.res.bs + (long long)r.t0 * p.res.rs);
    if (r.t1 >= 0) d.g1 = (int)((long long)r.b * p.res.bs + (long long)r.t1 * p.res.rs);
  }
}

// MODE: 0 plain, 1 act' mask on the output (Om), 2 one skip row (RES_SAME / RES_UP / RES_POOL_BWD), 3 two skip rows (RES_POOL / RES_UP_BWD)
template <int MODE>
__device__ __forceinline__ void t2_epilogue(const TcArgs& p, const float (&acc)[128], const T2Desc* __restrict__ desc,
                                            const float* __restrict__ rsv, int c, bool c_ok) {
  const float bias = (p.bias && c_ok) ? p.bias[c] : 0.f;
  const float* __restrict__ G = MODE == 1 ? p.Om : p.res.R;
  constexpr int NB = 8;                      // rows whose operand loads are in flight before the first is consumed
#pragma unroll
  for (int r0 = 0; r0 < 128; r0 += NB) {
    float ga[NB], gb[NB];
    if (MODE >= 1) {
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        const int4 d = *reinterpret_cast<const int4*>(desc + r0 + j);
        ga[j] = 0.f; gb[j] = 0.f;
        if (c_ok && d.z >= 0) ga[j] = G[d.z + c];
        if (MODE == 3 && c_ok && d.w >= 0) gb[j] = G[d.w + c];
      }
    }
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      const int4 d = *reinterpret_cast<const int4*>(desc + r0 + j);
      if (!c_ok || d.x == -1) continue;
      float x = acc[r0 + j];
      if (d.x < -1) { p.side[(-2 - d.x) + c] = x; continue; }
      x += bias;
      if (MODE == 1) x *= (ga[j] > 0.f ? 1.f : p.slope);
      if (p.act) x = x > 0.f ? x : x * p.slope;
      if (p.Y2) p.Y2[d.y + c] = x;
      if (MODE == 2) x += ga[j] * rsv[r0 + j];
      if (MODE == 3) x += (ga[j] + gb[j]) * rsv[r0 + j];
      p.Y[d.x + c] = x;
    }
  }
}

__global__ void __launch_bounds__(kT2Threads, 1) conv_tc2_kernel(const TcArgs p) {
  extern __shared__ __align__(128) unsigned char t2_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* Xs = reinterpret_cast<float*>(t2_smem);                                              // [2 stages][2 planes][kT2XPlane]
  unsigned char* Ws = t2_smem + (size_t)kT2XStages * 2 * kT2XPlane * 4;                        // [5 slots][16 KB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(Ws + (size_t)kT2WSlots * kT2WPlaneBytes);
  long long* rowtab = reinterpret_cast<long long*>(bars + 32);                                 // [2][kT2Rows]: element offsets of the window rows (A, mask)
  T2Desc* desc = reinterpret_cast<T2Desc*>(rowtab + 2 * kT2Rows);                               // [kT2N] output row descriptors of the current tile
  float* rsv = reinterpret_cast<float*>(desc + kT2N);                                          // [kT2N] skip-row weights
  const uint32_t bar0 = smem_u32(bars);
  auto x_full = [&](int s) { return bar0 + 8 * s; };
  auto x_empty = [&](int s) { return bar0 + 8 * (2 + s); };
  auto w_full = [&](int s) { return bar0 + 8 * (4 + s); };
  auto w_empty = [&](int s) { return bar0 + 8 * (9 + s); };
  const uint32_t acc_full0 = bar0 + 8 * 14, acc_empty0 = bar0 + 8 * 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kT2XStages; ++s) { mbar_init(x_full(s), kT2Loaders / 32); mbar_init(x_empty(s), 1); }
    for (int s = 0; s < kT2WSlots; ++s) { mbar_init(w_full(s), 1); mbar_init(w_empty(s), 1); }
    for (int s = 0; s < kT2AccBufs; ++s) { mbar_init(acc_full0 + 8 * s, 1); mbar_init(acc_empty0 + 8 * s, 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();

  const int n_tiles = (int)((p.Mv + kT2N - 1) / kT2N);
  const int n_work = n_tiles * p.n_pass;
  const int w_first = (int)blockIdx.x, w_step = (int)gridDim.x;

  if (warp < 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 88;");
    if (warp == 0) {
      // ===== weight producer =====
      if (lane == 0) {
        int sw = 0; uint32_t pw = 0;
        for (int wk = w_first; wk < n_work; wk += w_step) {
          const TcPass& ps = p.pass[wk % p.n_pass];
          const uint32_t N = (uint32_t)ps.N;
#pragma unroll 1
          for (int s = ps.s_begin; s < ps.s_end; ++s) {
            const TcStage e = p.st[s];
            const uint32_t bytes = 32u * e.nks * N;                  // ONE plane: 8 nks channels x N x 4 bytes
            const float* src = p.g[e.gi].Wp + e.w_off;
#pragma unroll 1
            for (int u = 0; u < 2; ++u) {
              mbar_wait(w_empty(sw), pw ^ 1);
              mbar_expect_tx(w_full(sw), bytes);
              bulk_g2s(smem_u32(Ws + (size_t)sw * kT2WPlaneBytes), src + (size_t)u * (bytes >> 2), bytes, w_full(sw));
              if (++sw == kT2WSlots) { sw = 0; pw ^= 1; }
            }
          }
        }
      }
    } else if (warp == 1) {
      // ===== MMA issuer: the whole warp walks the table, one elected lane issues =====
      const bool leader = elect_one();
      int sx = 0, sw = 0; uint32_t px = 0, pw = 0;
      int chunk = 0;
      const int terms = p.terms;
      const uint32_t x_base = smem_u32(Xs), w_base = smem_u32(Ws);
      constexpr uint32_t x_lbo = kT2Rows * 16;
      constexpr uint64_t x_ks = (uint64_t)((2 * x_lbo) >> 4), x_lo_off = (uint64_t)((kT2XPlane * 4) >> 4);
      for (int wk = w_first; wk < n_work; wk += w_step) {
        const TcPass& ps = p.pass[wk % p.n_pass];
        const uint32_t N = (uint32_t)ps.N;                                                       // valid output channels (M is always 128)
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kT2N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t idesc16 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kT2N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t w_lbo = N * 16;
        const uint64_t w_ks = (uint64_t)((2 * w_lbo) >> 4);
        uint64_t x_desc0 = 0;
        uint32_t d_tmem = 0, acc = 0;
#pragma unroll 1
        for (int s = ps.s_begin; s < ps.s_end; ++s) {
          const TcStage e = p.st[s];
          if (e.flags & kTcKbFirst) {
            mbar_wait(x_full(sx), px);
            x_desc0 = tc_desc(x_base + (uint32_t)sx * (2 * kT2XPlane * 4), x_lbo, 128);
          }
          if (e.flags & kTcChunkFirst) {
            const int buf = chunk % kT2AccBufs;
            mbar_wait(acc_empty0 + 8 * buf, ((chunk / kT2AccBufs) & 1) ^ 1);
            d_tmem = tmem_base + (uint32_t)(buf * kT2N);
            acc = 0;
          }
          const int nks = e.nks;
          // plane 0: kind::tf32  D^T += W_hi * X_hi^T
          mbar_wait(w_full(sw), pw);
          tc_fence_after();
          {
            uint64_t dw = tc_desc(w_base + (uint32_t)sw * kT2WPlaneBytes, w_lbo, 128);
            uint64_t dx = x_desc0 + (uint64_t)e.tap;
#pragma unroll 4
            for (int ks = 0; ks < nks; ++ks) {
              if (leader) tc_mma_tf32(d_tmem, dw, dx, idesc, acc);
              acc = 1;
              dw += w_ks; dx += x_ks;
            }
            if (leader) tc_commit(w_empty(sw));
            if (++sw == kT2WSlots) { sw = 0; pw ^= 1; }
          }
          // plane 1: kind::f16 (bf16, K = 16)  D^T += [W_hi | W_lo] * [X_lo | X_hi]^T
          mbar_wait(w_full(sw), pw);
          tc_fence_after();
          {
            uint64_t dw = tc_desc(w_base + (uint32_t)sw * kT2WPlaneBytes, w_lbo, 128);
            uint64_t dx = x_desc0 + (uint64_t)e.tap + x_lo_off;
            if (terms >= 2) {
#pragma unroll 4
              for (int ks = 0; ks < nks; ++ks) {
                if (leader) tc_mma_bf16(d_tmem, dw, dx, idesc16);
                dw += w_ks; dx += x_ks;
              }
            }
            if (leader) {
              tc_commit(w_empty(sw));
              if (e.flags & kTcChunkLast) tc_commit(acc_full0 + 8 * (chunk % kT2AccBufs));
              if (e.flags & kTcKbLast) tc_commit(x_empty(sx));
            }
            if (++sw == kT2WSlots) { sw = 0; pw ^= 1; }
          }
          __syncwarp();
          if (e.flags & kTcChunkLast) ++chunk;
          if (e.flags & kTcKbLast) { if (++sx == kT2XStages) { sx = 0; px ^= 1; } }
        }
      }
    } else {
      // ===== window loaders (192 threads) =====
      // unit q of a K block = (row q / 4, k-step q % 4): four consecutive lanes read the 128 contiguous bytes of one row's
      // 32 channels; every thread handles units tl, tl + 192, ... (5.5 on average), three at a time.
      pdl_wait();
      const int tl = threadIdx.x - 64;                    // 0..191
      int sx = 0; uint32_t px = 0;
      for (int wk = w_first; wk < n_work; wk += w_step) {
        const TcPass& ps = p.pass[wk % p.n_pass];
        const long long v0 = (long long)(wk / p.n_pass) * kT2N;
        for (int gi = ps.g_begin; gi < ps.g_end; ++gi) {
          const TcGroup& G = p.g[gi];
          const int n_rows = kT2N + G.n_taps - 1;
          // row table: element offset of every window row in A (and in the mask tensor), -1 = zeros
          asm volatile("bar.sync 1, %0;" ::"r"(kT2Loaders) : "memory");          // the previous group's readers are done
          for (int i = tl; i < n_rows; i += kT2Loaders) {
            const long long u = v0 + i;
            const int b = (int)(u / p.Pv);
            const int pos = G.sg * (int)(u - (long long)b * p.Pv) + G.off0;
            int rr = pos;
            if (!p.bwd) {
              rr = rr < 0 ? -rr : rr;
              if (rr >= p.T_a) rr = 2 * (p.T_a - 1) - rr;
            }
            const bool ok = b < p.B && rr >= 0 && rr < p.T_a;
            rowtab[i] = ok ? (long long)b * p.a_bs + (long long)rr * p.a_rs + G.a_ch_off : -1;
            rowtab[kT2Rows + i] = (ok && p.Mk) ? (long long)b * p.m_bs + (long long)rr * p.m_rs + G.a_ch_off : -1;
          }
          asm volatile("bar.sync 1, %0;" ::"r"(kT2Loaders) : "memory");
          const int nkb = (G.kc + kTcKB - 1) / kTcKB;
          const int n_units = n_rows * 4;
          const bool masked = p.Mk != nullptr;
          // L2 prefetch kTcPfDist K blocks ahead: one 128-byte line per row and K block (lanes with ks == 0)
          auto prefetch = [&](int kb) {
            if (kb >= nkb) return;
            for (int q = tl; q < n_units; q += kT2Loaders)
              if ((q & 3) == 0) {
                const long long o = rowtab[q >> 2];
                if (o >= 0) { prefetch_l2(p.A + o + kb * kTcKB); if (masked) prefetch_l2(p.Mk + rowtab[kT2Rows + (q >> 2)] + kb * kTcKB); }
              }
          };
#pragma unroll 1
          for (int d = 0; d < kTcPfDist; ++d) prefetch(d);
#pragma unroll 1
          for (int kb = 0; kb < nkb; ++kb) {
            const int kb0 = kb * kTcKB, kbs = min(kTcKB, G.kc - kb0);
            prefetch(kb + kTcPfDist);
            float* hi = Xs + (size_t)sx * 2 * kT2XPlane;
            float* lo = hi + kT2XPlane;
            bool waited = false;
#pragma unroll 1
            for (int q0 = tl; q0 < n_units; q0 += 3 * kT2Loaders) {
              float4 va[3], vb[3], ma[3], mb[3];
#pragma unroll
              for (int j = 0; j < 3; ++j) {
                const int q = q0 + j * kT2Loaders;
                va[j] = f4zero(); vb[j] = f4zero(); ma[j] = f4zero(); mb[j] = f4zero();
                if (q < n_units) {
                  const int ks = q & 3;
                  const long long o = rowtab[q >> 2];
                  if (o >= 0 && 8 * ks < kbs) {
                    const float* src = p.A + o + kb0 + 8 * ks;
                    va[j] = ld4(src); vb[j] = ld4(src + 4);
                    if (masked) { const float* ms = p.Mk + rowtab[kT2Rows + (q >> 2)] + kb0 + 8 * ks; ma[j] = ld4(ms); mb[j] = ld4(ms + 4); }
                  }
                }
              }
              if (!waited) { mbar_wait(x_empty(sx), px ^ 1); waited = true; }     // the first loads fly while the stage is released
#pragma unroll
              for (int j = 0; j < 3; ++j) {
                const int q = q0 + j * kT2Loaders;
                if (q < n_units) {
                  const int ks = q & 3, row = q >> 2;
                  if (8 * ks < kbs) {
                    float4 a = va[j], b = vb[j];
                    if (masked) { a = dact4mul(a, ma[j], p.slope); b = dact4mul(b, mb[j], p.slope); }
                    const float4 ha = make_float4(tf32_hi(a.x), tf32_hi(a.y), tf32_hi(a.z), tf32_hi(a.w));
                    const float4 hb = make_float4(tf32_hi(b.x), tf32_hi(b.y), tf32_hi(b.z), tf32_hi(b.w));
                    st4(hi + ((size_t)(2 * ks) * kT2Rows + row) * 4, ha);
                    st4(hi + ((size_t)(2 * ks + 1) * kT2Rows + row) * 4, hb);
                    st_bf16x8(lo + ((size_t)(2 * ks) * kT2Rows + row) * 4, f4sub(a, ha), f4sub(b, hb));
                    st_bf16x8(lo + ((size_t)(2 * ks + 1) * kT2Rows + row) * 4, ha, hb);
                  }
                }
              }
            }
            if (!waited) mbar_wait(x_empty(sx), px ^ 1);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(x_full(sx));
            if (++sx == kT2XStages) { sx = 0; px ^= 1; }
          }
        }
      }
    }
  } else {
    // ===== drain warps (256 threads): chunk sums in registers, then the epilogue straight from registers =====
    asm volatile("setmaxnreg.inc.sync.aligned.u32 168;");
    pdl_wait();
    const int quad = warp & 3, half = (warp - 8) >> 2;
    const int dt = threadIdx.x - 256;                   // 0..255: the tile row whose descriptor this thread works out
    int chunk = 0;
    int mode = 0;
    if (p.Om) mode = 1;
    else if (p.res.mode == RES_POOL || p.res.mode == RES_UP_BWD) mode = 3;
    else if (p.res.mode != RES_NONE) mode = 2;
    for (int wk = w_first; wk < n_work; wk += w_step) {
      const TcPass& ps = p.pass[wk % p.n_pass];
      const long long v0 = (long long)(wk / p.n_pass) * kT2N;
      const int cl = quad * 32 + lane;                  // channel within the pass (TMEM lane)
      const bool c_ok = cl < ps.N;
      const bool q_ok = quad * 32 < ps.N;               // this warp's lanes hold any valid channel at all
      // ---- row descriptors of this tile (while the first chunk's MMAs run) ----
      asm volatile("bar.sync 2, 256;" ::: "memory");    // the previous tile's epilogue has read its descriptors
      {
        T2Desc d; float rs;
        t2_make_desc(p, ps, v0 + dt, d, rs);
        desc[dt] = d; rsv[dt] = rs;
        if (kTcPfDist > 0 && d.yoff >= 0) {             // the epilogue's operands for this row: into L2 now, read ~20k clk later
          const float* G = p.Om ? p.Om : p.res.R;
          for (int c0 = 0; c0 < ps.N; c0 += 32) {
            if (d.g0 >= 0) prefetch_l2(G + d.g0 + ps.ch_off + c0);
            if (d.g1 >= 0) prefetch_l2(G + d.g1 + ps.ch_off + c0);
          }
        }
      }
      asm volatile("bar.sync 2, 256;" ::: "memory");
      float acc[128];
#pragma unroll
      for (int i = 0; i < 128; ++i) acc[i] = 0.f;
#pragma unroll 1
      for (int cc = 0; cc < ps.n_chunks; ++cc) {
        const int buf = chunk % kT2AccBufs;
        mbar_wait(acc_full0 + 8 * buf, (chunk / kT2AccBufs) & 1);
        tc_fence_after();
        if (q_ok) {
          const uint32_t t0 = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * kT2N + half * 128);
#pragma unroll
          for (int c0 = 0; c0 < 128; c0 += 32) {
            uint32_t r0[16], r1[16];
            tmem_ld16(t0 + c0, r0);
            tmem_ld16(t0 + c0 + 16, r1);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) { acc[c0 + i] += __uint_as_float(r0[i]); acc[c0 + 16 + i] += __uint_as_float(r1[i]); }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty0 + 8 * buf);
        ++chunk;
      }
      if (!q_ok) continue;
      const int c = ps.ch_off + cl;
      const T2Desc* dh = desc + half * 128;
      const float* rh = rsv + half * 128;
      switch (mode) {      // one specialised straight-line epilogue per launch stays hot in the instruction cache
        case 0: t2_epilogue<0>(p, acc, dh, rh, c, c_ok); break;
        case 1: t2_epilogue<1>(p, acc, dh, rh, c, c_ok); break;
        case 2: t2_epilogue<2>(p, acc, dh, rh, c, c_ok); break;
        default: t2_epilogue<3>(p, acc, dh, rh, c, c_ok); break;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

}  // namespace avc
