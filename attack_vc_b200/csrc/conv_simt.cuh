// conv_simt.cuh -- generic "tap-list" implicit-GEMM Conv1d on CUDA cores (exact fp32 FFMA).
//
// One kernel covers every contraction of AdaIN-VC (SURVEY.md §8 contraction table) in both
// directions:
//   forward  (reference pad_layer + nn.Conv1d, models.py:10-30): reflect padding by index math,
//            stride, bias, activation, optional second (pre-residual) output and pooled residual;
//   dgrad    (autograd of the same): zero-stuffed gather of dy, reversed taps, and the reflect-pad
//            "fold" (rows that were mirrored in forward receive their mirrored gradient) done by
//            giving edge rows up to three gather positions -- no atomics, fixed summation order.
// It is the path used when the GEMM M dimension is small (batch-1 attacks, config 1/2 of
// BASELINE.json), where spreading exact-fp32 FFMA work over all 148 SMs beats a two-CTA tensor
// core tile, and it is the bit-stable reference the tcgen05 path (conv_tc.cuh) is tested against.
//
//   Y[b, t, n] = epi( sum_g sum_i sum_c  Awin_g[b, pos(t) + off0_g + i, a_ch_off_g + c] * W_g[i][c][n] )
//
// Awin is the *window view* of the A tensor: forward -> A[reflect(r)], dgrad -> A[r/s] if r>=0,
// r%s==0, r/s<T_a, else 0.
#pragma once
#include "common.cuh"

namespace avc {

constexpr int kMaxGroups = 10;
constexpr int kMaxTaps = 8;
constexpr int kMaxZk = 3;
constexpr int kKSub = 32;        // K rows of a weight slab staged per pipeline step
constexpr int kSRow = 128 + 4;   // smem row pitch of the activation window (floats)

struct TapGroup {
  const float* W;   // [n_taps][wts][N], N contiguous; this group uses rows [0,kc) of every tap
  int a_ch_off;     // first channel of A this group contracts
  int kc;           // channels contracted (multiple of 4, <= 128)
  int n_taps;
  int off0;         // window-row offset of tap 0 (fwd: -pad_left ; dgrad: pad_left-(k-1))
  int pl, pr;       // reflect pads of the forward conv (dgrad fold only)
  int wts;          // rows between consecutive taps in W (>= kc; the full K of the packed image)
};

// Folding of the decoder's InstanceNorm/AdaIN/activation kernels into the small-M conv kernel (batch-1 attacks, where
// every separate launch is ~6 us of dependent latency).  The PRODUCER conv writes, next to its raw output, per-tile
// partial statistics; the CONSUMER conv finalises them (fixed order) and applies the normalisation while it loads
// its activation window.  Same arithmetic per element as norm_act_fwd_kernel / norm_act_bwd_kernel.
struct FoldPro {
  int mode;             // 0 none; 1 window = act(AdaIN(IN(A))) [+ res]; 2 window = d/dy of that, A = upstream gradient
  int C, T;             // channels / rows of the normalised tensor (window channel n -> n % C)
  const float* part;    // [B][n_part][C][2] producer partials: mode 1 (sum, M2 about the tile mean), mode 2 (S1, S2)
  int n_part, part_tm, part_up, part_T;   // every partial covers part_tm rows (part_T % part_tm == 0 or the host does not fold)
  const float* stats;   // finalised [B][C][2] (mean, rstd): mode 1 when n_part == 0, mode 2 always
  float* stats_out;     // mode 1: finalised statistics for the backward pass
  const float* cond; int cond_bs;   // AdaIN mean at [0,C), std at [C,2C) of row b
  const float* y;       // mode 2: raw conv output saved by the forward pass, indexed like A
  float* gcond; int gcond_bs;       // mode 2: d mean / d std
  ResArgs res;          // mode 1: skip connection added after the activation
  float* out;           // mode 1: optional materialised [B][T][C] result (own rows of the column-tile-0 CTAs)
};
struct FoldEpi {
  int mode;             // 0 none; 1 partial statistics of the output; 2 partial S1 = sum ga, S2 = sum ga*xhat
  int C;
  float* part; int n_part;
  const float* y; const float* stats; const float* cond; int cond_bs;   // mode 2 (y indexed like Y)
};
constexpr int kFoldPC = 6 * 128;
constexpr int kFoldMaxPart = 64;   // most partials one normalisation may be split into (conv_small_kernel keeps them in registers)   // per-channel constants in shared memory (floats)

struct ConvArgs {
  const float* A; long long a_bs; int a_rs; int T_a;     // operand rows per utterance
  const float* Mk; long long m_bs; int m_rs;             // optional: a *= act'(Mk) on load
  float slope;
  int bwd;        // 0 forward window (reflect), 1 dgrad window (zero-stuffed)
  int s;          // conv stride
  int T_y; int B;
  float* Y; long long y_bs; int y_rs; int N;
  float* Y2; long long y2_bs; int y2_rs;                 // optional copy taken before the residual add
  const float* bias;
  int act;                                               // apply act() in the epilogue
  const float* Om; long long om_bs; int om_rs;           // optional: v *= act'(Om[t]) (dgrad through an activation)
  ResArgs res;
  int zsplit;     // 1: blockIdx.z selects ONE group; its outputs go to channels [z*N, z*N+N) (conv bank)
  int n_groups;
  int win_rows;   // activation-window rows available in smem (zero rows follow)
  // conv_small_kernel only: first smem window row of every group (all windows are resident at once),
  // win_off[n_groups] = first zero row; ring = number of weight-slab buffers, slab_floats = size of one
  int win_off[kMaxGroups + 1];
  int ring, slab_floats;
  // K split over blockIdx.z (conv_small_kernel): CTA z contracts groups [zk_lo[z], zk_hi[z]) only; z = 0 writes Y (with bias /
  // residual), z > 0 writes the contiguous [B][T_y][N] partial y_part[z-1]; the consumer adds the partials in order
  int zk, zk_lo[kMaxZk], zk_hi[kMaxZk];
  float* y_part[kMaxZk - 1];
  int e_rows, e_off[kMaxGroups];   // dgrad: rows behind the zero rows holding the pre-summed operands of reflect-edge output rows (0: sum in the loop)
  int f_stage, f_s2, f_epi;   // float offsets of the fold regions behind the weight ring (conv_small_kernel)
  FoldPro pro;
  FoldEpi epi;
  TapGroup g[kMaxGroups];
};

// Rows of the activation window one output tile [t0, t1) of group G needs (shared by host and device).
struct WinGeom {
  int wlo, nrows;                      // first gather position (may be negative) and row count
  int lt_lo, lt_hi, rt_lo, rt_hi;      // dgrad: output rows that also receive a mirrored (reflect-pad) gradient
  bool edge;
};
__host__ __device__ inline WinGeom win_geom(int bwd, int s, int T_y, const TapGroup& G, int t0, int t1) {
  const int sv = bwd ? 1 : s;
  int pmin = t0 * sv, pmax = (t1 - 1) * sv;
  WinGeom w;
  w.edge = false; w.lt_lo = 0; w.lt_hi = -1; w.rt_lo = 0; w.rt_hi = -1;
  if (bwd) {
    w.lt_lo = t0 > 1 ? t0 : 1; w.lt_hi = (t1 - 1) < G.pl ? (t1 - 1) : G.pl;
    w.rt_lo = t0 > (T_y - 1 - G.pr) ? t0 : (T_y - 1 - G.pr); w.rt_hi = (t1 - 1) < (T_y - 2) ? (t1 - 1) : (T_y - 2);
    if (w.lt_lo <= w.lt_hi) { w.edge = true; if (-w.lt_hi < pmin) pmin = -w.lt_hi; if (-w.lt_lo > pmax) pmax = -w.lt_lo; }
    if (w.rt_lo <= w.rt_hi) {
      w.edge = true;
      const int a = 2 * (T_y - 1) - w.rt_hi, b = 2 * (T_y - 1) - w.rt_lo;
      if (a < pmin) pmin = a;
      if (b > pmax) pmax = b;
    }
  }
  w.wlo = pmin + G.off0;
  w.nrows = pmax + G.off0 + G.n_taps - 1 - w.wlo + 1;
  return w;
}

__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc, bool pred) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  int bytes = pred ? 16 : 0;   // src-size 0 => zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(sa), "l"(gsrc), "r"(bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }
// wait until at most n groups are pending (n is warp-uniform, 0..7)
__device__ __forceinline__ void cp_async_wait_dyn(int n) {
  switch (n) {
    case 0: cp_async_wait<0>(); break;
    case 1: cp_async_wait<1>(); break;
    case 2: cp_async_wait<2>(); break;
    case 3: cp_async_wait<3>(); break;
    case 4: cp_async_wait<4>(); break;
    case 5: cp_async_wait<5>(); break;
    case 6: cp_async_wait<6>(); break;
    default: cp_async_wait<7>(); break;
  }
}

template <int RM, int TXN, int TYN, bool EDGE>
__device__ __forceinline__ void conv_accumulate_slab(float (&acc)[RM][4], const float* __restrict__ S,
                                                     const float* __restrict__ Wsub, const int (&rb)[RM][3],
                                                     int tap, int kk, int tx) {
  constexpr int TN = TXN * 4;
#pragma unroll 2
  for (int k4 = 0; k4 < kKSub; k4 += 4) {
    float4 a[RM];
#pragma unroll
    for (int i = 0; i < RM; ++i) {
      a[i] = ld4(S + (rb[i][0] + tap) * kSRow + kk + k4);
      if (EDGE) {
        a[i] = f4add(a[i], ld4(S + (rb[i][1] + tap) * kSRow + kk + k4));
        a[i] = f4add(a[i], ld4(S + (rb[i][2] + tap) * kSRow + kk + k4));
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 w = ld4(Wsub + (k4 + q) * TN + tx * 4);
#pragma unroll
      for (int i = 0; i < RM; ++i) {
        const float av = q == 0 ? a[i].x : q == 1 ? a[i].y : q == 2 ? a[i].z : a[i].w;
        acc[i][0] = fmaf(av, w.x, acc[i][0]);
        acc[i][1] = fmaf(av, w.y, acc[i][1]);
        acc[i][2] = fmaf(av, w.z, acc[i][2]);
        acc[i][3] = fmaf(av, w.w, acc[i][3]);
      }
    }
  }
}

// RM rows x 4 cols per thread; TXN threads along N, TYN along time.  Tile = (RM*TYN) x (4*TXN).
template <int RM, int TXN, int TYN>
__global__ void __launch_bounds__(TXN* TYN) conv_simt_kernel(const ConvArgs p) {
  constexpr int TM = RM * TYN, TN = TXN * 4, NT = TXN * TYN;
  extern __shared__ __align__(16) float smem[];
  float* S = smem;                                         // [(win_rows + kMaxTaps)][kSRow]
  float* Wb = smem + (size_t)(p.win_rows + kMaxTaps) * kSRow;  // [2][kKSub][TN]

  pdl_enter();
  const int tid = threadIdx.x, tx = tid % TXN, ty = tid / TXN;
  const int tiles_t = (p.T_y + TM - 1) / TM;
  const int b = blockIdx.x / tiles_t, t0 = (blockIdx.x % tiles_t) * TM;
  const int t1 = min(t0 + TM, p.T_y);
  const int n0 = blockIdx.y * TN;
  const int zr = p.win_rows;  // first zero row

  // zero rows (never overwritten afterwards)
  for (int i = tid; i < kMaxTaps * kSRow; i += NT) S[zr * kSRow + i] = 0.f;

  float acc[RM][4];
#pragma unroll
  for (int i = 0; i < RM; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;

  const int g_lo = p.zsplit ? blockIdx.z : 0;
  const int g_hi = p.zsplit ? blockIdx.z + 1 : p.n_groups;
  const int sv = p.bwd ? 1 : p.s;

  for (int gi = g_lo; gi < g_hi; ++gi) {
    const TapGroup& G = p.g[gi];
    const int kc = G.kc, nt = G.n_taps;
    const int kcp = (kc + kKSub - 1) / kKSub * kKSub;   // padded K (zero filled)
    // ---- gather positions of this tile -------------------------------------------------
    int pmin = t0 * sv, pmax = (t1 - 1) * sv;
    bool edge = false;
    int lt_lo = 0, lt_hi = -1, rt_lo = 0, rt_hi = -1;
    if (p.bwd) {
      lt_lo = max(t0, 1); lt_hi = min(t1 - 1, G.pl);                     // rows mirrored at the left edge
      rt_lo = max(t0, p.T_y - 1 - G.pr); rt_hi = min(t1 - 1, p.T_y - 2); // ... at the right edge
      if (lt_lo <= lt_hi) { edge = true; pmin = min(pmin, -lt_hi); pmax = max(pmax, -lt_lo); }
      if (rt_lo <= rt_hi) { edge = true; pmin = min(pmin, 2 * (p.T_y - 1) - rt_hi); pmax = max(pmax, 2 * (p.T_y - 1) - rt_lo); }
    }
    const int wlo = pmin + G.off0;
    const int nrows = pmax + G.off0 + nt - 1 - wlo + 1;   // <= win_rows by construction (host)

    if (nrows > p.win_rows) __trap();   // host sized the window too small: fail loudly, never corrupt
    __syncthreads();   // previous group's math (and the zero-row fill) is done
    // ---- activation window -> smem ------------------------------------------------------
    {
      const int c4n = kcp >> 2;
      const float* Ab = p.A + (long long)b * p.a_bs + G.a_ch_off;
      const float* Mb = p.Mk ? p.Mk + (long long)b * p.m_bs + G.a_ch_off : nullptr;
      for (int idx = tid; idx < nrows * c4n; idx += NT) {
        const int row = idx / c4n, c = (idx - row * c4n) << 2;
        const int r = wlo + row;
        int rr; bool ok;
        if (!p.bwd) {
          rr = r < 0 ? -r : r;
          if (rr >= p.T_a) rr = 2 * (p.T_a - 1) - rr;
          ok = rr >= 0 && rr < p.T_a;
        } else {
          ok = r >= 0 && (r % p.s) == 0;
          rr = r / p.s;
          ok = ok && rr < p.T_a;
        }
        float4 v = f4zero();
        if (ok && c < kc) {
          v = ld4(Ab + (long long)rr * p.a_rs + c);
          if (Mb) v = dact4mul(v, ld4(Mb + (long long)rr * p.m_rs + c), p.slope);
        }
        st4(S + row * kSRow + c, v);
      }
    }
    // ---- per-thread window rows ---------------------------------------------------------
    int rb[RM][3];
#pragma unroll
    for (int i = 0; i < RM; ++i) {
      const int t = t0 + ty + i * TYN;
      const bool tv = t < t1;
      rb[i][0] = tv ? t * sv + G.off0 - wlo : zr;
      rb[i][1] = (tv && t >= lt_lo && t <= lt_hi) ? -t + G.off0 - wlo : zr;
      rb[i][2] = (tv && t >= rt_lo && t <= rt_hi) ? 2 * (p.T_y - 1) - t + G.off0 - wlo : zr;
    }
    // ---- weight slabs, double buffered with cp.async --------------------------------------
    const int nk = kcp / kKSub, nslab = nt * nk;
    const int Ntot = p.N;
    auto issue = [&](int slab, int buf) {
      const int tap = slab / nk, kk = (slab - tap * nk) * kKSub;
      const float* Wg = G.W + ((long long)tap * G.wts + kk) * Ntot + n0;
      float* dst = Wb + buf * (kKSub * TN);
      for (int idx = tid; idx < kKSub * (TN / 4); idx += NT) {
        const int k = idx / (TN / 4), n = (idx - k * (TN / 4)) << 2;
        const bool ok = (kk + k) < kc && (n0 + n) < Ntot;
        cp_async16(dst + k * TN + n, ok ? Wg + (long long)k * Ntot + n : G.W, ok);
      }
      cp_async_commit();
    };
    issue(0, 0);
    for (int slab = 0; slab < nslab; ++slab) {
      const int buf = slab & 1;
      if (slab + 1 < nslab) { issue(slab + 1, buf ^ 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
      __syncthreads();   // slab `slab` (and, first time, the window) visible to all
      const int tap = slab / nk, kk = (slab - tap * nk) * kKSub;
      const float* Wsub = Wb + buf * (kKSub * TN);
      if (edge) conv_accumulate_slab<RM, TXN, TYN, true>(acc, S, Wsub, rb, tap, kk, tx);
      else      conv_accumulate_slab<RM, TXN, TYN, false>(acc, S, Wsub, rb, tap, kk, tx);
      __syncthreads();   // all reads of this buffer done before it is refilled
    }
  }

  // ---- epilogue ---------------------------------------------------------------------------
  const int n = n0 + tx * 4;
  if (n >= p.N) return;
  const int ch = p.zsplit ? blockIdx.z * p.N + n : n;   // output channel
  float4 bias4 = p.bias ? ld4(p.bias + ch) : f4zero();
#pragma unroll
  for (int i = 0; i < RM; ++i) {
    const int t = t0 + ty + i * TYN;
    if (t >= t1) continue;
    float4 v = make_float4(acc[i][0] + bias4.x, acc[i][1] + bias4.y, acc[i][2] + bias4.z, acc[i][3] + bias4.w);
    if (p.Om) v = dact4mul(v, ld4(p.Om + (long long)b * p.om_bs + (long long)t * p.om_rs + ch), p.slope);
    if (p.act) v = act4(v, p.slope);
    if (p.Y2) st4(p.Y2 + (long long)b * p.y2_bs + (long long)t * p.y2_rs + ch, v);
    if (p.res.mode != RES_NONE) v = f4add(v, res_load4(p.res, b, t, p.T_y, ch));
    st4(p.Y + (long long)b * p.y_bs + (long long)t * p.y_rs + ch, v);
  }
}


// =================================================================================================
// conv_small_kernel -- the same contraction for SMALL M (batch-1 attacks: M = 32..256 rows).
// There a layer is latency-bound, not throughput-bound, so the tile is tiny (TM x 32 outputs, one
// CTA per SM), every group's activation window is resident at once, weight slabs (one tap x <=128
// channels x 32 columns) stream through a deep cp.async ring issued before anything else, the 8
// warps split K (16 rows of every slab each) and their partial tiles are summed in a fixed order.
// =================================================================================================
constexpr int kSmTN = 32;
constexpr int kSmWarps = 16;        // K slices per slab: warp w contracts rows [kSmKW*w, kSmKW*(w+1))
constexpr int kSmKW = 128 / kSmWarps;

#ifdef AVC_SMALL_PROFILE
// debug build only: per-launch phase stamps of CTA 0 (globaltimer ns at entry/exit, SM clock in between)
__device__ unsigned long long g_small_prof[8192][12];
__device__ unsigned int g_small_prof_n;
#define SMALL_STAMP(i) do { if (prof_slot < 8192u && threadIdx.x == 0) g_small_prof[prof_slot][i] = clock64(); } while (0)
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#else
#define SMALL_STAMP(i) do {} while (0)
#endif
#ifndef AVC_SMALL_MINB
#define AVC_SMALL_MINB 1
#endif
// n % C for the few wraps the folded decoder has (n < 2C with pixel shuffle x2): no integer division
__device__ __forceinline__ int fold_ch(int n, int C) { while (n >= C) n -= C; return n; }

template <int TM>
__global__ void __launch_bounds__(32 * kSmWarps, AVC_SMALL_MINB) conv_small_kernel(const ConvArgs p) {
  constexpr int RM = TM / 4, NT = 32 * kSmWarps;
  extern __shared__ __align__(16) float smem[];
  const int g_lo = p.zk > 1 ? p.zk_lo[blockIdx.z] : p.zsplit ? blockIdx.z : 0;
  const int g_hi = p.zk > 1 ? p.zk_hi[blockIdx.z] : p.zsplit ? blockIdx.z + 1 : p.n_groups;
  float* S = smem;                                                    // windows, then kMaxTaps zero rows
  const int zr = p.win_off[p.n_groups];
  float* Wr = smem + (size_t)(zr + kMaxTaps + p.e_rows) * kSRow;      // [ring][slab_floats]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tx = lane & 7, ty = lane >> 3;
  const int tiles_t = (p.T_y + TM - 1) / TM;
  const int b = blockIdx.x / tiles_t, t0 = (blockIdx.x % tiles_t) * TM;
  const int t1 = min(t0 + TM, p.T_y);
  const int n0 = blockIdx.y * kSmTN;
  const int Ntot = p.N;
  const int sv = p.bwd ? 1 : p.s;

  int n_slabs = 0;
  for (int gi = g_lo; gi < g_hi; ++gi) n_slabs += p.g[gi].n_taps;
#ifdef AVC_SMALL_PROFILE
  unsigned prof_slot = 0xffffffffu;
  if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0) {
    prof_slot = atomicAdd(&g_small_prof_n, 1u);
    if (prof_slot < 8192u) { g_small_prof[prof_slot][0] = gtimer(); g_small_prof[prof_slot][7] = ((unsigned long long)gridDim.x << 32) | (gridDim.y << 16) | (p.pro.mode << 4) | p.epi.mode; }
  }
  SMALL_STAMP(1);
#endif

  // ---- weight ring: slab q = (group, tap) in issue order ----------------------------------------
  // thread constants of the copy: 16-byte column chunk wn of K rows wk, wk + 32, ...
  const int wk = tid >> 3, wn = (tid & 7) << 2;
  const bool wok = (n0 + wn) < Ntot;
  // Issue state kept as running pointers: the kernel is instruction-issue bound, and 7 slabs x ~70 instructions of
  // 64-bit index arithmetic per thread used to cost as much as the main loop.  Per slab now: two copies, three adds.
  constexpr int kRows = NT / 8;                  // K rows covered per pass of the CTA
  int is_g = g_lo, is_left = 0, is_slot = 0;     // group of the next slab, taps left in it, ring slot it goes to
  int is_kc = 0;
  const float* is_src = nullptr;                 // row wk of the next slab, this thread's 16-byte column
  long long is_tap_step = 0, is_row_step = 0;
  float* is_dst = Wr + wk * kSmTN + wn;
  auto issue_group = [&]() {
    const TapGroup& G = p.g[is_g];
    is_left = G.n_taps; is_kc = G.kc;
    is_src = wok ? G.W + (long long)wk * Ntot + n0 + wn : G.W;
    is_tap_step = wok ? (long long)G.wts * Ntot : 0;
    is_row_step = wok ? (long long)kRows * Ntot : 0;
  };
  if (n_slabs > 0) issue_group();
  auto issue = [&](int q) {
    if (q < n_slabs) {
      if (wk < is_kc) cp_async16(is_dst, is_src, wok);
      if (wk + kRows < is_kc) cp_async16(is_dst + kRows * kSmTN, is_src + is_row_step, wok);
      is_src += is_tap_step;
      is_dst += p.slab_floats;
      if (++is_slot == p.ring) { is_slot = 0; is_dst -= (size_t)p.ring * p.slab_floats; }
      if (--is_left == 0 && ++is_g < g_hi) issue_group();
    }
    cp_async_commit();
  };
  pdl_launch_dependents();
  const bool all_resident = n_slabs <= p.ring;   // every slab has its own buffer: no refills, no per-slab barrier
  const int n_pre = all_resident ? n_slabs : p.ring - 1;
  for (int q = 0; q < n_pre; ++q) issue(q);
  pdl_wait();   // weights are loop constants; the windows below are the predecessor's output
  SMALL_STAMP(2);

  // ---- activation windows -> smem (loads of all groups in flight together) --------------------------
  // Folded normalisation (FoldPro / FoldEpi): EVERY global read this CTA needs -- producer partials, raw window,
  // the second operand of the transform (skip connection or saved conv output), the epilogue's operands -- is issued
  // here as one batch of cp.async, so the fold adds no extra L2 round trip to the dependent chain of the iteration.
  float* PC = Wr + (size_t)p.ring * p.slab_floats;                    // [6][128]: mu, rstd, std, mean, m1, m2
  float* stage = PC + p.f_stage;                                      // [n_part][C][2] producer partials
  float* S2 = PC + p.f_s2;                                            // second-operand window, same geometry as S
  float* EP = PC + p.f_epi;                                           // epilogue operands: y tile [TM][32], stats [64], mean [32], std [32]
  const int fmode = p.pro.mode;
  const bool second = fmode == 2 || (fmode == 1 && p.pro.res.mode != RES_NONE);
  for (int i = tid; i < kMaxTaps * kSRow; i += NT) S[zr * kSRow + i] = 0.f;
  // walk(fn): fn(G, smem float offset of the element, c, r, rr, ok) for every float4 of every window
  auto walk = [&](auto&& fn) {
    for (int gi = g_lo; gi < g_hi; ++gi) {
      const TapGroup& G = p.g[gi];
      const WinGeom wg = win_geom(p.bwd, p.s, p.T_y, G, t0, t1);
      const int woff = p.zsplit ? 0 : p.win_off[gi];
      if (woff + wg.nrows > zr) __trap();   // host sized the window too small
      const int c4n = G.kc >> 2;
      // idx += NT without a division per element (and none at all for the usual 128-channel group)
      int drow, row;
      if (c4n == 32) { drow = NT / 32; row = tid >> 5; } else { drow = NT / c4n; row = tid / c4n; }
      const int dcol = NT - drow * c4n;
      int col = tid - row * c4n;
      for (int idx = tid; idx < wg.nrows * c4n; idx += NT, row += drow, col += dcol) {
        if (col >= c4n) { col -= c4n; ++row; }
        const int c = col << 2;
        const int r = wg.wlo + row;
        int rr; bool ok;
        if (!p.bwd) {
          rr = r < 0 ? -r : r;
          if (rr >= p.T_a) rr = 2 * (p.T_a - 1) - rr;
          ok = rr >= 0 && rr < p.T_a;
        } else if (p.s == 1) {
          rr = r;
          ok = r >= 0 && r < p.T_a;
        } else {
          ok = r >= 0 && (r % p.s) == 0;
          rr = r / p.s;
          ok = ok && rr < p.T_a;
        }
        fn(G, (woff + row) * kSRow + c, c, r, rr, ok);
      }
    }
  };
  if (!fmode && !p.Mk) {
    // nothing to do to the operand on the way in: asynchronous copies, all in flight at once (they join the first weight
    // group the main loop waits for)
    walk([&](const TapGroup& G, int so, int c, int r, int rr, bool ok) {
      cp_async16(S + so, ok ? p.A + (long long)b * p.a_bs + G.a_ch_off + (long long)rr * p.a_rs + c : p.A, ok);
    });
    cp_async_commit();
  } else if (!fmode) {
    walk([&](const TapGroup& G, int so, int c, int r, int rr, bool ok) {
      float4 v = f4zero();
      if (ok) {
        v = ld4(p.A + (long long)b * p.a_bs + G.a_ch_off + (long long)rr * p.a_rs + c);
        if (p.Mk) v = dact4mul(v, ld4(p.Mk + (long long)b * p.m_bs + G.a_ch_off + (long long)rr * p.m_rs + c), p.slope);
      }
      st4(S + so, v);
    });
  } else {
    const FoldPro& f = p.pro;
    const int nfl = f.n_part * f.C * 2;
    const float* src = f.part + (long long)b * nfl;
    for (int i = tid * 4; i < nfl; i += NT * 4) cp_async16(stage + i, src + i, true);
    walk([&](const TapGroup& G, int so, int c, int r, int rr, bool ok) {
      const float* a = p.A + (long long)b * p.a_bs + G.a_ch_off + (long long)rr * p.a_rs + c;
      cp_async16(S + so, ok ? a : p.A, ok);
      if (second) {
        const float* q;
        if (fmode == 2) q = f.y + (long long)b * p.a_bs + G.a_ch_off + (long long)rr * p.a_rs + c;
        else q = f.res.R + (long long)b * f.res.bs + (long long)(f.res.mode == RES_UP ? (f.res.rf == 2 ? rr >> 1 : rr / f.res.rf) : rr) * f.res.rs + fold_ch(G.a_ch_off + c, f.C);
        cp_async16(S2 + so, ok ? q : p.A, ok);
      }
    });
  }
  if (p.epi.mode == 2) {
    const FoldEpi& e = p.epi;
    const int cc0 = fold_ch(n0, e.C);
    if (tid < TM * 8) {
      const int er = tid >> 3, en = (tid & 7) << 2;
      const bool eok = t0 + er < t1 && n0 + en < p.N;
      cp_async16(EP + er * kSmTN + en, eok ? e.y + (long long)b * p.y_bs + (long long)(t0 + er) * p.y_rs + n0 + en : e.y, eok);
    } else if (tid < TM * 8 + 16) {
      const int j = tid - TM * 8;
      cp_async16(EP + TM * kSmTN + j * 4, e.stats + ((long long)b * e.C + cc0) * 2 + j * 4, true);
    } else if (tid < TM * 8 + 32) {
      const int j = tid - TM * 8 - 16;      // 0..7 mean, 8..15 std
      if (e.cond) cp_async16(EP + TM * kSmTN + 64 + j * 4, e.cond + (long long)b * e.cond_bs + (j >> 3) * e.C + cc0 + (j & 7) * 4, true);
      else st4(EP + TM * kSmTN + 64 + j * 4, j < 8 ? f4zero() : make_float4(1.f, 1.f, 1.f, 1.f));
    }
  }
  if (fmode) {
    const FoldPro& f = p.pro;
    // finalised statistics / AdaIN row of channel c (4 threads per channel, sub = partial residue class)
    const int c = tid >> 2, sub = tid & 3;
    const bool cv = c < f.C;
    float mu = 0.f, rstd = 0.f, cs = 1.f, cm = 0.f;
    if (cv) {
      if (f.mode == 2 || f.n_part == 0) {
        mu = f.stats[((long long)b * f.C + c) * 2];
        rstd = f.stats[((long long)b * f.C + c) * 2 + 1];
      }
      if (f.cond) { cs = f.cond[(long long)b * f.cond_bs + f.C + c]; cm = f.cond[(long long)b * f.cond_bs + c]; }
    }
    cp_async_commit();
    cp_async_wait<0>();
    SMALL_STAMP(10);
    __syncthreads();
    SMALL_STAMP(8);
    // fixed-order reduction of the partials: residue classes q = sub (mod 4) summed in order, then ((0+1)+(2+3))
    auto quad = [](float v) {
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      return v;
    };
    const int cq = cv ? c : 0;
    const float invT = __frcp_rn((float)f.T);
    // partial q of this channel: st2[q * C]; this thread sums q = sub, sub+4, ...  (every partial covers part_tm rows:
    // the host folds only when the producer's tiles are not ragged)
    const float2* st2 = reinterpret_cast<const float2*>(stage) + cq + sub * f.C;
    const int qstep = 4 * f.C, n_it = (f.n_part - sub + 3) >> 2;
    if (f.mode == 1 && f.n_part > 0) {
      const float cnt = (float)f.part_tm, inv_cnt = __frcp_rn(cnt);
      float sum = 0.f;
#pragma unroll 2
      for (int i = 0; i < n_it; ++i) sum += st2[i * qstep].x;
      mu = quad(sum) * invT;
      float m2 = 0.f;
#pragma unroll 2
      for (int i = 0; i < n_it; ++i) {
        const float2 pq = st2[i * qstep];
        const float d = fmaf(pq.x, inv_cnt, -mu);
        m2 += fmaf(cnt * d, d, pq.y);
      }
      rstd = rsqrtf(quad(m2) * invT + 1e-5f);
      if (cv && sub == 0 && f.stats_out && blockIdx.y == 0 && t0 == 0) {
        f.stats_out[((long long)b * f.C + c) * 2] = mu;
        f.stats_out[((long long)b * f.C + c) * 2 + 1] = rstd;
      }
    }
    float s1 = 0.f, s2 = 0.f;
    if (f.mode == 2) {
#pragma unroll 2
      for (int i = 0; i < n_it; ++i) { const float2 pq = st2[i * qstep]; s1 += pq.x; s2 += pq.y; }
      s1 = quad(s1); s2 = quad(s2);
      if (cv && sub == 0 && f.gcond && blockIdx.y == 0 && t0 == 0) {
        f.gcond[(long long)b * f.gcond_bs + c] = s1;
        f.gcond[(long long)b * f.gcond_bs + f.C + c] = s2;
      }
    }
    if (cv && sub == 0) {
      PC[c] = mu; PC[128 + c] = rstd; PC[256 + c] = cs; PC[384 + c] = cm;
      PC[512 + c] = s1 * invT; PC[640 + c] = s2 * invT;
    }
    SMALL_STAMP(9);
    __syncthreads();
    SMALL_STAMP(11);
    // transform the raw windows in place
    walk([&](const TapGroup& G, int so, int cch, int r, int rr, bool ok) {
      if (!ok) return;
      const int cc = fold_ch(G.a_ch_off + cch, f.C);
      float4 v = ld4(S + so);
      const float4 mu4 = ld4(PC + cc), rs = ld4(PC + 128 + cc), cs4 = ld4(PC + 256 + cc), cm4 = ld4(PC + 384 + cc);
      if (fmode == 1) {
        float4 a;
        a.x = fmaf((v.x - mu4.x) * rs.x, cs4.x, cm4.x); a.y = fmaf((v.y - mu4.y) * rs.y, cs4.y, cm4.y);
        a.z = fmaf((v.z - mu4.z) * rs.z, cs4.z, cm4.z); a.w = fmaf((v.w - mu4.w) * rs.w, cs4.w, cm4.w);
        v = act4(a, p.slope);
        if (second) v = f4add(v, ld4(S2 + so));
        if (f.out && blockIdx.y == 0 && r >= t0 && r < t1) st4(f.out + ((long long)b * f.T + rr) * f.C + cc, v);
      } else {
        const float4 yv = ld4(S2 + so);
        const float4 m1 = ld4(PC + 512 + cc), m2 = ld4(PC + 640 + cc);
        const float4 xh = make_float4((yv.x - mu4.x) * rs.x, (yv.y - mu4.y) * rs.y, (yv.z - mu4.z) * rs.z, (yv.w - mu4.w) * rs.w);
        const float4 a = make_float4(fmaf(xh.x, cs4.x, cm4.x), fmaf(xh.y, cs4.y, cm4.y), fmaf(xh.z, cs4.z, cm4.z), fmaf(xh.w, cs4.w, cm4.w));
        const float4 ga = dact4mul(v, a, p.slope);
        v.x = rs.x * cs4.x * (ga.x - m1.x - xh.x * m2.x); v.y = rs.y * cs4.y * (ga.y - m1.y - xh.y * m2.y);
        v.z = rs.z * cs4.z * (ga.z - m1.z - xh.z * m2.z); v.w = rs.w * cs4.w * (ga.w - m1.w - xh.w * m2.w);
      }
      st4(S + so, v);
    });
  } else if (p.epi.mode == 2) {
    cp_async_commit();     // the epilogue operands ride in their own group, older than every ring refill
  }

  // ---- dgrad tiles that touch a reflect-padded border: an output row t there also receives the gradient of the mirrored
  // positions.  Sum its 2-3 operand rows ONCE per tap into extra rows, so the main loop reads one operand row per output row.
  if (p.e_rows && p.bwd) {
    bool any = false;
    for (int gi = g_lo; gi < g_hi; ++gi) any = any || win_geom(p.bwd, p.s, p.T_y, p.g[gi], t0, t1).edge;
    if (any) {
      cp_async_wait<0>();
      __syncthreads();       // windows complete
      for (int gi = g_lo; gi < g_hi; ++gi) {
        const TapGroup& G = p.g[gi];
        const WinGeom wg = win_geom(p.bwd, p.s, p.T_y, G, t0, t1);
        if (!wg.edge) continue;
        const int nl = max(0, wg.lt_hi - wg.lt_lo + 1), nr = max(0, wg.rt_hi - wg.rt_lo + 1);
        const float* Sg0 = S + (size_t)p.win_off[gi] * kSRow;
        float* Eg = S + (size_t)p.e_off[gi] * kSRow;
        const int c4n = G.kc >> 2, n_pair = (nl + nr) * G.n_taps;
        // thread -> (16-byte column, first (row, tap) pair); no integer division (each thread has one or two pairs)
        int c4, pair, dpair;
        if (c4n == 32) { c4 = tid & 31; pair = tid >> 5; dpair = NT / 32; }
        else { c4 = tid % c4n; pair = tid / c4n; dpair = NT / c4n; if (dpair == 0) __trap(); }
        for (; pair < n_pair; pair += dpair) {
          int j = 0, tap = pair;
          while (tap >= G.n_taps) { tap -= G.n_taps; ++j; }
          const int t = j < nl ? wg.lt_lo + j : wg.rt_lo + (j - nl);
          const int r0 = t + G.off0 - wg.wlo + tap;
          const int r1 = (j < nl ? -t : 2 * (p.T_y - 1) - t) + G.off0 - wg.wlo + tap;
          st4(Eg + (size_t)pair * kSRow + c4 * 4, f4add(ld4(Sg0 + (size_t)r0 * kSRow + c4 * 4), ld4(Sg0 + (size_t)r1 * kSRow + c4 * 4)));
        }
      }
    }
  }
  SMALL_STAMP(3);
  // ---- main loop over slabs ------------------------------------------------------------------------
  float acc[RM][4];
#pragma unroll
  for (int i = 0; i < RM; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
  int rb[RM][3];
  bool edge = false;
  int cg = g_lo - 1, ctap = 0, cnt = 0;   // group / tap of the slab being consumed
  int c_slot = 0;                          // ring slot of the slab being consumed
  const float* Sg = S;
  int kc = 0;
  const int k_lo = warp * kSmKW;
  // running pointers (the loop is instruction-issue bound): operand rows of this thread at the current tap, and the
  // weight rows of this warp's K slice in the current ring slot
  const float* ap[RM];
  const float* wp = Wr + k_lo * kSmTN + tx * 4;
  for (int q = 0; q < n_slabs; ++q) {
    if (cnt == 0) {   // first slab of the next group: per-thread window rows
      ++cg; ctap = 0;
      const TapGroup& G = p.g[cg];
      cnt = G.n_taps; kc = G.kc;
      const WinGeom wg = win_geom(p.bwd, p.s, p.T_y, G, t0, t1);
      const bool pre = p.e_rows && p.bwd;      // mirrored operands were pre-summed above
      edge = wg.edge && !pre;
      const int wbase = p.zsplit ? 0 : p.win_off[cg];
      Sg = S + (size_t)wbase * kSRow;
      const int zrel = zr - wbase;
      const int nl = max(0, wg.lt_hi - wg.lt_lo + 1);
#pragma unroll
      for (int i = 0; i < RM; ++i) {
        const int t = t0 + ty * RM + i;
        const bool tv = t < t1;
        const bool lt = tv && t >= wg.lt_lo && t <= wg.lt_hi, rt = tv && t >= wg.rt_lo && t <= wg.rt_hi;
        rb[i][0] = tv ? t * sv + G.off0 - wg.wlo : zrel;
        rb[i][1] = lt ? -t + G.off0 - wg.wlo : zrel;
        rb[i][2] = rt ? 2 * (p.T_y - 1) - t + G.off0 - wg.wlo : zrel;
        if (pre && lt) rb[i][0] = p.e_off[cg] - wbase + (t - wg.lt_lo) * G.n_taps;
        if (pre && rt) rb[i][0] = p.e_off[cg] - wbase + (nl + t - wg.rt_lo) * G.n_taps;
        ap[i] = Sg + rb[i][0] * kSRow + k_lo;
      }
    }
    if (!all_resident) {
      if (q == 0) cp_async_wait<0>();    // the activation windows may ride in the youngest group
      else cp_async_wait_dyn(p.ring - 2);
      __syncthreads();                   // slab q landed for everyone; everyone is done with slab q-1 (and, first time, the windows are visible)
      issue(q + p.ring - 1);             // refills the buffer slab q-1 used
    } else if (q == 0) {
      cp_async_wait<0>();                // every slab was requested before the dependency wait: one barrier for the whole loop
      __syncthreads();
    }
    if (k_lo < kc) {     // kc is a multiple of kSmKW: a warp has all of its K rows or none
      // all shared-memory loads of the slab are issued before the first FMA
      float4 a[kSmKW / 4][RM], w[kSmKW];
#pragma unroll
      for (int j = 0; j < kSmKW / 4; ++j) {
#pragma unroll
        for (int i = 0; i < RM; ++i) {
          a[j][i] = ld4(ap[i] + 4 * j);
          if (edge) {
            a[j][i] = f4add(a[j][i], ld4(Sg + (rb[i][1] + ctap) * kSRow + k_lo + 4 * j));
            a[j][i] = f4add(a[j][i], ld4(Sg + (rb[i][2] + ctap) * kSRow + k_lo + 4 * j));
          }
        }
      }
#pragma unroll
      for (int k = 0; k < kSmKW; ++k) w[k] = ld4(wp + k * kSmTN);
#pragma unroll
      for (int k = 0; k < kSmKW; ++k) {
#pragma unroll
        for (int i = 0; i < RM; ++i) {
          const float4 av4 = a[k >> 2][i];
          const float av = (k & 3) == 0 ? av4.x : (k & 3) == 1 ? av4.y : (k & 3) == 2 ? av4.z : av4.w;
          acc[i][0] = fmaf(av, w[k].x, acc[i][0]);
          acc[i][1] = fmaf(av, w[k].y, acc[i][1]);
          acc[i][2] = fmaf(av, w[k].z, acc[i][2]);
          acc[i][3] = fmaf(av, w[k].w, acc[i][3]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < RM; ++i) ap[i] += kSRow;
    wp += p.slab_floats;
    if (++c_slot == p.ring) { c_slot = 0; wp -= (size_t)p.ring * p.slab_floats; }
    ++ctap; --cnt;
  }
  cp_async_wait<0>();
  __syncthreads();                       // all slabs consumed: the ring becomes the reduction buffer
  SMALL_STAMP(4);

  // ---- fixed-order sum of the 8 K-slices, then the epilogue ----------------------------------------
  float* red = Wr;                       // [kSmWarps][TM][kSmTN]
#pragma unroll
  for (int i = 0; i < RM; ++i)
    st4(red + ((size_t)warp * TM + ty * RM + i) * kSmTN + tx * 4, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
  __syncthreads();
  const int row = tid >> 3, nn = (tid & 7) << 2;
  const int t = t0 + row, n = n0 + nn;
  const bool active = tid < TM * (kSmTN / 4) && t < t1 && n < p.N;
#ifndef AVC_SMALL_PROFILE
  if (!active && !p.epi.mode) return;
#endif
  const int ch = p.zsplit ? blockIdx.z * p.N + n : n;   // output channel
  float4 v = f4zero();
  if (active) {
    v = ld4(red + (size_t)row * kSmTN + nn);
#pragma unroll
    for (int w = 1; w < kSmWarps; ++w) v = f4add(v, ld4(red + ((size_t)w * TM + row) * kSmTN + nn));
    const bool first = p.zk <= 1 || blockIdx.z == 0;
    if (p.bias && first) v = f4add(v, ld4(p.bias + ch));
    if (p.Om) v = dact4mul(v, ld4(p.Om + (long long)b * p.om_bs + (long long)t * p.om_rs + ch), p.slope);
    if (p.act) v = act4(v, p.slope);
    if (p.Y2) st4(p.Y2 + (long long)b * p.y2_bs + (long long)t * p.y2_rs + ch, v);
    if (p.res.mode != RES_NONE && first) v = f4add(v, res_load4(p.res, b, t, p.T_y, ch));
    if (first) st4(p.Y + (long long)b * p.y_bs + (long long)t * p.y_rs + ch, v);
    else st4(p.y_part[blockIdx.z - 1] + ((long long)b * p.T_y + t) * p.N + ch, v);
  }
#ifdef AVC_SMALL_PROFILE
  SMALL_STAMP(5);
  if (prof_slot < 8192u && threadIdx.x == 0) g_small_prof[prof_slot][6] = gtimer();
#endif
  if (!p.epi.mode) return;
  // ---- folded normalisation: per-tile partial sums over the rows of this tile, one thread per column ----
  const FoldEpi& e = p.epi;
  float* tile = S;                       // [2][TM][kSmTN]; the windows are dead
  float4 e1 = v, e2 = f4zero();
  if (e.mode == 2 && active) {
    const float* EPc = Wr + (size_t)p.ring * p.slab_floats + p.f_epi;
    const float4 s0 = ld4(EPc + TM * kSmTN + nn * 2), s1 = ld4(EPc + TM * kSmTN + nn * 2 + 4);
    const float4 mu = make_float4(s0.x, s0.z, s1.x, s1.z), rs = make_float4(s0.y, s0.w, s1.y, s1.w);
    const float4 cm = ld4(EPc + TM * kSmTN + 64 + nn), cs = ld4(EPc + TM * kSmTN + 96 + nn);
    const float4 yv = ld4(EPc + row * kSmTN + nn);
    const float4 xh = make_float4((yv.x - mu.x) * rs.x, (yv.y - mu.y) * rs.y, (yv.z - mu.z) * rs.z, (yv.w - mu.w) * rs.w);
    const float4 a = make_float4(fmaf(xh.x, cs.x, cm.x), fmaf(xh.y, cs.y, cm.y), fmaf(xh.z, cs.z, cm.z), fmaf(xh.w, cs.w, cm.w));
    e1 = dact4mul(v, a, p.slope);
    e2 = make_float4(e1.x * xh.x, e1.y * xh.y, e1.z * xh.z, e1.w * xh.w);
  }
  if (tid < TM * (kSmTN / 4)) {
    st4(tile + row * kSmTN + nn, e1);
    st4(tile + (TM + row) * kSmTN + nn, e2);
  }
  __syncthreads();
  if (tid < kSmTN && n0 + tid < p.N) {
    const int rows = t1 - t0, col = n0 + tid;
    const int half = col >= e.C ? 1 : 0;                    // N is C or 2C (pixel shuffle x2)
    const int q = (t0 / TM) * (p.N > e.C ? 2 : 1) + half;
    float* dst = e.part + (((long long)b * e.n_part + q) * e.C + (col - half * e.C)) * 2;
    float s1 = 0.f, s2 = 0.f;
    for (int r = 0; r < rows; ++r) s1 += tile[r * kSmTN + tid];
    if (e.mode == 1) {
      const float m = s1 * __frcp_rn((float)rows);
      for (int r = 0; r < rows; ++r) { const float d = tile[r * kSmTN + tid] - m; s2 = fmaf(d, d, s2); }
    } else {
      for (int r = 0; r < rows; ++r) s2 += tile[(TM + r) * kSmTN + tid];
    }
    dst[0] = s1; dst[1] = s2;
  }
}

}  // namespace avc
