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

template <int TM>
__global__ void __launch_bounds__(32 * kSmWarps) conv_small_kernel(const ConvArgs p) {
  constexpr int RM = TM / 4, NT = 32 * kSmWarps;
  extern __shared__ __align__(16) float smem[];
  const int g_lo = p.zsplit ? blockIdx.z : 0;
  const int g_hi = p.zsplit ? blockIdx.z + 1 : p.n_groups;
  float* S = smem;                                                    // windows, then kMaxTaps zero rows
  const int zr = p.win_off[p.n_groups];
  float* Wr = smem + (size_t)(zr + kMaxTaps) * kSRow;                 // [ring][slab_floats]

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

  // ---- weight ring: slab q = (group, tap) in issue order ----------------------------------------
  // thread constants of the copy: 16-byte column chunk wn of K rows wk, wk + 32, ...
  const int wk = tid >> 3, wn = (tid & 7) << 2;
  const bool wok = (n0 + wn) < Ntot;
  int is_g = g_lo, is_tap = 0, is_slot = 0;   // next slab to issue and the ring slot it goes to
  auto issue = [&](int q) {
    if (q < n_slabs) {
      const TapGroup& G = p.g[is_g];
      const float* src = wok ? G.W + ((long long)is_tap * G.wts + wk) * Ntot + n0 + wn : G.W;
      float* dst = Wr + (size_t)is_slot * p.slab_floats + wk * kSmTN + wn;
      constexpr int kRows = NT / 8;          // K rows covered per pass of the CTA
      const long long sstep = wok ? (long long)kRows * Ntot : 0;
      for (int k = wk; k < G.kc; k += kRows) {
        cp_async16(dst, src, wok);
        src += sstep; dst += kRows * kSmTN;
      }
      if (++is_tap == G.n_taps) { is_tap = 0; ++is_g; }
      if (++is_slot == p.ring) is_slot = 0;
    }
    cp_async_commit();
  };
  pdl_launch_dependents();
  for (int q = 0; q < p.ring - 1; ++q) issue(q);
  pdl_wait();   // weights are loop constants; the windows below are the predecessor's output

  // ---- all activation windows -> smem (loads of all groups in flight together) -------------------
  for (int i = tid; i < kMaxTaps * kSRow; i += NT) S[zr * kSRow + i] = 0.f;
  for (int gi = g_lo; gi < g_hi; ++gi) {
    const TapGroup& G = p.g[gi];
    const WinGeom wg = win_geom(p.bwd, p.s, p.T_y, G, t0, t1);
    const int woff = p.zsplit ? 0 : p.win_off[gi];
    if (wg.nrows > (p.zsplit ? zr : p.win_off[gi + 1] - p.win_off[gi])) __trap();   // host sized the window too small
    const int c4n = G.kc >> 2;
    const float* Ab = p.A + (long long)b * p.a_bs + G.a_ch_off;
    const float* Mb = p.Mk ? p.Mk + (long long)b * p.m_bs + G.a_ch_off : nullptr;
    const int drow = NT / c4n, dcol = NT - drow * c4n;   // idx += NT without a division per element
    int row = tid / c4n, col = tid - row * c4n;
    for (int idx = tid; idx < wg.nrows * c4n; idx += NT, row += drow, col += dcol) {
      if (col >= c4n) { col -= c4n; ++row; }
      const int c = col << 2;
      const int r = wg.wlo + row;
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
      if (ok) {
        v = ld4(Ab + (long long)rr * p.a_rs + c);
        if (Mb) v = dact4mul(v, ld4(Mb + (long long)rr * p.m_rs + c), p.slope);
      }
      st4(S + (size_t)(woff + row) * kSRow + c, v);
    }
  }

  // ---- main loop over slabs ------------------------------------------------------------------------
  float acc[RM][4];
#pragma unroll
  for (int i = 0; i < RM; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
  int rb[RM][3];
  bool edge = false;
  int cg = g_lo - 1, ctap = 0, cnt = 0;   // group / tap of the slab being consumed
  int c_slot = 0;                          // ring slot of the slab being consumed
  const bool all_resident = n_slabs <= p.ring - 1;   // the prologue already requested every slab: no ring traffic, no per-slab barrier
  const float* Sg = S;
  int kc = 0;
  for (int q = 0; q < n_slabs; ++q) {
    if (cnt == 0) {   // first slab of the next group: per-thread window rows
      ++cg; ctap = 0;
      const TapGroup& G = p.g[cg];
      cnt = G.n_taps; kc = G.kc;
      const WinGeom wg = win_geom(p.bwd, p.s, p.T_y, G, t0, t1);
      edge = wg.edge;
      Sg = S + (size_t)(p.zsplit ? 0 : p.win_off[cg]) * kSRow;
      const int zrel = zr - (p.zsplit ? 0 : p.win_off[cg]);
#pragma unroll
      for (int i = 0; i < RM; ++i) {
        const int t = t0 + ty * RM + i;
        const bool tv = t < t1;
        rb[i][0] = tv ? t * sv + G.off0 - wg.wlo : zrel;
        rb[i][1] = (tv && t >= wg.lt_lo && t <= wg.lt_hi) ? -t + G.off0 - wg.wlo : zrel;
        rb[i][2] = (tv && t >= wg.rt_lo && t <= wg.rt_hi) ? 2 * (p.T_y - 1) - t + G.off0 - wg.wlo : zrel;
      }
    }
    if (!all_resident) {
      cp_async_wait_dyn(p.ring - 2);
      __syncthreads();                   // slab q landed for everyone; everyone is done with slab q-1 (and, first time, the windows are visible)
      issue(q + p.ring - 1);             // refills the buffer slab q-1 used
    } else if (q == 0) {
      cp_async_wait<0>();                // every slab was requested before the dependency wait: one barrier for the whole loop
      __syncthreads();
    }
    const float* Wsub = Wr + (size_t)c_slot * p.slab_floats;
    if (++c_slot == p.ring) c_slot = 0;
    const int k_lo = warp * kSmKW;
    if (k_lo < kc) {     // kc is a multiple of kSmKW: a warp has all of its K rows or none
      // all shared-memory loads of the slab are issued before the first FMA: with few warps per
      // scheduler the loop is bound by load latency, not by bandwidth
      float4 a[kSmKW / 4][RM], w[kSmKW];
#pragma unroll
      for (int j = 0; j < kSmKW / 4; ++j) {
#pragma unroll
        for (int i = 0; i < RM; ++i) {
          a[j][i] = ld4(Sg + (rb[i][0] + ctap) * kSRow + k_lo + 4 * j);
          if (edge) {
            a[j][i] = f4add(a[j][i], ld4(Sg + (rb[i][1] + ctap) * kSRow + k_lo + 4 * j));
            a[j][i] = f4add(a[j][i], ld4(Sg + (rb[i][2] + ctap) * kSRow + k_lo + 4 * j));
          }
        }
      }
#pragma unroll
      for (int k = 0; k < kSmKW; ++k) w[k] = ld4(Wsub + (k_lo + k) * kSmTN + tx * 4);
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
    ++ctap; --cnt;
  }
  cp_async_wait<0>();
  __syncthreads();                       // all slabs consumed: the ring becomes the reduction buffer

  // ---- fixed-order sum of the 8 K-slices, then the epilogue ----------------------------------------
  float* red = Wr;                       // [kSmWarps][TM][kSmTN]
#pragma unroll
  for (int i = 0; i < RM; ++i)
    st4(red + ((size_t)warp * TM + ty * RM + i) * kSmTN + tx * 4, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
  __syncthreads();
  if (tid >= TM * (kSmTN / 4)) return;
  const int row = tid >> 3, nn = (tid & 7) << 2;
  const int t = t0 + row, n = n0 + nn;
  if (t >= t1 || n >= p.N) return;
  float4 v = ld4(red + (size_t)row * kSmTN + nn);
#pragma unroll
  for (int w = 1; w < kSmWarps; ++w) v = f4add(v, ld4(red + ((size_t)w * TM + row) * kSmTN + nn));
  const int ch = p.zsplit ? blockIdx.z * p.N + n : n;   // output channel
  if (p.bias) v = f4add(v, ld4(p.bias + ch));
  if (p.Om) v = dact4mul(v, ld4(p.Om + (long long)b * p.om_bs + (long long)t * p.om_rs + ch), p.slope);
  if (p.act) v = act4(v, p.slope);
  if (p.Y2) st4(p.Y2 + (long long)b * p.y2_bs + (long long)t * p.y2_rs + ch, v);
  if (p.res.mode != RES_NONE) v = f4add(v, res_load4(p.res, b, t, p.T_y, ch));
  st4(p.Y + (long long)b * p.y_bs + (long long)t * p.y_rs + ch, v);
}

}  // namespace avc
