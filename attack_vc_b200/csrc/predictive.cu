// predictive.cu -- VSMask PredictiveModel (reference models/predictive_model.py:6-110)
// on B200: eval forward, training forward (BatchNorm batch statistics) and the full backward
// (dgrad + wgrad + BatchNorm / PReLU / bias gradients) of   loss = mean(out^2)   (BASELINE config 5).
//
// Layout: NHWC fp32 (the [B,1,F,T] input and output of the reference are NHWC with C = 1 as they are).
// All reductions are two-stage with a fixed order (no float atomics).  Round-1 status: direct
// CUDA-core kernels, correct and measured, not yet tuned (see DESIGN.md "Row P").
#include <cuda_runtime.h>

#include <cmath>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/avc_b200.h"
#include "common.cuh"
#include "host_util.h"
#include "wgrad_tc.cuh"
#include "conv2d_tc.cuh"

using namespace avc;

namespace {

thread_local std::string g_pm_create_error;

// ---- generic 3x3 gather convolution ---------------------------------------------------------------
// y[b,oh,ow,co] = epi( sum_{kh,kw,ci} x[b, ih(oh,kh), iw(ow,kw), ci] * w[kh][kw][ci][co] )
enum PmMode : int {
  PM_REFLECT = 0,   // ih = reflect(oh*sh + kh - 1)                     Conv2d after ReflectionPad2d(1)
  PM_TRANSPOSED = 1,// ih = (oh - kh)/sh when divisible and in range     ConvTranspose2d forward; conv dgrad on padded coords
  PM_PLAIN = 2      // ih = oh*sh + kh                                   ConvTranspose2d dgrad
};
struct PmConv {
  const float* x; int Hi, Wi, Ci;
  const float* w;                 // [9][Ci][Cop], Cop = Co rounded up to 4
  const float* bias;              // [Cop] or nullptr
  const float* scale; const float* shift;   // [Cop] or nullptr: v = v*scale + shift (folded eval BatchNorm)
  const float* dmask;             // optional tensor shaped like y: v *= dmask > 0 ? 1 : mslope (backward through LeakyReLU)
  float mslope;
  float* y; int Ho, Wo, Co, Cop;
  int B, mode, sh, sw;
  float slope; int act;           // 0 none, 1 x>0?x:slope*x, 2 the same then tanh
  // tensor-core path (conv2d_tc.cuh), taken when the caller provides the operand planes:
  const float* xh; const float* xl;     // hi / lo planes of x; for PM_REFLECT of the reflect-padded x [B,Hi+2,Wi+2,Ci]
  const float* wkh; const float* wkl;   // hi / lo planes of the K-major weight image [9][Co][Ci]
  const float* xp; const float* wkp;    // bf16 pair planes of both (wgrad_tc.cuh st_pair4): with them the two cross terms are ONE kind::f16 MMA
  const float* slope_ptr;               // PReLU slope on the device (overrides slope)
};

__device__ __forceinline__ int pm_src(int o, int k, int n_in, int s, int mode) {
  if (mode == PM_REFLECT) {
    int i = o * s + k - 1;
    i = i < 0 ? -i : i;
    if (i >= n_in) i = 2 * (n_in - 1) - i;
    return i;
  }
  if (mode == PM_TRANSPOSED) {
    const int t = o - k;
    if (t < 0 || t % s) return -1;
    const int i = t / s;
    return i < n_in ? i : -1;
  }
  return o * s + k;
}

// The shared epilogue of the gather kernels: bias, folded BatchNorm, act' mask, activation, store of 4 channels of one pixel
__device__ __forceinline__ void pm_conv_store4(const PmConv& p, long long o, int co, const float (&accq)[4], float4 bias, float4 sc, float4 sf) {
  float v[4] = {fmaf(accq[0] + bias.x, sc.x, sf.x), fmaf(accq[1] + bias.y, sc.y, sf.y),
                fmaf(accq[2] + bias.z, sc.z, sf.z), fmaf(accq[3] + bias.w, sc.w, sf.w)};
  const bool vec = (p.Co & 3) == 0 && co + 3 < p.Co;      // one 16-byte store per pixel (the scalar form wrote 4 bytes at a 16-byte stride)
  float m[4] = {1.f, 1.f, 1.f, 1.f};
  if (p.dmask) {
    if (vec) { const float4 d = ld4(p.dmask + o); m[0] = d.x; m[1] = d.y; m[2] = d.z; m[3] = d.w; }
    else for (int j = 0; j < 4; ++j) if (co + j < p.Co) m[j] = p.dmask[o + j];
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float x = v[j];
    if (p.dmask) x *= m[j] > 0.f ? 1.f : p.mslope;
    if (p.act) x = x > 0.f ? x : x * p.slope;
    if (p.act == 2) x = tanhf(x);
    v[j] = x;
  }
  if (vec) st4(p.y + o, make_float4(v[0], v[1], v[2], v[3]));
  else for (int j = 0; j < 4; ++j) if (co + j < p.Co) p.y[o + j] = v[j];
}

// MODE: the gather mode as a compile-time constant for the single-input-channel launches (-1: read p.mode) -- with the mode
// known the source-index function is two or three instructions instead of a three-way branch with an integer division
template <bool CI1, int MODE = -1>
__global__ void __launch_bounds__(256) pm_conv_kernel(const PmConv p) {
  const int mode = MODE >= 0 ? MODE : p.mode;
  const int co = (blockIdx.y * blockDim.x + threadIdx.x) * 4;
  const long long pg = (long long)blockIdx.x * blockDim.y + threadIdx.y;
  const int wg = (p.Wo + 3) / 4;
  const long long npg = (long long)p.B * p.Ho * wg;
  if (pg >= npg || co >= p.Cop) return;
  const int ow0 = (int)(pg % wg) * 4;
  const long long t = pg / wg;
  const int oh = (int)(t % p.Ho), b = (int)(t / p.Ho);
  float acc[4][4];
#pragma unroll
  for (int q = 0; q < 4; ++q) acc[q][0] = acc[q][1] = acc[q][2] = acc[q][3] = 0.f;
  if (CI1) {
    // single input channel (the first down block and the dgrad of the last up block): the 12 source columns, the 3 source
    // rows and the 9 weight vectors are loop constants of the thread -- computed once, then 36 loads and 144 FMAs straight
    // (the generic loop below recomputed the columns per row and fetched the weights inside it: 125 us for a 131 MB output)
    int iw[3][4];
#pragma unroll
    for (int kw = 0; kw < 3; ++kw)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        int i = (ow0 + q < p.Wo) ? pm_src(ow0 + q, kw, p.Wi, p.sw, mode) : -1;
        iw[kw][q] = i >= p.Wi ? -1 : i;
      }
    float4 w4[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) w4[k] = ld4(p.w + (size_t)k * p.Cop + co);
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int ih = pm_src(oh, kh, p.Hi, p.sh, mode);
      if (ih < 0 || ih >= p.Hi) continue;
      const float* xr = p.x + (long long)(b * p.Hi + ih) * p.Wi;
      float a[3][4];
#pragma unroll
      for (int kw = 0; kw < 3; ++kw)
#pragma unroll
        for (int q = 0; q < 4; ++q) a[kw][q] = iw[kw][q] >= 0 ? __ldg(xr + iw[kw][q]) : 0.f;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const float4 w = w4[kh * 3 + kw];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          acc[q][0] = fmaf(a[kw][q], w.x, acc[q][0]); acc[q][1] = fmaf(a[kw][q], w.y, acc[q][1]);
          acc[q][2] = fmaf(a[kw][q], w.z, acc[q][2]); acc[q][3] = fmaf(a[kw][q], w.w, acc[q][3]);
        }
      }
    }
  } else {
    for (int kh = 0; kh < 3; ++kh) {
      const int ih = pm_src(oh, kh, p.Hi, p.sh, mode);
      if (ih < 0 || ih >= p.Hi) continue;
      for (int kw = 0; kw < 3; ++kw) {
        int iw[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          iw[q] = (ow0 + q < p.Wo) ? pm_src(ow0 + q, kw, p.Wi, p.sw, mode) : -1;
          if (iw[q] >= p.Wi) iw[q] = -1;
        }
        const float* wt = p.w + (size_t)(kh * 3 + kw) * p.Ci * p.Cop + co;
        const float* xr = p.x + ((long long)(b * p.Hi + ih) * p.Wi) * p.Ci;
        for (int ci = 0; ci < p.Ci; ci += 4) {
          float4 a[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) a[q] = iw[q] >= 0 ? ld4(xr + (long long)iw[q] * p.Ci + ci) : f4zero();
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float4 w4 = ld4(wt + (size_t)(ci + c) * p.Cop);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float av = c == 0 ? a[q].x : c == 1 ? a[q].y : c == 2 ? a[q].z : a[q].w;
              acc[q][0] = fmaf(av, w4.x, acc[q][0]); acc[q][1] = fmaf(av, w4.y, acc[q][1]);
              acc[q][2] = fmaf(av, w4.z, acc[q][2]); acc[q][3] = fmaf(av, w4.w, acc[q][3]);
            }
          }
        }
      }
    }
  }
  const float4 bias = p.bias ? ld4(p.bias + co) : f4zero();
  const float4 sc = p.scale ? ld4(p.scale + co) : make_float4(1.f, 1.f, 1.f, 1.f);
  const float4 sf = p.shift ? ld4(p.shift + co) : f4zero();
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    if (ow0 + q >= p.Wo) continue;
    const long long o = ((long long)(b * p.Ho + oh) * p.Wo + ow0 + q) * p.Co + co;
    pm_conv_store4(p, o, co, acc[q], bias, sc, sf);
  }
}

// Single OUTPUT channel (the last up block: ConvTranspose2d 32 -> 1, stride 2, + tanh).  Eight lanes share a 2 x 2 block of
// output pixels (2i + a, 2j + c), each lane with four of the input channels: the block reads input pixels (i, j), (i, j-1),
// (i-1, j), (i-1, j-1) once (coalesced 128-byte pixel reads) and every output meets exactly its own taps (kh = a mod 2, ...):
//   out(2i,   2j)   = x[i,j] w00 + x[i,j-1] w02 + x[i-1,j] w20 + x[i-1,j-1] w22      out(2i,   2j+1) = x[i,j] w01 + x[i-1,j] w21
//   out(2i+1, 2j)   = x[i,j] w10 + x[i,j-1] w12                                      out(2i+1, 2j+1) = x[i,j] w11
// folded over the eight lanes by three shuffles in a fixed order.  The generic kernel gave a pixel to one thread: eight
// uncoalesced 16-byte loads per tap, 181 us for a 6 MB output; one thread group per output pixel with generic index
// arithmetic was slower still (593 us: the integer divisions of 12 M threads).
__global__ void __launch_bounds__(256) pm_convT_co1_kernel(const PmConv p) {
  const int lane_c = threadIdx.x & 7, grp = threadIdx.x >> 3;            // 32 groups per CTA
  const int nbi = p.Hi + 1, nbj = p.Wi + 1;                              // blocks: i in [0, Hi], j in [0, Wi]
  const int ci = lane_c * 4;
  float4 w[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    w[k] = f4zero();
    for (int c = ci; c < p.Ci; c += 32) {                                // Ci = 32: one trip
      const float* wt = p.w + ((size_t)k * p.Ci + c) * p.Cop;
      w[k] = make_float4(__ldg(wt), __ldg(wt + p.Cop), __ldg(wt + 2 * p.Cop), __ldg(wt + 3 * p.Cop));
    }
  }
  const float bias = p.bias ? p.bias[0] : 0.f;
  const long long nblk = (long long)p.B * nbi * nbj;
  for (long long q = (long long)blockIdx.x * 32 + grp; q < nblk; q += (long long)gridDim.x * 32) {
    const int j = (int)(q % nbj);
    const long long t = q / nbj;
    const int i = (int)(t % nbi), b = (int)(t / nbi);
    const float* xb = p.x + (long long)b * p.Hi * p.Wi * p.Ci + ci;
    const bool i0 = i < p.Hi, i1 = i > 0, j0 = j < p.Wi, j1 = j > 0;
    const float4 x00 = i0 && j0 ? ld4(xb + ((long long)i * p.Wi + j) * p.Ci) : f4zero();
    const float4 x01 = i0 && j1 ? ld4(xb + ((long long)i * p.Wi + j - 1) * p.Ci) : f4zero();
    const float4 x10 = i1 && j0 ? ld4(xb + ((long long)(i - 1) * p.Wi + j) * p.Ci) : f4zero();
    const float4 x11 = i1 && j1 ? ld4(xb + ((long long)(i - 1) * p.Wi + j - 1) * p.Ci) : f4zero();
    auto dot = [](float4 a, float4 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w))); };
    float o[4];
    o[0] = dot(x00, w[0]) + dot(x01, w[2]) + dot(x10, w[6]) + dot(x11, w[8]);
    o[1] = dot(x00, w[1]) + dot(x10, w[7]);
    o[2] = dot(x00, w[3]) + dot(x01, w[5]);
    o[3] = dot(x00, w[4]);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      o[k] += __shfl_xor_sync(0xffffffffu, o[k], 1);
      o[k] += __shfl_xor_sync(0xffffffffu, o[k], 2);
      o[k] += __shfl_xor_sync(0xffffffffu, o[k], 4);
    }
    if (lane_c < 4) {                                                    // lane k of the group finishes output k of the block
      const int oh = 2 * i + (lane_c >> 1), ow = 2 * j + (lane_c & 1);
      if (oh < p.Ho && ow < p.Wo) {
        float x = (lane_c == 0 ? o[0] : lane_c == 1 ? o[1] : lane_c == 2 ? o[2] : o[3]) + bias;
        if (p.scale) x = fmaf(x, p.scale[0], p.shift[0]);
        const long long e = (long long)(b * p.Ho + oh) * p.Wo + ow;
        if (p.dmask) x *= p.dmask[e] > 0.f ? 1.f : p.mslope;
        if (p.act) x = x > 0.f ? x : x * p.slope;
        if (p.act == 2) x = tanhf(x);
        p.y[e] = x;
      }
    }
  }
}

// ---- the same convolution as a shared-memory tiled implicit GEMM (Ci % 16 == 0, Cop >= 32) ------------------------
// M = output pixels, N = Cop, K = taps x Ci.  64 pixels x 64 channels per CTA, 4 x 4 per thread, K in slabs of 16
// channels of one tap.  For PM_TRANSPOSED the pixels are enumerated by residue class (oh % sh, ow % sw): a class
// meets only the taps with kh = oh (mod sh), kw = ow (mod sw), so a CTA (one class, blockIdx.z) loops over its valid
// taps only -- the direct kernel multiplied through the structural zeros of the stride-2 transposed convs (4x).
constexpr int kPmTK = 16;

// MP x NP outputs per thread, 16 x 16 threads: tile = 16*MP pixels x 16*NP channels (64 x 64, 128 x 64 or 128 x 128)
template <int MP, int NP>
__global__ void __launch_bounds__(256) pm_conv_tiled_kernel(const PmConv p) {
  constexpr int TM = 16 * MP, TN = 16 * NP, LA = TM / 64, LW = TN / 64;   // float4 loads per thread and slab
  __shared__ __align__(16) float sA[kPmTK][TM + 4];
  __shared__ __align__(16) float sW[kPmTK][TN + 4];
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int co0 = blockIdx.y * TN;
  const bool tr = p.mode == PM_TRANSPOSED;
  const int csh = tr ? p.sh : 1, csw = tr ? p.sw : 1;
  const int ph = (int)blockIdx.z / csw, pw = (int)blockIdx.z % csw;      // residue class of this CTA
  const int Hc = (p.Ho - ph + csh - 1) / csh, Wc = (p.Wo - pw + csw - 1) / csw;
  const long long npix = (long long)p.B * Hc * Wc;
  const long long pix0 = (long long)blockIdx.x * TM;
  if (pix0 >= npix) return;
  auto decode = [&](long long pix, int& b, int& oh, int& ow) {
    if (pix >= npix) { b = -1; oh = ow = 0; return; }
    const int wq = (int)(pix % Wc);
    const long long t = pix / Wc;
    ow = wq * csw + pw; oh = (int)(t % Hc) * csh + ph; b = (int)(t / Hc);
  };
  // the pixels this thread LOADS: (tid / 4) + 64 * j; the pixels it ACCUMULATES: MP * ty + q
  int lb[LA], loh[LA], low[LA];
#pragma unroll
  for (int j = 0; j < LA; ++j) decode(pix0 + (tid >> 2) + 64 * j, lb[j], loh[j], low[j]);
  const int lci = (tid & 3) * 4;
  float acc[MP][NP];
#pragma unroll
  for (int q = 0; q < MP; ++q)
#pragma unroll
    for (int r = 0; r < NP; ++r) acc[q][r] = 0.f;
  // valid taps of this class as a bit mask; the K slabs (tap, 16 channels) form one sequence whose next slab is
  // loaded into registers while the current one is multiplied out of shared memory
  unsigned tmask = 0;
  for (int t9 = 0; t9 < 9; ++t9)
    if (!tr || ((t9 / 3) % csh == ph % csh && (t9 % 3) % csw == pw % csw)) tmask |= 1u << t9;   // (oh - kh) % sh == 0 for the whole class
  const int wr = tid >> 4, wc = (tid & 15) * 4;
  long long xoff[LA];
  const float* wt = p.w;
  auto setup_tap = [&](int t9) {
    const int kh = t9 / 3, kw = t9 % 3;
#pragma unroll
    for (int j = 0; j < LA; ++j) {
      xoff[j] = -1;
      if (lb[j] >= 0) {
        const int ih = pm_src(loh[j], kh, p.Hi, p.sh, p.mode), iw = pm_src(low[j], kw, p.Wi, p.sw, p.mode);
        if (ih >= 0 && ih < p.Hi && iw >= 0 && iw < p.Wi) xoff[j] = (((long long)lb[j] * p.Hi + ih) * p.Wi + iw) * p.Ci;
      }
    }
    wt = p.w + (size_t)t9 * p.Ci * p.Cop;
  };
  int ci0 = 0;
  float4 a[LA], w4[LW];
  auto load_slab = [&]() {
#pragma unroll
    for (int j = 0; j < LA; ++j) a[j] = xoff[j] >= 0 ? ld4(p.x + xoff[j] + ci0 + lci) : f4zero();
#pragma unroll
    for (int j = 0; j < LW; ++j) w4[j] = (co0 + wc + 64 * j < p.Cop) ? ld4(wt + (size_t)(ci0 + wr) * p.Cop + co0 + wc + 64 * j) : f4zero();
  };
  bool more = tmask != 0;
  if (more) { setup_tap(__ffs(tmask) - 1); tmask &= tmask - 1; load_slab(); }
  while (more) {
    __syncthreads();
#pragma unroll
    for (int j = 0; j < LA; ++j) {
      const int px = (tid >> 2) + 64 * j;
      sA[lci][px] = a[j].x; sA[lci + 1][px] = a[j].y; sA[lci + 2][px] = a[j].z; sA[lci + 3][px] = a[j].w;
    }
#pragma unroll
    for (int j = 0; j < LW; ++j) st4(&sW[wr][wc + 64 * j], w4[j]);
    __syncthreads();
    ci0 += kPmTK;
    if (ci0 >= p.Ci) {
      ci0 = 0;
      more = tmask != 0;
      if (more) { setup_tap(__ffs(tmask) - 1); tmask &= tmask - 1; }
    }
    if (more) load_slab();
#pragma unroll
    for (int kk = 0; kk < kPmTK; ++kk) {
      float aa[MP], bb[NP];
#pragma unroll
      for (int q = 0; q < MP; q += 4) { const float4 v = ld4(&sA[kk][ty * MP + q]); aa[q] = v.x; aa[q + 1] = v.y; aa[q + 2] = v.z; aa[q + 3] = v.w; }
#pragma unroll
      for (int r = 0; r < NP; r += 4) { const float4 v = ld4(&sW[kk][tx * NP + r]); bb[r] = v.x; bb[r + 1] = v.y; bb[r + 2] = v.z; bb[r + 3] = v.w; }
#pragma unroll
      for (int q = 0; q < MP; ++q)
#pragma unroll
        for (int r = 0; r < NP; ++r) acc[q][r] = fmaf(aa[q], bb[r], acc[q][r]);
    }
  }
#pragma unroll
  for (int r0 = 0; r0 < NP; r0 += 4) {
    const int co = co0 + tx * NP + r0;
    if (co >= p.Cop) continue;
    const float4 bias = p.bias ? ld4(p.bias + co) : f4zero();
    const float4 sc = p.scale ? ld4(p.scale + co) : make_float4(1.f, 1.f, 1.f, 1.f);
    const float4 sf = p.shift ? ld4(p.shift + co) : f4zero();
#pragma unroll
    for (int q = 0; q < MP; ++q) {
      int b, oh, ow;
      decode(pix0 + ty * MP + q, b, oh, ow);
      if (b < 0) continue;
      float v[4] = {fmaf(acc[q][r0] + bias.x, sc.x, sf.x), fmaf(acc[q][r0 + 1] + bias.y, sc.y, sf.y),
                    fmaf(acc[q][r0 + 2] + bias.z, sc.z, sf.z), fmaf(acc[q][r0 + 3] + bias.w, sc.w, sf.w)};
      const long long o = (((long long)b * p.Ho + oh) * p.Wo + ow) * p.Co + co;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (co + j >= p.Co) continue;
        float x = v[j];
        if (p.dmask) x *= p.dmask[o + j] > 0.f ? 1.f : p.mslope;
        if (p.act) x = x > 0.f ? x : x * p.slope;
        if (p.act == 2) x = tanhf(x);
        p.y[o + j] = x;
      }
    }
  }
}

// ---- per-channel reductions over the N = B*H*W pixels of an NHWC tensor --------------------------------
// stage 1: grid (G, 1); thread = (channel float4 lane, row lane); partial[G][4 quantities][C]
//   kind 0 (BatchNorm forward): q0 = sum y, q1 = sum y^2
//   kind 1 (BatchNorm/PReLU backward): with yn = y*scale+shift, gyn = gz * (yn > 0 ? 1 : a), xh = (y-mean)*rstd:
//                                      q0 = sum gyn, q1 = sum gyn*xh, q2 = sum gz*yn*[yn <= 0]  (PReLU slope gradient)
//   kind 2 (bias gradient): q0 = sum g
struct PmReduce {
  const float* y; const float* g; long long N; int C;
  const float* scale; const float* shift; const float* mean; const float* rstd; const float* a;   // a: the PReLU slope (device: the trainer's Adam step updates it)
  int kind;         // 0: sum y, sum y^2   1: the three sums of the BatchNorm + PReLU backward   2: sum g   3: kind 1's elementwise follow-up, fused (below)
  float* partial;   // [G][3][C]
  // kind 3: gy = gamma*rstd * (gyn - sums[0]/N - xh * sums[1]/N) written as gy (optional) and / or as the hi / lo operand planes of
  // the tensor-core dgrad and wgrad (optional), with partial[.][0][c] = sum of gy (the conv bias gradient) -- one pass over y
  // and g instead of three passes (elementwise, bias reduction, operand split) with gy stored and read back twice
  const float* sums; float invN; float* gy; float* hi; float* lo; float* pair;
};
__global__ void __launch_bounds__(256) pm_reduce_kernel(const PmReduce p) {
  extern __shared__ float4 pm_red[];   // [rowlanes][3][C4]
  const int C4 = p.C >> 2;
  const int cl = threadIdx.x % C4, rl = threadIdx.x / C4, RL = blockDim.x / C4;
  float4 q0 = f4zero(), q1 = f4zero(), q2 = f4zero();
  if (rl < RL) {
    const long long per = (p.N + gridDim.x - 1) / gridDim.x;
    const long long lo = (long long)blockIdx.x * per, hi = min(p.N, lo + per);
    float4 sc = f4zero(), sf = f4zero(), mu = f4zero(), rs = f4zero();
    const float pa = (p.kind == 1 || p.kind == 3) ? *p.a : 0.f;
    float4 t0 = f4zero(), t1 = f4zero();
    if (p.kind == 3) { t0 = f4scale(ld4(p.sums + 4 * cl), p.invN); t1 = f4scale(ld4(p.sums + p.C + 4 * cl), p.invN); }
    if (p.kind == 1 || p.kind == 3) { sc = ld4(p.scale + 4 * cl); sf = ld4(p.shift + 4 * cl); mu = ld4(p.mean + 4 * cl); rs = ld4(p.rstd + 4 * cl); }
    for (long long i = lo + rl; i < hi; i += RL) {
      if (p.kind == 0) {
        const float4 v = ld4(p.y + i * p.C + 4 * cl);
        q0 = f4add(q0, v);
        q1.x = fmaf(v.x, v.x, q1.x); q1.y = fmaf(v.y, v.y, q1.y); q1.z = fmaf(v.z, v.z, q1.z); q1.w = fmaf(v.w, v.w, q1.w);
      } else if (p.kind == 1) {
        const float4 v = ld4(p.y + i * p.C + 4 * cl), gz = ld4(p.g + i * p.C + 4 * cl);
        const float vv[4] = {v.x, v.y, v.z, v.w}, gg[4] = {gz.x, gz.y, gz.z, gz.w};
        const float s4[4] = {sc.x, sc.y, sc.z, sc.w}, f4[4] = {sf.x, sf.y, sf.z, sf.w}, m4[4] = {mu.x, mu.y, mu.z, mu.w}, r4[4] = {rs.x, rs.y, rs.z, rs.w};
        float o0[4], o1[4], o2[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float yn = fmaf(vv[j], s4[j], f4[j]);
          const float gyn = gg[j] * (yn > 0.f ? 1.f : pa);
          const float xh = (vv[j] - m4[j]) * r4[j];
          o0[j] = gyn; o1[j] = gyn * xh; o2[j] = yn > 0.f ? 0.f : gg[j] * yn;
        }
        q0 = f4add(q0, make_float4(o0[0], o0[1], o0[2], o0[3]));
        q1 = f4add(q1, make_float4(o1[0], o1[1], o1[2], o1[3]));
        q2 = f4add(q2, make_float4(o2[0], o2[1], o2[2], o2[3]));
      } else if (p.kind == 3) {
        const long long e = i * p.C + 4 * cl;
        const float4 v = ld4(p.y + e), gz = ld4(p.g + e);
        const float vv[4] = {v.x, v.y, v.z, v.w}, gg[4] = {gz.x, gz.y, gz.z, gz.w};
        const float s4[4] = {sc.x, sc.y, sc.z, sc.w}, f4[4] = {sf.x, sf.y, sf.z, sf.w}, m4[4] = {mu.x, mu.y, mu.z, mu.w}, r4[4] = {rs.x, rs.y, rs.z, rs.w};
        const float a0[4] = {t0.x, t0.y, t0.z, t0.w}, a1[4] = {t1.x, t1.y, t1.z, t1.w};
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float yn = fmaf(vv[j], s4[j], f4[j]);
          const float gyn = gg[j] * (yn > 0.f ? 1.f : pa);
          const float xh = (vv[j] - m4[j]) * r4[j];
          o[j] = s4[j] * (gyn - a0[j] - xh * a1[j]);     // scale = gamma*rstd
        }
        const float4 o4 = make_float4(o[0], o[1], o[2], o[3]);
        q0 = f4add(q0, o4);
        if (p.gy) st4(p.gy + e, o4);
        if (p.hi) {
          const float4 oh = make_float4(tf32_hi(o[0]), tf32_hi(o[1]), tf32_hi(o[2]), tf32_hi(o[3]));
          const float4 ol = f4sub(o4, oh);
          st4(p.hi + e, oh);
          st4(p.lo + e, ol);
          if (p.pair) st_pair4(p.pair, e, oh, ol, false);
        }
      } else {
        q0 = f4add(q0, ld4(p.g + i * p.C + 4 * cl));
      }
    }
    pm_red[(rl * 3 + 0) * C4 + cl] = q0; pm_red[(rl * 3 + 1) * C4 + cl] = q1; pm_red[(rl * 3 + 2) * C4 + cl] = q2;
  }
  __syncthreads();
  if (rl == 0) {
    for (int k = 0; k < 3; ++k) {
      float4 s = f4zero();
      for (int r = 0; r < RL; ++r) s = f4add(s, pm_red[(r * 3 + k) * C4 + cl]);
      st4(p.partial + ((size_t)blockIdx.x * 3 + k) * p.C + 4 * cl, s);
    }
  }
}

// stage 2 of the BatchNorm forward statistics: mean / biased variance -> scale, shift, saved mean / rstd,
// and PyTorch's running-statistics update (momentum 0.1, unbiased variance).
// fixed-order sum of n floats `stride` apart by ONE WARP: lane j adds elements j, j+32, ... in fp64, then a xor tree.
// (The stage-2 kernels used to walk up to 1184 partials with one thread: 30-95 us of pure latency each.)
__device__ __forceinline__ double pm_warp_sum(const float* v, int n, size_t stride) {
  double s = 0.0;
  for (int i = threadIdx.x & 31; i < n; i += 32) s += (double)v[(size_t)i * stride];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  return s;
}
// one warp per channel: the G partials of both sums are folded by pm_warp_sum (single-device runs hand the CTA partials of
// pm_reduce_kernel straight in: no separate stage-2 launch)
__global__ void pm_bn_finalize_kernel(const float* partial, int G, int C, long long N, const float* gamma, const float* beta,
                                      float* scale, float* shift, float* mean, float* rstd,
                                      const float* run_mean, const float* run_var, float* new_mean, float* new_var) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (c >= C) return;
  const double s = pm_warp_sum(partial + c, G, (size_t)3 * C), ss = pm_warp_sum(partial + (size_t)C + c, G, (size_t)3 * C);
  if (threadIdx.x & 31) return;
  const double m = s / (double)N;
  double var = ss / (double)N - m * m;
  if (var < 0.0) var = 0.0;
  const float r = (float)(1.0 / sqrt(var + 1e-5));
  mean[c] = (float)m; rstd[c] = r;
  const float sc = gamma[c] * r;
  scale[c] = sc; shift[c] = beta[c] - (float)m * sc;
  if (new_mean) {
    const double unb = N > 1 ? var * (double)N / (double)(N - 1) : var;
    new_mean[c] = (float)(0.9 * run_mean[c] + 0.1 * m);
    new_var[c] = (float)(0.9 * run_var[c] + 0.1 * unb);
  }
}

// generic stage 2: out[k][c] = sum_g partial[g][k][c] (fixed order), k < 3; one warp per output, launch 3*C warps
__global__ void pm_sum_partials_kernel(const float* partial, int G, int C, float* out) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= 3 * C) return;
  const int k = i / C, c = i - k * C;
  const double s = pm_warp_sum(partial + (size_t)k * C + c, G, (size_t)3 * C);
  if ((threadIdx.x & 31) == 0) out[i] = (float)s;
}

// z = prelu(y*scale + shift)
__global__ void pm_affine_prelu_kernel(const float* y, const float* scale, const float* shift, const float* ap, float* z, long long n4, int C) {
  const float a = *ap;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)((i * 4) % C);
    const float4 v = ld4(y + i * 4), sc = ld4(scale + c), sf = ld4(shift + c);
    float4 o = make_float4(fmaf(v.x, sc.x, sf.x), fmaf(v.y, sc.y, sf.y), fmaf(v.z, sc.z, sf.z), fmaf(v.w, sc.w, sf.w));
    st4(z + i * 4, act4(o, a));
  }
}

// The same, written straight into the NEXT layer's tensor-core operand planes: hi / lo (3xTF32) of prelu(y*scale + shift) with
// ReflectionPad2d(pad) materialised -- z itself is never stored (it was written once and read once, by the split pass)
__global__ void __launch_bounds__(256) pm_affine_prelu_split_kernel(const float* __restrict__ y, const float* __restrict__ scale, const float* __restrict__ shift,
                                                                    const float* __restrict__ ap, float* __restrict__ hi, float* __restrict__ lo,
                                                                    float* __restrict__ pair, int B, int H, int W, int C, int pad) {
  const float a = *ap;
  const int C4 = C >> 2, Wp = W + 2 * pad, Hp = H + 2 * pad;
  const long long n4 = (long long)B * Hp * Wp * C4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4) << 2;
    long long t = i / C4;
    const int wp = (int)(t % Wp); t /= Wp;
    const int hp = (int)(t % Hp), b = (int)(t / Hp);
    int w = wp - pad, hh = hp - pad;
    w = w < 0 ? -w : w; if (w >= W) w = 2 * (W - 1) - w;
    hh = hh < 0 ? -hh : hh; if (hh >= H) hh = 2 * (H - 1) - hh;
    const float4 v = ld4(y + (((long long)b * H + hh) * W + w) * C + c), sc = ld4(scale + c), sf = ld4(shift + c);
    const float4 z = act4(make_float4(fmaf(v.x, sc.x, sf.x), fmaf(v.y, sc.y, sf.y), fmaf(v.z, sc.z, sf.z), fmaf(v.w, sc.w, sf.w)), a);
    const float4 zh = make_float4(tf32_hi(z.x), tf32_hi(z.y), tf32_hi(z.z), tf32_hi(z.w));
    const float4 zl = f4sub(z, zh);
    st4(hi + i * 4, zh);
    st4(lo + i * 4, zl);
    if (pair) st_pair4(pair, i * 4, zh, zl, false);
  }
}

// BatchNorm (training) + PReLU backward, elementwise part:
//   gy = gamma*rstd * (gyn - sum_gyn/N - xh * sum_gyn_xh/N),  gyn = gz * prelu'(yn)
__global__ void pm_bn_bwd_kernel(const float* y, const float* gz, const float* scale, const float* shift, const float* mean,
                                 const float* rstd, const float* sums /*[3][C]*/, const float* ap, float invN, float* gy, long long n4, int C) {
  const float a = *ap;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)((i * 4) % C);
    const float4 v = ld4(y + i * 4), g = ld4(gz + i * 4), sc = ld4(scale + c), sf = ld4(shift + c), mu = ld4(mean + c), rs = ld4(rstd + c);
    const float4 s0 = ld4(sums + c), s1 = ld4(sums + C + c);
    const float vv[4] = {v.x, v.y, v.z, v.w}, gg[4] = {g.x, g.y, g.z, g.w}, s4[4] = {sc.x, sc.y, sc.z, sc.w}, f4[4] = {sf.x, sf.y, sf.z, sf.w};
    const float m4[4] = {mu.x, mu.y, mu.z, mu.w}, r4[4] = {rs.x, rs.y, rs.z, rs.w}, a0[4] = {s0.x, s0.y, s0.z, s0.w}, a1[4] = {s1.x, s1.y, s1.z, s1.w};
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float yn = fmaf(vv[j], s4[j], f4[j]);
      const float gyn = gg[j] * (yn > 0.f ? 1.f : a);
      const float xh = (vv[j] - m4[j]) * r4[j];
      o[j] = s4[j] * (gyn - a0[j] * invN - xh * a1[j] * invN);     // scale = gamma*rstd
    }
    st4(gy + i * 4, make_float4(o[0], o[1], o[2], o[3]));
  }
}

// loss = mean(out^2): g_pre = 2*out/N * (1 - out^2) * leaky'(out)   (tanh and LeakyReLU(0.2) of the last block)
// plus per-CTA partial sums of out^2.  out has C = 1, so plain scalars.
__global__ void __launch_bounds__(256) pm_loss_bwd_kernel(const float* out, float* gpre, long long n, float inv_n, float* partial) {
  __shared__ float ws[8];
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float o = out[i];
    acc = fmaf(o, o, acc);
    if (gpre) gpre[i] = 2.f * o * inv_n * (1.f - o * o) * (o > 0.f ? 1.f : 0.2f);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += ws[i];
    partial[blockIdx.x] = s;
  }
}
__global__ void pm_loss_final_kernel(const float* partial, int G, float inv_n, float* loss) {   // one warp
  const double s = pm_warp_sum(partial, G, 1);
  if (threadIdx.x == 0) *loss = (float)(s * inv_n);
}

// sum of all n elements of v: per-CTA partials, then pm_sum_n_kernel
__global__ void __launch_bounds__(256) pm_sum_all_kernel(const float* v, long long n, float* partial) {
  __shared__ float ws[8];
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) acc += v[i];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += ws[i];
    partial[blockIdx.x] = s;
  }
}
__global__ void pm_sum_n_kernel(const float* v, int n, float* out) {   // one warp
  const double s = pm_warp_sum(v, n, 1);
  if (threadIdx.x == 0) *out = (float)s;
}

// fold the gradient w.r.t. the reflect-padded input ([B,H+2,W+2,C]) back onto the input ([B,H,W,C])
__global__ void pm_fold_kernel(const float* gxp, float* gx, int B, int H, int W, int C) {
  const long long n4 = (long long)B * H * W * (C >> 2);
  const int C4 = C >> 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4) * 4;
    long long t = i / C4;
    const int w = (int)(t % W); t /= W;
    const int h = (int)(t % H), b = (int)(t / H);
    float4 s = f4zero();
    // padded row 0 mirrors input row 1, padded row H+1 mirrors input row H-2 (both when H == 3)
    for (int a = 0; a < 3; ++a) {
      int ph;
      if (a == 0) ph = h + 1; else if (a == 1) { if (h != 1) continue; ph = 0; } else { if (h != H - 2) continue; ph = H + 1; }
      for (int e = 0; e < 3; ++e) {
        int pw;
        if (e == 0) pw = w + 1; else if (e == 1) { if (w != 1) continue; pw = 0; } else { if (w != W - 2) continue; pw = W + 1; }
        s = f4add(s, ld4(gxp + ((long long)(b * (H + 2) + ph) * (W + 2) + pw) * C + c));
      }
    }
    st4(gx + i * 4, s);
  }
}
// C = 1 variant (gradient w.r.t. the model input)
__global__ void pm_fold1_kernel(const float* gxp, float* gx, int B, int H, int W) {
  const long long n = (long long)B * H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    long long t = i;
    const int w = (int)(t % W); t /= W;
    const int h = (int)(t % H), b = (int)(t / H);
    float s = 0.f;
    for (int a = 0; a < 3; ++a) {
      int ph;
      if (a == 0) ph = h + 1; else if (a == 1) { if (h != 1) continue; ph = 0; } else { if (h != H - 2) continue; ph = H + 1; }
      for (int e = 0; e < 3; ++e) {
        int pw;
        if (e == 0) pw = w + 1; else if (e == 1) { if (w != 1) continue; pw = 0; } else { if (w != W - 2) continue; pw = W + 1; }
        s += gxp[(long long)(b * (H + 2) + ph) * (W + 2) + pw];
      }
    }
    gx[i] = s;
  }
}

// ---- weight gradient -----------------------------------------------------------------------------------
// gw[tap][ci][co] = sum over base pixels of A[pixA(base,tap)][ci] * G[pixG(base,tap)][co]
//   down conv (amode PM_REFLECT): base = output pixel (oh,ow); A = block input at reflect(oh*sh+kh-1, ...); G = gy at base
//   conv-transpose (amode PM_PLAIN): base = input pixel (ih,iw); A = x at base; G = gpre at (2ih+kh, 2iw+kw)
// grid (S pixel slices, ci tiles of 64, 9 taps x co tiles of 64)
struct PmWgrad {
  const float* A; int Ha, Wa, Ci;
  const float* G; int Hg, Wg, Co;
  int B, Hb, Wb;        // base pixel grid
  int up;               // 0 down conv, 1 conv-transpose
  int sh, sw;
  float* partial;       // [S][9][Ci][Cop]
  int Cop;
};
// 64 x 64 tile of gw[tap] per CTA, the pixel axis (the GEMM K) staged 16 pixels at a time in shared memory
// and split over gridDim.x slices; thread = 4 ci x 4 co.
constexpr int kWgKC = 16, kWgT = 64, kWgPitch = kWgT + 4;
__global__ void __launch_bounds__(256) pm_wgrad_kernel(const PmWgrad p) {
  __shared__ __align__(16) float As[kWgKC][kWgPitch];
  __shared__ __align__(16) float Gs[kWgKC][kWgPitch];
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int n_cot = (p.Cop + kWgT - 1) / kWgT;
  const int cot = blockIdx.z % n_cot, tap = blockIdx.z / n_cot;
  const int ci0 = blockIdx.y * kWgT, co0 = cot * kWgT;
  const int kh = tap / 3, kw = tap % 3;
  const long long N = (long long)p.B * p.Hb * p.Wb;
  const long long per = ((N + gridDim.x - 1) / gridDim.x + kWgKC - 1) / kWgKC * kWgKC;
  const long long lo = (long long)blockIdx.x * per, hi = min(N, lo + per);
  // loader role: pixel lk = tid / 16 of the chunk, channels 4*(tid % 16) .. +3
  const int lk = tid >> 4, lc = (tid & 15) << 2;
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a) acc[a][0] = acc[a][1] = acc[a][2] = acc[a][3] = 0.f;
  float4 av = f4zero(), gv = f4zero();
  auto load_chunk = [&](long long base) {
    const long long i = base + lk;
    av = f4zero(); gv = f4zero();
    if (i < hi) {
      const int wb = (int)(i % p.Wb);
      const long long t = i / p.Wb;
      const int hb = (int)(t % p.Hb), b = (int)(t / p.Hb);
      int ha, wa, hg, wg;
      if (!p.up) {
        ha = pm_src(hb, kh, p.Ha, p.sh, PM_REFLECT); wa = pm_src(wb, kw, p.Wa, p.sw, PM_REFLECT);
        hg = hb; wg = wb;
      } else {
        ha = hb; wa = wb; hg = 2 * hb + kh; wg = 2 * wb + kw;
      }
      const float* ap = p.A + ((long long)(b * p.Ha + ha) * p.Wa + wa) * p.Ci + ci0 + lc;
      const float* gp = p.G + ((long long)(b * p.Hg + hg) * p.Wg + wg) * p.Co + co0 + lc;
      if (ci0 + lc + 3 < p.Ci && (p.Ci & 3) == 0) av = ld4(ap);
      else av = make_float4(ci0 + lc < p.Ci ? ap[0] : 0.f, ci0 + lc + 1 < p.Ci ? ap[1] : 0.f, ci0 + lc + 2 < p.Ci ? ap[2] : 0.f, ci0 + lc + 3 < p.Ci ? ap[3] : 0.f);
      if (co0 + lc + 3 < p.Co && (p.Co & 3) == 0) gv = ld4(gp);
      else gv = make_float4(co0 + lc < p.Co ? gp[0] : 0.f, co0 + lc + 1 < p.Co ? gp[1] : 0.f, co0 + lc + 2 < p.Co ? gp[2] : 0.f, co0 + lc + 3 < p.Co ? gp[3] : 0.f);
    }
  };
  if (lo < hi) load_chunk(lo);
  for (long long base = lo; base < hi; base += kWgKC) {
    __syncthreads();                 // the previous chunk has been consumed
    st4(&As[lk][lc], av);
    st4(&Gs[lk][lc], gv);
    __syncthreads();
    if (base + kWgKC < hi) load_chunk(base + kWgKC);   // the next chunk's loads fly while this one is multiplied
#pragma unroll
    for (int k = 0; k < kWgKC; ++k) {
      const float4 a4 = ld4(&As[k][ty * 4]), g4 = ld4(&Gs[k][tx * 4]);
      const float aa[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        acc[a][0] = fmaf(aa[a], g4.x, acc[a][0]); acc[a][1] = fmaf(aa[a], g4.y, acc[a][1]);
        acc[a][2] = fmaf(aa[a], g4.z, acc[a][2]); acc[a][3] = fmaf(aa[a], g4.w, acc[a][3]);
      }
    }
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int ci = ci0 + ty * 4 + a, co = co0 + tx * 4;
    if (ci < p.Ci && co < p.Cop) st4(p.partial + (((size_t)blockIdx.x * 9 + tap) * p.Ci + ci) * p.Cop + co, make_float4(acc[a][0], acc[a][1], acc[a][2], acc[a][3]));
  }
}
// gw_out (PyTorch layout) = sum_s partial[s]; down: out[co][ci][kh][kw]; up: out[ci][co][kh][kw]
// SL split lanes per output (blockDim = 32 outputs x SL): lane j adds partials j, j + SL, ... in fp64, the SL sums meet in
// shared memory in a fixed order.  With one thread per output the 138 splits of the 32 -> 64 layer were 138 dependent-latency
// loads in a row on 18 k threads (66 us for 10 MB).
template <int SL>
__global__ void __launch_bounds__(256) pm_wgrad_final_kernel(const float* partial, int S, int Ci, int Co, int Cop, int up, float* out) {
  constexpr int OPB = 256 / SL;                      // outputs per CTA
  __shared__ double red[SL][OPB + 1];
  const long long n = 9LL * Ci * Co;
  const int ol = threadIdx.x % OPB, sl = threadIdx.x / OPB;
  for (long long base = (long long)blockIdx.x * OPB; base < n; base += (long long)gridDim.x * OPB) {
    const long long i = base + ol;
    const bool ok = i < n;
    const int co = ok ? (int)(i % Co) : 0;
    const long long t = ok ? i / Co : 0;
    const int ci = (int)(t % Ci), tap = (int)(t / Ci);
    const float* src = partial + ((size_t)tap * Ci + ci) * Cop + co;
    const size_t stride = (size_t)9 * Ci * Cop;
    double s = 0.0;
    if (ok) {
      int k = sl;
      for (; k + 3 * SL < S; k += 4 * SL) {          // four independent loads in flight
        const float a = src[(size_t)k * stride], b = src[(size_t)(k + SL) * stride], c = src[(size_t)(k + 2 * SL) * stride], d = src[(size_t)(k + 3 * SL) * stride];
        s += (double)a; s += (double)b; s += (double)c; s += (double)d;
      }
      for (; k < S; k += SL) s += (double)src[(size_t)k * stride];
    }
    if (SL > 1) {
      red[sl][ol] = s;
      __syncthreads();
      if (sl == 0) {
#pragma unroll
        for (int j = 1; j < SL; ++j) s += red[j][ol];
      }
    }
    if (ok && sl == 0) {
      const long long o = up ? ((long long)ci * Co + co) * 9 + tap : ((long long)co * Ci + ci) * 9 + tap;
      out[o] = (float)s;
    }
    if (SL > 1) __syncthreads();
  }
}
void launch_pm_wgrad_final(const float* partial, int S, int Ci, int Co, int Cop, int up, float* out, int sm_count, cudaStream_t st) {
  const long long n = 9LL * Ci * Co;
  // enough split lanes per output to put ~300 k threads on the machine, no more than the splits there are
  const int sl = (S >= 16 && n * 4 <= 300000) ? 8 : (S >= 8 && n * 2 <= 300000) ? 4 : (S >= 4 && n <= 300000) ? 2 : 1;
  const int opb = 256 / sl;
  const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>((n + opb - 1) / opb, (long long)sm_count * 16));
  if (sl == 8) pm_wgrad_final_kernel<8><<<grid, 256, 0, st>>>(partial, S, Ci, Co, Cop, up, out);
  else if (sl == 4) pm_wgrad_final_kernel<4><<<grid, 256, 0, st>>>(partial, S, Ci, Co, Cop, up, out);
  else if (sl == 2) pm_wgrad_final_kernel<2><<<grid, 256, 0, st>>>(partial, S, Ci, Co, Cop, up, out);
  else pm_wgrad_final_kernel<1><<<grid, 256, 0, st>>>(partial, S, Ci, Co, Cop, up, out);
  CK(cudaGetLastError());
}

// ---- weight gradient of the two single-channel ends of the model ----------------------------------------------------
// First down block (c_in = 1) and last up block (c_out = 1): out[tap][c] = sum_pix V[pix][c] * s[src(pix, tap)] with V a
// C-channel tensor and s a scalar field -- a reduction over ~1M pixels into 9 x C numbers, bound by reading V once.
// (The tiled GEMM kernel spends 2.1 ms on each: one useful row per 64 x 64 tile.)
//   down: base pixel = output (oh, ow): V = gy, s = x[b, reflect(oh*sh + kh - 1), reflect(ow*sw + kw - 1)]
//   up:   base pixel = input (ih, iw):  V = x,  s = g[b, 2 ih + kh, 2 iw + kw]
struct PmWgradC1 {
  const float* V; int C;
  const float* S; int Hs, Ws;
  int B, Hb, Wb, up, sh, sw;
  float* partial;     // [gridDim.x][9][C]
};
__global__ void __launch_bounds__(256) pm_wgrad_c1_kernel(const PmWgradC1 p) {
  __shared__ float4 red[8][9][8];
  const int C4 = p.C >> 2;                       // 8 channel lanes (C = 32)
  const int cl = threadIdx.x % C4, pl = threadIdx.x / C4, PL = blockDim.x / C4;
  const long long N = (long long)p.B * p.Hb * p.Wb;
  float4 acc[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) acc[t] = f4zero();
  const unsigned Np = (unsigned)N;                 // < 2^31 pixels: 32-bit index arithmetic (the 64-bit divisions dominated the loop)
  // four pixels per trip: their 16-byte loads of V and 36 scalar loads of S are all issued before the first FMA (one pixel per
  // trip left 16 KB in flight per SM: 1.0-1.4 TB/s on a 131 MB tensor)
  const unsigned stride = gridDim.x * PL;
  for (unsigned i0 = blockIdx.x * PL + pl; i0 < Np; i0 += 4 * stride) {
    float4 v[4];
    float sv[4][9];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const unsigned i = i0 + u * stride;
      const bool in = i < Np;
      const unsigned ii = in ? i : 0u;
      const unsigned r = ii / (unsigned)p.Wb;
      const int wb = (int)(ii - r * (unsigned)p.Wb);
      const unsigned bq = r / (unsigned)p.Hb;
      const int hb = (int)(r - bq * (unsigned)p.Hb), b = (int)bq;
      v[u] = in ? ld4(p.V + (long long)ii * p.C + 4 * cl) : f4zero();
      const float* sp = p.S + (long long)b * p.Hs * p.Ws;
      int wsv[3];
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) wsv[kw] = p.up ? 2 * wb + kw : pm_src(wb, kw, p.Ws, p.sw, PM_REFLECT);
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const int hs = p.up ? 2 * hb + kh : pm_src(hb, kh, p.Hs, p.sh, PM_REFLECT);
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) sv[u][kh * 3 + kw] = __ldg(sp + hs * p.Ws + wsv[kw]);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        float4& a = acc[t];
        a.x = fmaf(v[u].x, sv[u][t], a.x); a.y = fmaf(v[u].y, sv[u][t], a.y); a.z = fmaf(v[u].z, sv[u][t], a.z); a.w = fmaf(v[u].w, sv[u][t], a.w);
      }
  }
  // lanes of a warp = 4 pixels x 8 channel lanes: fold the pixel lanes, then the 8 warps (fixed order)
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    float4 a = acc[t];
#pragma unroll
    for (int o = 8; o < 32; o <<= 1) {
      a.x += __shfl_xor_sync(0xffffffffu, a.x, o); a.y += __shfl_xor_sync(0xffffffffu, a.y, o);
      a.z += __shfl_xor_sync(0xffffffffu, a.z, o); a.w += __shfl_xor_sync(0xffffffffu, a.w, o);
    }
    if ((threadIdx.x & 31) < 8) red[threadIdx.x >> 5][t][threadIdx.x & 7] = a;
  }
  __syncthreads();
  if (threadIdx.x < 9 * 8) {
    const int t = threadIdx.x / 8, c = threadIdx.x % 8;
    float4 s4 = red[0][t][c];
    for (int w = 1; w < 8; ++w) s4 = f4add(s4, red[w][t][c]);
    st4(p.partial + ((size_t)blockIdx.x * 9 + t) * p.C + 4 * c, s4);
  }
}
// out[c*9 + tap] = sum_g partial[g][tap][c]   (PyTorch layout of both [C,1,3,3] weights)
__global__ void pm_wgrad_c1_final_kernel(const float* partial, int G, int C, float* out) {   // one warp per output
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= 9 * C) return;
  const int tap = i / C, c = i % C;
  const double s = pm_warp_sum(partial + (size_t)tap * C + c, G, (size_t)9 * C);
  if ((threadIdx.x & 31) == 0) out[c * 9 + tap] = (float)s;
}

// ---- model ---------------------------------------------------------------------------------------------
struct DownSpec { int ci, co, sh, sw; };
struct UpSpec { int ci, co; };
const DownSpec kDown[7] = {{1, 32, 1, 2}, {32, 64, 2, 2}, {64, 128, 2, 2}, {128, 256, 2, 2}, {256, 256, 2, 2}, {256, 512, 2, 2}, {512, 512, 2, 2}};
const UpSpec kUp[5] = {{512, 256}, {256, 128}, {128, 64}, {64, 32}, {32, 1}};

int pad4(int c) { return (c + 3) / 4 * 4; }

struct DownW {
  float *w = nullptr, *wt = nullptr;    // [9][ci][cop], transposed [9][co][cip]
  float *w_h = nullptr, *w_l = nullptr, *wt_h = nullptr, *wt_l = nullptr;   // their hi / lo planes (3xTF32 operands of the tensor-core kernels)
  float *w_p = nullptr, *wt_p = nullptr;                                     // and bf16 pair planes (weight order)
  float *bias = nullptr, *gamma = nullptr, *beta = nullptr, *rmean = nullptr, *rvar = nullptr;
  float *escale = nullptr, *eshift = nullptr;   // folded eval BatchNorm (+ conv bias)
  float a = 0.25f;          // host copy for the eval conv epilogue (refreshed from a_dev after training steps)
  float* a_dev = nullptr;   // the PReLU slope on the device
};
struct UpW {
  float *w = nullptr, *wt = nullptr, *bias = nullptr;
  float *w_h = nullptr, *w_l = nullptr, *wt_h = nullptr, *wt_l = nullptr, *w_p = nullptr, *wt_p = nullptr;
};

}  // namespace

// one tensor of model.parameters() inside the flat PyTorch-layout parameter buffer, and the packed images derived from it
struct PmParam {
  std::string name;
  long long off = 0, n = 0;
  int kind = 0;            // 0: vector copied to d0 (bias / BatchNorm weight / BatchNorm bias / PReLU slope), 1: Conv2d weight [co][ci][3][3], 2: ConvTranspose2d weight [ci][co][3][3]
  int ci = 0, co = 0;
  float *d0 = nullptr, *d1 = nullptr;   // kind 1/2: d0 = forward image [9][ci][cop], d1 = transposed image [9][co][cip]
};
struct PmParamDev { long long off, n; int kind, ci, co, cop, cip; float* d0; float* d1; };

struct avc_pm_handle {
  int device = 0, sm_count = 148;
  Arena wmem;
  SlabPool pool;
  bool have_weights = false;
  DownW down[7];
  UpW up[5];
  std::string err;
  long long launches = 0;
  // training state: every parameter in PyTorch layout in ONE buffer (the trainer's Adam step walks it and rewrites the
  // packed images in the same pass)
  std::vector<PmParam> params;
  long long n_params = 0;
  float* P = nullptr;
  PmParamDev* pdev = nullptr;
  bool eval_stale = false;       // escale / eshift / host PReLU slopes are older than the parameters
  bool planes_stale = true;      // the hi / lo weight planes are older than the weight images
  float* c2_part = nullptr;      // partial sums of K-split tensor-core convs (grows on demand, lives in wmem)
  size_t c2_part_floats = 0;
  // data-parallel training: sums that couple the ranks go through the caller's all-reduce
  avc_allreduce_fn ar = nullptr;
  void* ar_ctx = nullptr;
  float* comm = nullptr;
  long long comm_cap = 0;
  int world = 1;
  // step buffers kept between calls of the same kind and shape (Arena::rewind): the same allocation sequence gets the same
  // addresses back without zero-filling 2-3 GB of slabs per step (declared after `pool`: destroyed before it)
  std::unique_ptr<Arena> ws;
  long long ws_key[6] = {0, 0, 0, 0, 0, 0};
};

namespace {

// kind: 0 forward, 1 train step, 2 trainer step; flags: whatever else changes the allocation sequence
Arena& pm_workspace(avc_pm_handle* h, cudaStream_t st, int kind, int B, int H, int W, int flags) {
  static const bool off = getenv("AVC_PM_NO_WS") != nullptr;     // A/B: a fresh zero-filled arena per call
  const long long key[6] = {kind, B, H, W, flags, (long long)(uintptr_t)st};
  if (off || !h->ws || memcmp(key, h->ws_key, sizeof key) != 0) {
    h->ws.reset();
    h->ws.reset(new Arena(&h->pool, st));
    memcpy(h->ws_key, key, sizeof key);
  } else {
    h->ws->rewind();
  }
  return *h->ws;
}

template <class Fn>
int pm_guarded(avc_pm_handle* h, Fn&& fn) {
  try {
    DeviceGuard dg(h ? h->device : -1);
    fn();
    return AVC_OK;
  } catch (const Fail& f) {
    if (h) h->err = f.msg; else g_pm_create_error = f.msg;
    return f.code;
  } catch (const std::exception& e) {
    if (h) h->err = e.what(); else g_pm_create_error = e.what();
    return AVC_ERR_INVALID;
  }
}

bool pm_tc_enabled() {
  static const bool off = getenv("AVC_PM_NO_TC") != nullptr;
  return !off;
}
bool pm_pair_enabled() {
  // Opt-in (AVC_PM_PAIR=1).  Measured on B200, 256 windows: step 4.53 -> 4.36 ms (58.8 k windows/s) with the MMAs of a stage
  // grouped by kind (alternating kinds per k-step: 4.51 -- a change of kind costs ~64 clk); the third operand plane costs 57 us
  // in the producers.  Not the default: tests/test_predictive_gpu.py::test_train_step_gradients[3-64-77] then misses 1e-3
  // (d loss / d x off by 2e-2: the last down block normalises over 3 values per channel there, which amplifies the ~1e-6
  // forward error of the bf16 cross terms; with 3xTF32 the same case is at 2e-5).
  static const bool on = getenv("AVC_PM_PAIR") != nullptr;
  return on;
}
bool pm_tc_layer(int ci, int co) { return pm_tc_enabled() && ci % 32 == 0 && co % 32 == 0; }
// A/B switch per call site (AVC_PM_TC_MASK): 1 forward down, 2 forward up, 4 dgrad of the transposed convs, 8 dgrad of the down convs, 16 wgrad
bool pm_tc_site(int bit) {
  static const int mask = getenv("AVC_PM_TC_MASK") ? atoi(getenv("AVC_PM_TC_MASK")) : 31;
  return (mask & bit) != 0;
}

// the same convolution on the tensor cores (conv2d_tc.cuh); transposed forms: one launch per output residue class
void launch_pm_conv_tc(avc_pm_handle* h, const PmConv& c, cudaStream_t st) {
  C2Args base{};
  base.B = c.B; base.Ci = c.Ci; base.Cop = c.Cop; base.y = c.y; base.Ho = c.Ho; base.Wo = c.Wo; base.Co = c.Co;
  base.bias = c.bias; base.scale = c.scale; base.shift = c.shift; base.dmask = c.dmask; base.mslope = c.mslope;
  base.slope_ptr = c.slope_ptr; base.slope = c.slope; base.act = c.act; base.ksplit = 1;
  struct Cls { C2Args a; WtOperand X; };
  std::vector<Cls> cls;
  if (c.mode != PM_TRANSPOSED) {
    const int pad = c.mode == PM_REFLECT ? 2 : 0;
    C2Args a = base;
    a.Hb = c.Ho; a.Wb = c.Wo; a.a_wmul = c.sw; a.a_hmul = c.sh; a.n_taps = 9;
    for (int t = 0; t < 9; ++t) { a.tap[t] = t; a.a_woff[t] = t % 3; a.a_hoff[t] = t / 3; }
    a.oh_mul = a.ow_mul = 1; a.oh_off = a.ow_off = 0;
    c2_pick_boxes(a);
    cls.push_back({a, WtOperand{c.xh, c.xl, c.Ci, c.Wi + pad, c.Hi + pad, c.B, c.sw, c.sh}});
  } else {
    for (int rh = 0; rh < c.sh; ++rh)
      for (int rw = 0; rw < c.sw; ++rw) {
        C2Args a = base;
        a.Hb = (c.Ho - rh + c.sh - 1) / c.sh; a.Wb = (c.Wo - rw + c.sw - 1) / c.sw;
        if (a.Hb <= 0 || a.Wb <= 0) continue;
        a.a_wmul = a.a_hmul = 1; a.n_taps = 0;
        for (int t = 0; t < 9; ++t) {
          const int kh = t / 3, kw = t % 3;
          if ((kh - rh) % c.sh || (kw - rw) % c.sw) continue;
          a.tap[a.n_taps] = t; a.a_hoff[a.n_taps] = (rh - kh) / c.sh; a.a_woff[a.n_taps] = (rw - kw) / c.sw; ++a.n_taps;
        }
        a.oh_mul = c.sh; a.oh_off = rh; a.ow_mul = c.sw; a.ow_off = rw;
        c2_pick_boxes(a);
        cls.push_back({a, WtOperand{c.xh, c.xl, c.Ci, c.Wi, c.Hi, c.B, 1, 1}});
      }
  }
  // a layer with fewer tiles than half the machine (the deep 2x1 .. 5x4-pixel layers) splits the (tap, K block) stages of
  // every tile over several CTAs; the partial sums meet in c2_finish_kernel (fixed order), which also applies the epilogue
  int tiles = 0, min_stages = 1 << 30;
  for (const Cls& q : cls) { tiles += c2_tiles(q.a); min_stages = std::min(min_stages, c2_stages(q.a)); }
  static const bool no_split = getenv("AVC_PM_NO_KSPLIT") != nullptr;
  int ksplit = 1;
  if (!no_split && 2 * tiles <= h->sm_count) ksplit = std::max(1, std::min({8, h->sm_count / std::max(1, tiles), min_stages}));
  if (ksplit > 1) {
    const size_t n = (size_t)c.B * c.Ho * c.Wo * c.Co;
    if (h->c2_part_floats < n * ksplit) {
      h->c2_part = h->wmem.f(n * ksplit); h->c2_part_floats = n * ksplit;
      CK(cudaDeviceSynchronize());     // the arena's zero fill runs on the legacy stream: it must not overtake kernels of a non-blocking stream
    }
    for (Cls& q : cls) { q.a.ksplit = ksplit; q.a.part = h->c2_part; q.a.part_stride = (long long)n; }
  }
  const bool pair = c.xp && c.wkp && c.Ci % 8 == 0 && pm_pair_enabled();
  for (const Cls& q : cls) {
    if (pair) {
      C2Args a2 = q.a; a2.pair = 1;
      WtOperand X2 = q.X; X2.lo = c.xp;
      launch_conv2d_tc(X2, c.wkh, c.wkp, c.Ci, c.Cop, a2, h->sm_count, st);
    } else {
      launch_conv2d_tc(q.X, c.wkh, c.wkl, c.Ci, c.Cop, q.a, h->sm_count, st);
    }
    h->launches++;
  }
  if (ksplit > 1) {
    launch_c2_finish(cls[0].a, h->sm_count, st);
    h->launches++;
  }
}

void launch_pm_conv(avc_pm_handle* h, const PmConv& c, cudaStream_t st) {
  if (c.xh && c.wkh && pm_tc_layer(c.Ci, c.Cop)) { launch_pm_conv_tc(h, c, st); return; }
  static const bool no_tiled = getenv("AVC_PM_NO_TILED") != nullptr;
  if (!no_tiled && c.Ci % kPmTK == 0 && c.Cop >= 32) {
    const int csh = c.mode == PM_TRANSPOSED ? c.sh : 1, csw = c.mode == PM_TRANSPOSED ? c.sw : 1;
    const long long npix = (long long)c.B * ((c.Ho + csh - 1) / csh) * ((c.Wo + csw - 1) / csw);   // the largest class
    static const int force = getenv("AVC_PM_TILE") ? atoi(getenv("AVC_PM_TILE")) : 0;   // 1: 64x64, 2: 128x64, 3: 128x128
    auto ctas = [&](int TM, int TN) { return ((npix + TM - 1) / TM) * ((c.Cop + TN - 1) / TN) * csh * csw; };
    // the largest tile that still fills the machine about twice
    int tile = c.Cop >= 128 && ctas(128, 128) >= 2LL * h->sm_count ? 3 : ctas(128, 64) >= 2LL * h->sm_count ? 2 : 1;
    if (force) tile = force;
    if (tile == 3) {
      dim3 grid((unsigned)((npix + 127) / 128), (unsigned)((c.Cop + 127) / 128), (unsigned)(csh * csw));
      pm_conv_tiled_kernel<8, 8><<<grid, 256, 0, st>>>(c);
    } else if (tile == 2) {
      dim3 grid((unsigned)((npix + 127) / 128), (unsigned)((c.Cop + 63) / 64), (unsigned)(csh * csw));
      pm_conv_tiled_kernel<8, 4><<<grid, 256, 0, st>>>(c);
    } else {
      dim3 grid((unsigned)((npix + 63) / 64), (unsigned)((c.Cop + 63) / 64), (unsigned)(csh * csw));
      pm_conv_tiled_kernel<4, 4><<<grid, 256, 0, st>>>(c);
    }
    CK(cudaGetLastError());
    h->launches++;
    return;
  }
  static const bool no_co1 = getenv("AVC_PM_NO_CO1") != nullptr;
  if (!no_co1 && c.Co == 1 && c.Ci == 32 && c.mode == PM_TRANSPOSED && c.sh == 2 && c.sw == 2 && c.Ho <= 2 * c.Hi + 1 && c.Wo <= 2 * c.Wi + 1) {
    const long long nblk = (long long)c.B * (c.Hi + 1) * (c.Wi + 1);
    const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>((nblk + 31) / 32, (long long)h->sm_count * 32));
    pm_convT_co1_kernel<<<grid, 256, 0, st>>>(c);
    CK(cudaGetLastError());
    h->launches++;
    return;
  }
  const int lanes = std::min(c.Cop / 4, 32);
  dim3 block(lanes, 256 / lanes);
  const long long npg = (long long)c.B * c.Ho * ((c.Wo + 3) / 4);
  dim3 grid((unsigned)((npg + block.y - 1) / block.y), (unsigned)((c.Cop / 4 + lanes - 1) / lanes));
  if (c.Ci == 1) {
    if (c.mode == PM_REFLECT) pm_conv_kernel<true, PM_REFLECT><<<grid, block, 0, st>>>(c);
    else if (c.mode == PM_PLAIN) pm_conv_kernel<true, PM_PLAIN><<<grid, block, 0, st>>>(c);
    else pm_conv_kernel<true><<<grid, block, 0, st>>>(c);
  }
  else {
    if (c.Ci % 4) fail(AVC_ERR_INVALID, "predictive conv: c_in %d not a multiple of 4", c.Ci);
    pm_conv_kernel<false><<<grid, block, 0, st>>>(c);
  }
  CK(cudaGetLastError());
  h->launches++;
}

// CTAs of a per-channel reduction over N pixels of C channels: about 8 rows per row lane (a CTA has 256 / (C/4) row lanes), at most 4 per SM
int reduce_slices(long long N, int C) {
  const int RL = std::max(1, 256 / std::max(1, C / 4));
  return (int)std::max<long long>(1, std::min<long long>(592, N / (8LL * RL)));
}

// stage 1 + (for kind != 0) stage 2 of a per-channel reduction; returns G
int launch_pm_reduce(avc_pm_handle* h, PmReduce r, cudaStream_t st) {
  const int G = reduce_slices(r.N, r.C);
  const int C4 = r.C / 4, RL = 256 / C4;
  if (r.C % 4 || C4 > 256 || RL < 1) fail(AVC_ERR_INVALID, "predictive reduce: bad channel count %d", r.C);
  const size_t smem = (size_t)RL * 3 * C4 * sizeof(float4);
  pm_reduce_kernel<<<G, RL * C4, smem, st>>>(r);
  CK(cudaGetLastError());
  h->launches++;
  return G;
}

// sum comm[0..n) over the ranks, ordered on `st` (the callback enqueues the collective there)
void pm_allreduce(avc_pm_handle* h, long long n, cudaStream_t st) {
  if (h->world <= 1) return;
  if (!h->ar || n > h->comm_cap) fail(AVC_ERR_STATE, "all-reduce of %lld floats requested but the registered buffer holds %lld", n, h->comm_cap);
  const int rc = h->ar(h->ar_ctx, h->comm, (int64_t)n, (void*)st);
  if (rc != 0) fail(AVC_ERR_STATE, "the caller's all-reduce callback failed (%d)", rc);
}

unsigned ew_grid(long long n, int sm) { return (unsigned)std::max<long long>(1, std::min<long long>((n + 255) / 256, (long long)sm * 8)); }

struct HostSD {
  std::map<std::string, std::vector<float>> t;
  std::map<std::string, std::vector<int64_t>> shape;
  const std::vector<float>& get(const std::string& k, std::initializer_list<int64_t> want) {
    auto it = t.find(k);
    if (it == t.end()) fail(AVC_ERR_WEIGHTS, "missing weight '%s'", k.c_str());
    if (shape[k] != std::vector<int64_t>(want)) fail(AVC_ERR_WEIGHTS, "weight '%s' has an unexpected shape", k.c_str());
    return it->second;
  }
};

struct Planes { float* h = nullptr; float* l = nullptr; float* p = nullptr; };   // hi, lo (wgrad), bf16 pair (convs)


// hi / lo planes of an NHWC tensor for the 3xTF32 tensor-core kernels, optionally with ReflectionPad2d(pad) materialised
Planes pm_split(avc_pm_handle* h, Arena& mem, const float* src, int B, int H, int W, int C, int pad, cudaStream_t st) {
  Planes p;
  const size_t n = (size_t)B * (H + 2 * pad) * (W + 2 * pad) * C;
  p.h = mem.f(n); p.l = mem.f(n);
  if (pm_pair_enabled() && C % 8 == 0) p.p = mem.f(n);
  wt_split_pad_kernel<<<ew_grid((long long)n / 4, h->sm_count), 256, 0, st>>>(src, p.h, p.l, B, H, W, C, pad, pad, pad, p.p, 0);
  CK(cudaGetLastError());
  h->launches++;
  return p;
}

// the weight images change with every optimiser step: their planes follow lazily, once per forward pass
void pm_refresh_weight_planes(avc_pm_handle* h, cudaStream_t st) {
  if (!h->planes_stale || !pm_tc_enabled()) return;
  auto one = [&](const float* src, float*& hi, float*& lo, float*& pr, size_t n) {
    if (!hi) {
      hi = h->wmem.f(n); lo = h->wmem.f(n); pr = h->wmem.f(n);
      CK(cudaDeviceSynchronize());     // the weight arena zero-fills on the legacy stream: it must not overtake kernels of a non-blocking stream
    }
    wt_split_pad_kernel<<<ew_grid((long long)n / 4, h->sm_count), 256, 0, st>>>(src, hi, lo, 1, 1, (int)(n / 4), 4, 0, 0, 0, pr, 1);
    CK(cudaGetLastError());
    h->launches++;
  };
  for (int l = 1; l < 7; ++l) {
    DownW& d = h->down[l];
    const size_t n = (size_t)9 * kDown[l].ci * kDown[l].co;
    one(d.w, d.w_h, d.w_l, d.w_p, n); one(d.wt, d.wt_h, d.wt_l, d.wt_p, n);
  }
  for (int i = 0; i < 4; ++i) {
    UpW& u = h->up[i];
    const size_t n = (size_t)9 * kUp[i].ci * kUp[i].co;
    one(u.w, u.w_h, u.w_l, u.w_p, n); one(u.wt, u.wt_h, u.wt_l, u.wt_p, n);
  }
  h->planes_stale = false;
}

// activations of one forward pass (kept for the backward)
struct PmActs {
  int B = 0, H[8]{}, W[8]{};        // down: spatial size after block l is H[l+1], W[l+1]; H[0], W[0] = input
  int Hu[6]{}, Wu[6]{};             // up: Hu[0] = H[7]
  const float* x = nullptr;
  float* y[7]{};                    // conv outputs (pre-BatchNorm), training only
  float* z[7]{};                    // block outputs
  float *scale[7]{}, *shift[7]{}, *mean[7]{}, *rstd[7]{};
  float* u[5]{};                    // up block outputs (u[4] = model output when out == nullptr)
  // tensor-core path: hi / lo operand planes of every block input (down: reflect-padded [B,H+2,W+2,C]), kept for wgrad
  Planes zin[7]{}, uin[5]{};
};

void pm_shapes(PmActs& A, int B, int H, int W) {
  A.B = B; A.H[0] = H; A.W[0] = W;
  for (int l = 0; l < 7; ++l) {
    if (A.H[l] < 2 || A.W[l] < 2) fail(AVC_ERR_INVALID, "predictive model: %dx%d is too small for ReflectionPad2d(1) at block %d", A.H[l], A.W[l], l);
    A.H[l + 1] = (A.H[l] + 2 - 3) / kDown[l].sh + 1;
    A.W[l + 1] = (A.W[l] + 2 - 3) / kDown[l].sw + 1;
  }
  A.Hu[0] = A.H[7]; A.Wu[0] = A.W[7];
  for (int i = 0; i < 5; ++i) { A.Hu[i + 1] = (A.Hu[i] - 1) * 2 + 3; A.Wu[i + 1] = (A.Wu[i] - 1) * 2 + 3; }
}

// forward; training: batch statistics, conv outputs kept.  new_stats (optional): [7] pairs of device pointers
void pm_forward(avc_pm_handle* h, Arena& mem, PmActs& A, const float* x, float* out, bool training,
                float* const* new_mean, float* const* new_var, cudaStream_t st) {
  pm_refresh_weight_planes(h, st);
  A.x = x;
  const float* in = x;
  static const bool no_fuse = getenv("AVC_PM_NO_FUSE") != nullptr;     // A/B: activation and operand split as two passes
  Planes ready;                                                          // operand planes of `in` made by the producing layer
  for (int l = 0; l < 7; ++l) {
    const DownSpec& s = kDown[l];
    const DownW& w = h->down[l];
    const long long npix = (long long)A.B * A.H[l + 1] * A.W[l + 1];
    A.z[l] = mem.f((size_t)npix * s.co);
    PmConv c{};
    c.x = in; c.Hi = A.H[l]; c.Wi = A.W[l]; c.Ci = s.ci;
    c.w = w.w; c.Ho = A.H[l + 1]; c.Wo = A.W[l + 1]; c.Co = s.co; c.Cop = pad4(s.co);
    c.B = A.B; c.mode = PM_REFLECT; c.sh = s.sh; c.sw = s.sw;
    if (pm_tc_layer(s.ci, s.co) && pm_tc_site(1)) {
      A.zin[l] = ready.h ? ready : pm_split(h, mem, in, A.B, A.H[l], A.W[l], s.ci, 1, st);
      c.xh = A.zin[l].h; c.xl = A.zin[l].l; c.xp = A.zin[l].p; c.wkh = w.wt_h; c.wkl = w.wt_l; c.wkp = w.wt_p;
    }
    if (!training) {
      c.scale = w.escale; c.shift = w.eshift; c.slope = w.a; c.act = 1; c.y = A.z[l];
      launch_pm_conv(h, c, st);
    } else {
      A.y[l] = mem.f((size_t)npix * s.co);
      A.scale[l] = mem.f(s.co); A.shift[l] = mem.f(s.co); A.mean[l] = mem.f(s.co); A.rstd[l] = mem.f(s.co);
      c.bias = w.bias; c.y = A.y[l];
      launch_pm_conv(h, c, st);
      PmReduce r{};
      r.y = A.y[l]; r.N = npix; r.C = s.co; r.kind = 0;
      r.partial = mem.f((size_t)reduce_slices(npix, s.co) * 3 * s.co);
      int G = launch_pm_reduce(h, r, st);
      const float* part = r.partial;
      long long n_stat = npix;
      if (h->world > 1) {
        // one BatchNorm over the GLOBAL batch (a single-device batch of world x B windows): sum, sum of squares across ranks
        pm_sum_partials_kernel<<<(3 * s.co * 32 + 255) / 256, 256, 0, st>>>(r.partial, G, s.co, h->comm);
        CK(cudaGetLastError());
        h->launches++;
        pm_allreduce(h, 3LL * s.co, st);
        part = h->comm; G = 1;
        n_stat = npix * h->world;
      }
      pm_bn_finalize_kernel<<<(s.co * 32 + 255) / 256, 256, 0, st>>>(part, G, s.co, n_stat, w.gamma, w.beta, A.scale[l], A.shift[l], A.mean[l],
                                                                  A.rstd[l], w.rmean, w.rvar, new_mean ? new_mean[l] : nullptr, new_var ? new_var[l] : nullptr);
      CK(cudaGetLastError());
      // the consumer of z[l]: the next down block (reflect-padded planes) or the first up block.  When it runs on the tensor
      // cores (and so does every wgrad, which otherwise reads z), the activation is written as operand planes directly
      const bool next_tc = l < 6 ? pm_tc_layer(kDown[l + 1].ci, kDown[l + 1].co) && pm_tc_site(1) : pm_tc_layer(kUp[0].ci, kUp[0].co) && pm_tc_site(2);
      ready = Planes();
      if (next_tc && pm_tc_site(16) && !no_fuse) {
        const int pad = l < 6 ? 1 : 0;
        const size_t n = (size_t)A.B * (A.H[l + 1] + 2 * pad) * (A.W[l + 1] + 2 * pad) * s.co;
        ready.h = mem.f(n); ready.l = mem.f(n);
        if (pm_pair_enabled()) ready.p = mem.f(n);
        pm_affine_prelu_split_kernel<<<ew_grid((long long)n / 4, h->sm_count), 256, 0, st>>>(A.y[l], A.scale[l], A.shift[l], w.a_dev, ready.h, ready.l, ready.p,
                                                                                              A.B, A.H[l + 1], A.W[l + 1], s.co, pad);
      } else {
        const long long n4 = npix * s.co / 4;
        pm_affine_prelu_kernel<<<ew_grid(n4, h->sm_count), 256, 0, st>>>(A.y[l], A.scale[l], A.shift[l], w.a_dev, A.z[l], n4, s.co);
      }
      CK(cudaGetLastError());
      h->launches += 2;
    }
    in = A.z[l];
  }
  for (int i = 0; i < 5; ++i) {
    const UpSpec& s = kUp[i];
    const UpW& w = h->up[i];
    const long long npix = (long long)A.B * A.Hu[i + 1] * A.Wu[i + 1];
    A.u[i] = (i == 4 && out) ? out : mem.f((size_t)npix * s.co);
    PmConv c{};
    c.x = in; c.Hi = A.Hu[i]; c.Wi = A.Wu[i]; c.Ci = s.ci;
    c.w = w.w; c.bias = w.bias; c.y = A.u[i]; c.Ho = A.Hu[i + 1]; c.Wo = A.Wu[i + 1]; c.Co = s.co; c.Cop = pad4(s.co);
    c.B = A.B; c.mode = PM_TRANSPOSED; c.sh = 2; c.sw = 2; c.slope = 0.2f; c.act = i == 4 ? 2 : 1;
    if (pm_tc_layer(s.ci, s.co) && pm_tc_site(2)) {
      A.uin[i] = (i == 0 && ready.h) ? ready : pm_split(h, mem, in, A.B, A.Hu[i], A.Wu[i], s.ci, 0, st);
      c.xh = A.uin[i].h; c.xl = A.uin[i].l; c.xp = A.uin[i].p; c.wkh = w.wt_h; c.wkl = w.wt_l; c.wkp = w.wt_p;
    }
    launch_pm_conv(h, c, st);
    in = A.u[i];
  }
}

// backward of one training forward pass.  g: gradient w.r.t. the LAST up block's pre-activation [B,Hu5,Wu5] (consumed).
// want(key) -> device buffer of that parameter's gradient (PyTorch shape) or nullptr.
template <class Want>
void pm_backward(avc_pm_handle* h, Arena& mem, PmActs& A, float* g, Want&& want, float* grad_x, cudaStream_t st) {
  const int B = A.B;
  auto wgrad = [&](const float* Ain, int Ha, int Wa, int Ci, const float* G, int Hg, int Wg, int Co, int Hb, int Wb, int up, int sh, int sw, float* dst,
                   Planes ap, Planes gp) {
    if (!dst) return;
    static const bool no_tc = getenv("AVC_PM_NO_TC") != nullptr;
    if (!no_tc && pm_tc_site(16) && Ci % 4 == 0 && Ci >= 32 && Co % 4 == 0 && Co >= 32) {
      // tensor cores: TMA-fed tcgen05 wgrad (wgrad_tc.cuh).  The operands are split hi / lo for 3xTF32 by one elementwise
      // pass each; for a down block that pass also materialises ReflectionPad2d(1), so that a tap is a box offset and the
      // conv stride the tensor map's element stride (for a transposed conv the stride sits on the gradient instead).
      const int pad = up ? 0 : 1, Hap = Ha + 2 * pad, Wap = Wa + 2 * pad;
      if (!ap.h) ap = pm_split(h, mem, Ain, B, Ha, Wa, Ci, pad, st);       // normally made by the forward pass / the dgrad of this block
      if (!gp.h) gp = pm_split(h, mem, G, B, Hg, Wg, Co, 0, st);
      float *ah = ap.h, *al = ap.l, *gh = gp.h, *gl = gp.l;
      WtArgs p{};
      p.a_wmul = up ? 1 : sw; p.a_hmul = up ? 1 : sh; p.g_wmul = up ? 2 : 1; p.g_hmul = up ? 2 : 1;
      for (int t = 0; t < 9; ++t) {
        p.a_woff[t] = up ? 0 : t % 3; p.a_hoff[t] = up ? 0 : t / 3;
        p.g_woff[t] = up ? t % 3 : 0; p.g_hoff[t] = up ? t / 3 : 0;
      }
      p.n_taps = 9; p.Ci = Ci; p.Co = Co; p.Cop = pad4(Co);
      wt_pick_boxes(p, Wb, Hb, B);          // after the taps: the stage geometry depends on which operand moves with the tap
      const int S = wt_splits(p, h->sm_count);
      p.partial = mem.f((size_t)S * 9 * Ci * p.Cop);
      const WtOperand Aop{ah, al, Ci, Wap, Hap, B, up ? 1 : sw, up ? 1 : sh}, Gop{gh, gl, Co, Wg, Hg, B, up ? 2 : 1, up ? 2 : 1};
      launch_wgrad_tc(Aop, Gop, p, S, st);
      launch_pm_wgrad_final(p.partial, S, Ci, Co, p.Cop, up, dst, h->sm_count, st);
      h->launches += 2;
      return;
    }
    if (((up && Co == 1 && Ci == 32) || (!up && Ci == 1 && Co == 32))) {
      PmWgradC1 c{};
      c.V = up ? Ain : G; c.C = 32; c.S = up ? G : Ain; c.Hs = up ? Hg : Ha; c.Ws = up ? Wg : Wa;
      c.B = B; c.Hb = Hb; c.Wb = Wb; c.up = up; c.sh = sh; c.sw = sw;
      const long long N = (long long)B * Hb * Wb;
      const int Gc = (int)std::max<long long>(1, std::min<long long>((long long)h->sm_count * 4, N / (32 * 16)));
      c.partial = mem.f((size_t)Gc * 9 * 32);
      pm_wgrad_c1_kernel<<<Gc, 256, 0, st>>>(c);
      CK(cudaGetLastError());
      pm_wgrad_c1_final_kernel<<<(9 * 32 * 32 + 255) / 256, 256, 0, st>>>(c.partial, Gc, 32, dst);
      CK(cudaGetLastError());
      h->launches += 2;
      return;
    }
    PmWgrad q{};
    q.A = Ain; q.Ha = Ha; q.Wa = Wa; q.Ci = Ci; q.G = G; q.Hg = Hg; q.Wg = Wg; q.Co = Co;
    q.B = B; q.Hb = Hb; q.Wb = Wb; q.up = up; q.sh = sh; q.sw = sw; q.Cop = pad4(Co);
    const long long N = (long long)B * Hb * Wb;
    const int tiles = 9 * ((Ci + kWgT - 1) / kWgT) * ((q.Cop + kWgT - 1) / kWgT);
    static const int oversub = getenv("AVC_PM_WG_CTAS") ? atoi(getenv("AVC_PM_WG_CTAS")) : 8;   // pixel-axis slices: CTAs per SM aimed at
    const int S = (int)std::max<long long>(1, std::min<long long>(std::min<long long>(256, ((long long)oversub * h->sm_count + tiles - 1) / tiles), N / (4 * kWgKC) + 1));
    q.partial = mem.f((size_t)S * 9 * Ci * q.Cop);
    dim3 grid(S, (Ci + kWgT - 1) / kWgT, 9 * ((q.Cop + kWgT - 1) / kWgT));
    pm_wgrad_kernel<<<grid, 256, 0, st>>>(q);
    CK(cudaGetLastError());
    launch_pm_wgrad_final(q.partial, S, Ci, Co, q.Cop, up, dst, h->sm_count, st);
    h->launches += 2;
  };
  auto bias_grad = [&](const float* G, long long N, int C, float* dst) {
    if (!dst) return;
    if (C % 4) {   // single output channel (last block): sum of every element
      if (C != 1) fail(AVC_ERR_INVALID, "bias gradient: unsupported channel count %d", C);
      const int G2 = (int)ew_grid(N, h->sm_count);
      float* part = mem.f(G2);
      pm_sum_all_kernel<<<G2, 256, 0, st>>>(G, N, part);
      CK(cudaGetLastError());
      pm_sum_n_kernel<<<1, 32, 0, st>>>(part, G2, dst);
      CK(cudaGetLastError());
      h->launches += 2;
      return;
    }
    PmReduce r{};
    r.g = G; r.N = N; r.C = C; r.kind = 2; r.partial = mem.f((size_t)reduce_slices(N, C) * 3 * C);
    const int Gs = launch_pm_reduce(h, r, st);
    float* s3 = mem.f(3 * (size_t)C);
    pm_sum_partials_kernel<<<(3 * C * 32 + 255) / 256, 256, 0, st>>>(r.partial, Gs, C, s3);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(dst, s3, C * sizeof(float), cudaMemcpyDeviceToDevice, st));
    h->launches += 1;
  };
  // ---- up blocks, last to first: g = gradient w.r.t. block i's pre-activation ----
  for (int i = 4; i >= 0; --i) {
    const UpSpec& s = kUp[i];
    const std::string p = "up_blocks." + std::to_string(i) + ".conv_transpose.0.";
    const float* xin = i > 0 ? A.u[i - 1] : A.z[6];
    const long long npix = (long long)B * A.Hu[i + 1] * A.Wu[i + 1];
    bias_grad(g, npix, s.co, want(p + "bias"));
    Planes gp;
    if (pm_tc_layer(s.ci, s.co)) gp = pm_split(h, mem, g, B, A.Hu[i + 1], A.Wu[i + 1], s.co, 0, st);
    wgrad(xin, A.Hu[i], A.Wu[i], s.ci, g, A.Hu[i + 1], A.Wu[i + 1], s.co, A.Hu[i], A.Wu[i], 1, 2, 2, want(p + "weight"), A.uin[i], gp);
    // dgrad: gx[ih,iw,ci] = sum_{kh,kw,co} g[2ih+kh, 2iw+kw, co] * W[ci][co][kh][kw]; then through the previous LeakyReLU
    float* gx = mem.f((size_t)B * A.Hu[i] * A.Wu[i] * s.ci);
    PmConv c{};
    c.x = g; c.Hi = A.Hu[i + 1]; c.Wi = A.Wu[i + 1]; c.Ci = s.co;
    c.w = h->up[i].wt; c.y = gx; c.Ho = A.Hu[i]; c.Wo = A.Wu[i]; c.Co = s.ci; c.Cop = pad4(s.ci);
    c.B = B; c.mode = PM_PLAIN; c.sh = 2; c.sw = 2;
    if (gp.h && pm_tc_site(4)) { c.xh = gp.h; c.xl = gp.l; c.xp = gp.p; c.wkh = h->up[i].w_h; c.wkl = h->up[i].w_l; c.wkp = h->up[i].w_p; }
    if (i > 0) { c.dmask = A.u[i - 1]; c.mslope = 0.2f; }
    launch_pm_conv(h, c, st);
    g = gx;
  }
  // ---- down blocks, last to first: g = gradient w.r.t. block l's output z_l ----
  for (int l = 6; l >= 0; --l) {
    const DownSpec& s = kDown[l];
    const DownW& w = h->down[l];
    const std::string p = "down_blocks." + std::to_string(l) + ".conv.";
    const long long npix = (long long)B * A.H[l + 1] * A.W[l + 1];
    PmReduce r{};
    r.y = A.y[l]; r.g = g; r.N = npix; r.C = s.co; r.kind = 1;
    r.scale = A.scale[l]; r.shift = A.shift[l]; r.mean = A.mean[l]; r.rstd = A.rstd[l]; r.a = w.a_dev;
    r.partial = mem.f((size_t)reduce_slices(npix, s.co) * 3 * s.co);
    const int Gs = launch_pm_reduce(h, r, st);
    float* sums = mem.f(3 * (size_t)s.co);
    pm_sum_partials_kernel<<<(3 * s.co * 32 + 255) / 256, 256, 0, st>>>(r.partial, Gs, s.co, sums);
    CK(cudaGetLastError());
    // this rank's sums are its share of the BatchNorm / PReLU parameter gradients; d x needs the sums over the GLOBAL batch
    const float* gsums = sums;
    long long n_stat = npix;
    if (h->world > 1) {
      CK(cudaMemcpyAsync(h->comm, sums, 3 * (size_t)s.co * sizeof(float), cudaMemcpyDeviceToDevice, st));
      pm_allreduce(h, 3LL * s.co, st);
      float* gs = mem.f(3 * (size_t)s.co);
      CK(cudaMemcpyAsync(gs, h->comm, 3 * (size_t)s.co * sizeof(float), cudaMemcpyDeviceToDevice, st));
      gsums = gs; n_stat = npix * h->world;
    }
    if (float* d = want(p + "2.bias")) CK(cudaMemcpyAsync(d, sums, s.co * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (float* d = want(p + "2.weight")) CK(cudaMemcpyAsync(d, sums + s.co, s.co * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (float* d = want(p + "3.weight")) {   // PReLU slope: sum over channels of q2 (fixed order)
      pm_sum_n_kernel<<<1, 32, 0, st>>>(sums + 2 * (size_t)s.co, s.co, d);
      CK(cudaGetLastError());
    }
    static const bool no_fuse = getenv("AVC_PM_NO_FUSE") != nullptr;     // A/B: elementwise pass, bias reduction and operand split as three passes
    const bool tc = pm_tc_layer(s.ci, s.co);
    float* gy = nullptr;
    Planes gp;
    if (!no_fuse) {
      // one pass over y and g: gy as the operand planes of the tensor-core dgrad / wgrad (and as a tensor only where a
      // CUDA-core consumer remains: the single-channel first block, A/B switches), the bias gradient as per-CTA partial sums
      const bool need_gy = !tc || !pm_tc_site(16) || !pm_tc_site(8);
      if (need_gy) gy = mem.f((size_t)npix * s.co);
      if (tc) { gp.h = mem.f((size_t)npix * s.co); gp.l = mem.f((size_t)npix * s.co); if (pm_pair_enabled()) gp.p = mem.f((size_t)npix * s.co); }
      PmReduce f{};
      f.y = A.y[l]; f.g = g; f.N = npix; f.C = s.co; f.kind = 3;
      f.scale = A.scale[l]; f.shift = A.shift[l]; f.mean = A.mean[l]; f.rstd = A.rstd[l]; f.a = w.a_dev;
      f.sums = gsums; f.invN = 1.f / (float)n_stat; f.gy = gy; f.hi = gp.h; f.lo = gp.l; f.pair = gp.p;
      f.partial = mem.f((size_t)reduce_slices(npix, s.co) * 3 * s.co);
      const int Gf = launch_pm_reduce(h, f, st);
      if (float* d = want(p + "1.bias")) {
        float* s3 = mem.f(3 * (size_t)s.co);
        pm_sum_partials_kernel<<<(3 * s.co * 32 + 255) / 256, 256, 0, st>>>(f.partial, Gf, s.co, s3);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(d, s3, s.co * sizeof(float), cudaMemcpyDeviceToDevice, st));
        h->launches++;
      }
    } else {
      gy = mem.f((size_t)npix * s.co);
      const long long n4 = npix * s.co / 4;
      pm_bn_bwd_kernel<<<ew_grid(n4, h->sm_count), 256, 0, st>>>(A.y[l], g, A.scale[l], A.shift[l], A.mean[l], A.rstd[l], gsums, w.a_dev,
                                                              1.f / (float)n_stat, gy, n4, s.co);
      CK(cudaGetLastError());
      h->launches += 1;
      bias_grad(gy, npix, s.co, want(p + "1.bias"));
      if (tc) gp = pm_split(h, mem, gy, B, A.H[l + 1], A.W[l + 1], s.co, 0, st);
    }
    h->launches += 1;
    const float* xin = l > 0 ? A.z[l - 1] : A.x;
    wgrad(xin, A.H[l], A.W[l], s.ci, gy, A.H[l + 1], A.W[l + 1], s.co, A.H[l + 1], A.W[l + 1], 0, s.sh, s.sw, want(p + "1.weight"), A.zin[l], gp);
    if (l == 0 && !grad_x) break;
    // dgrad on the padded coordinates, then fold the reflect padding back
    const int Hp = A.H[l] + 2, Wp = A.W[l] + 2;
    const int cip = pad4(s.ci);
    float* gxp = mem.f((size_t)B * Hp * Wp * cip);
    PmConv c{};
    c.x = gy; c.Hi = A.H[l + 1]; c.Wi = A.W[l + 1]; c.Ci = s.co;
    c.w = w.wt; c.y = gxp; c.Ho = Hp; c.Wo = Wp; c.Co = s.ci; c.Cop = cip;
    c.B = B; c.mode = PM_TRANSPOSED; c.sh = s.sh; c.sw = s.sw;
    if (gp.h && pm_tc_site(8)) { c.xh = gp.h; c.xl = gp.l; c.xp = gp.p; c.wkh = w.w_h; c.wkl = w.w_l; c.wkp = w.w_p; }
    launch_pm_conv(h, c, st);
    if (l > 0) {
      float* gx = mem.f((size_t)B * A.H[l] * A.W[l] * s.ci);
      pm_fold_kernel<<<ew_grid((long long)B * A.H[l] * A.W[l] * s.ci / 4, h->sm_count), 256, 0, st>>>(gxp, gx, B, A.H[l], A.W[l], s.ci);
      CK(cudaGetLastError());
      g = gx;
    } else {
      pm_fold1_kernel<<<ew_grid((long long)B * A.H[0] * A.W[0], h->sm_count), 256, 0, st>>>(gxp, grad_x, B, A.H[0], A.W[0]);
      CK(cudaGetLastError());
    }
    h->launches += 1;
  }
}

// ---- VSMask training step glue (train_predictive.py:95-111, utils/audio.py:77-116) ------------------------------------
struct VsmaskArgs {
  const float* src;        // [B,F,T]   source mel windows (the reference's [B,1,F,T])
  const float* pert;       // [B,Fp,Tp] model output
  int B, F, T, Fp, Tp;
  int fs, fe;              // future_steps, min(fs + Tp, T)
  int lo_end, hi_start;    // int(F*0.3), int(F*0.7)
  float e1, e2, e3;
};
// the value the reference clamps: perturbed_mels - source_mels with perturbed_mels = source (+ prediction on [fs,fe)),
// rounded exactly as the reference rounds it ((s + p) - s in fp32, not p)
__device__ __forceinline__ float vsmask_delta(const VsmaskArgs& a, int b, int f, int t, float s) {
  if (t < a.fs || t >= a.fe) return 0.f;
  const float pm = s + a.pert[((long long)b * a.Fp + f) * a.Tp + (t - a.fs)];
  return pm - s;
}
__device__ __forceinline__ float vsmask_eps(const VsmaskArgs& a, int f) { return f < a.lo_end ? a.e1 : (f < a.hi_start ? a.e2 : a.e3); }

__global__ void vsmask_apply_kernel(const VsmaskArgs a, float* __restrict__ perturbed) {
  const long long n = (long long)a.B * a.F * a.T;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(i % a.T);
    const long long bf = i / a.T;
    const int f = (int)(bf % a.F), b = (int)(bf / a.F);
    const float s = a.src[i];
    const float d = vsmask_delta(a, b, f, t, s), e = vsmask_eps(a, f);
    perturbed[i] = s + fminf(fmaxf(d, -e), e);
  }
}
// d loss / d (last block's pre-activation) from d loss / d perturbed: the crop and the clamp mask (gradient passes where
// -eps <= delta <= eps, torch.clamp's backward), then tanh and LeakyReLU(0.2) of the last up block (predictive_model.py:47,108)
__global__ void vsmask_grad_kernel(const VsmaskArgs a, const float* __restrict__ gmel /*[B,F,T]*/, float* __restrict__ gpre /*[B,Fp,Tp]*/) {
  const long long n = (long long)a.B * a.Fp * a.Tp;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int tp = (int)(i % a.Tp);
    const long long bf = i / a.Tp;
    const int f = (int)(bf % a.Fp), b = (int)(bf / a.Fp);
    const int t = a.fs + tp;
    float g = 0.f;
    if (f < a.F && t < a.fe) {
      const long long j = ((long long)b * a.F + f) * a.T + t;
      const float d = vsmask_delta(a, b, f, t, a.src[j]), e = vsmask_eps(a, f);
      if (d >= -e && d <= e) g = gmel[j];
    }
    const float o = a.pert[i];
    gpre[i] = g * (1.f - o * o) * (o > 0.f ? 1.f : 0.2f);
  }
}

// ---- Adam over every parameter + refresh of the packed images, one pass (torch/optim/adam.py, defaults of :57) --------
struct PmAdamArgs {
  float* P; const float* G; float* M; float* V;
  const PmParamDev* tab; int n_tab; long long n;
  float step_size, bc2s, b1w, b2, b2w, eps;     // lr/(1-b1^t), sqrt(1-b2^t), 1-b1, b2, 1-b2, eps
};
__global__ void __launch_bounds__(256) pm_adam_kernel(const PmAdamArgs a) {
  __shared__ long long offs[64];
  for (int i = threadIdx.x; i < a.n_tab; i += blockDim.x) offs[i] = a.tab[i].off;
  __syncthreads();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (long long)gridDim.x * blockDim.x) {
    const float g = a.G[i];
    float m = a.M[i], v = a.V[i], p = a.P[i];
    m = m + (g - m) * a.b1w;
    v = v * a.b2 + (a.b2w * g) * g;
    const float denom = sqrtf(v) / a.bc2s + a.eps;
    p = p + (-a.step_size * m) / denom;
    a.M[i] = m; a.V[i] = v; a.P[i] = p;
    int lo = 0, hi = a.n_tab - 1;                      // the tensor this element belongs to
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (offs[mid] <= i) lo = mid; else hi = mid - 1; }
    const PmParamDev d = a.tab[lo];
    const long long r = i - d.off;
    if (d.kind == 0) { d.d0[r] = p; continue; }
    const int tap = (int)(r % 9);
    const long long q = r / 9;
    int ci, co;
    if (d.kind == 1) { ci = (int)(q % d.ci); co = (int)(q / d.ci); } else { co = (int)(q % d.co); ci = (int)(q / d.co); }
    d.d0[((long long)tap * d.ci + ci) * d.cop + co] = p;
    d.d1[((long long)tap * d.co + co) * d.cip + ci] = p;
  }
}

// folded eval BatchNorm of every down block from the current parameters and running statistics (one CTA per block)
struct PmRefoldArgs { const float* bias[7]; const float* gamma[7]; const float* beta[7]; const float* rmean[7]; const float* rvar[7]; float* escale[7]; float* eshift[7]; int co[7]; };
__global__ void pm_refold_kernel(const PmRefoldArgs a) {
  const int l = blockIdx.x;
  for (int c = threadIdx.x; c < a.co[l]; c += blockDim.x) {
    const float sc = a.gamma[l][c] / sqrtf(a.rvar[l][c] + 1e-5f);
    a.escale[l][c] = sc;
    a.eshift[l][c] = (a.bias[l][c] - a.rmean[l][c]) * sc + a.beta[l][c];
  }
}

// eval-mode constants follow the parameters lazily: a training step only marks them stale
void pm_refresh_eval(avc_pm_handle* h, cudaStream_t st) {
  if (!h->eval_stale) return;
  PmRefoldArgs r{};
  for (int l = 0; l < 7; ++l) {
    const DownW& w = h->down[l];
    r.bias[l] = w.bias; r.gamma[l] = w.gamma; r.beta[l] = w.beta; r.rmean[l] = w.rmean; r.rvar[l] = w.rvar;
    r.escale[l] = w.escale; r.eshift[l] = w.eshift; r.co[l] = kDown[l].co;
  }
  pm_refold_kernel<<<7, 256, 0, st>>>(r);
  CK(cudaGetLastError());
  h->launches++;
  for (int l = 0; l < 7; ++l) CK(cudaMemcpyAsync(&h->down[l].a, h->down[l].a_dev, sizeof(float), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  h->eval_stale = false;
}

}  // namespace

struct avc_pm_trainer {
  avc_pm_handle* pm = nullptr;
  avc_handle* se = nullptr;
  avc_pm_trainer_args a{};
  avc_session* spk = nullptr;             // speaker-embedding loss gradient service on `se`
  std::unique_ptr<Arena> mem;             // persistent buffers of the trainer
  float *G = nullptr, *M = nullptr, *V = nullptr;     // [n_params] gradient, Adam moments
  float *src = nullptr, *tgt = nullptr, *perturbed = nullptr, *gmel = nullptr;   // [B,F,T]
  float* loss = nullptr;
  int Fp = 0, Tp = 0;
  long long step = 0;
};

extern "C" {

int avc_pm_create(avc_pm_handle** out, int device) {
  if (!out) { g_pm_create_error = "null argument"; return AVC_ERR_INVALID; }
  *out = nullptr;
  return pm_guarded(nullptr, [&] {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) fail(AVC_ERR_CUDA, "no CUDA device available (%s); libavc_b200 has no CPU fallback", cudaGetErrorString(e));
    if (device < 0 || device >= n) fail(AVC_ERR_INVALID, "device %d out of range (%d devices)", device, n);
    DeviceGuard dg(device);
    cudaDeviceProp p{};
    CK(cudaGetDeviceProperties(&p, device));
    if (p.major < 10) fail(AVC_ERR_CUDA, "device %d is sm_%d%d; libavc_b200 is built for sm_100a only", device, p.major, p.minor);
    auto h = std::make_unique<avc_pm_handle>();
    h->device = device; h->sm_count = p.multiProcessorCount;
    CK(cudaFuncSetAttribute(pm_reduce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    wt_init_attributes();
    c2_init_attributes();
    *out = h.release();
  });
}

void avc_pm_destroy(avc_pm_handle* h) {
  if (!h) return;
  DeviceGuard dg(h->device);
  cudaDeviceSynchronize();
  delete h;
}

const char* avc_pm_last_error(const avc_pm_handle* h) { return h ? h->err.c_str() : g_pm_create_error.c_str(); }

int64_t avc_pm_kernel_launches(const avc_pm_handle* h) { return h ? h->launches : -1; }

int avc_pm_load_weights(avc_pm_handle* h, const avc_weight_view* tensors, int32_t n) {
  if (!h) return AVC_ERR_INVALID;
  return pm_guarded(h, [&] {
    if (!tensors || n <= 0) fail(AVC_ERR_INVALID, "no tensors");
    if (h->have_weights) fail(AVC_ERR_STATE, "weights already loaded; create a new handle");
    HostSD sd;
    for (int i = 0; i < n; ++i) {
      const avc_weight_view& v = tensors[i];
      if (!v.name || !v.data || v.ndim < 0 || v.ndim > 4) fail(AVC_ERR_WEIGHTS, "bad weight view %d", i);
      size_t cnt = 1;
      std::vector<int64_t> shp;
      for (int d = 0; d < v.ndim; ++d) { cnt *= (size_t)v.shape[d]; shp.push_back(v.shape[d]); }
      std::vector<float> host(cnt);
      CK(cudaMemcpy(host.data(), v.data, cnt * sizeof(float), cudaMemcpyDeviceToHost));
      sd.t[v.name] = std::move(host);
      sd.shape[v.name] = shp;
    }
    for (int l = 0; l < 7; ++l) {
      const DownSpec& s = kDown[l];
      const std::string p = "down_blocks." + std::to_string(l) + ".conv.";
      const auto& w = sd.get(p + "1.weight", {s.co, s.ci, 3, 3});
      const auto& b = sd.get(p + "1.bias", {s.co});
      const auto& g = sd.get(p + "2.weight", {s.co});
      const auto& be = sd.get(p + "2.bias", {s.co});
      const auto& rm = sd.get(p + "2.running_mean", {s.co});
      const auto& rv = sd.get(p + "2.running_var", {s.co});
      const auto& a = sd.get(p + "3.weight", {1});
      const int cop = pad4(s.co), cip = pad4(s.ci);
      std::vector<float> wp((size_t)9 * s.ci * cop, 0.f), wt((size_t)9 * s.co * cip, 0.f), es(cop, 0.f), ef(cop, 0.f), bp(cop, 0.f), gp(cop, 0.f), bep(cop, 0.f);
      for (int co = 0; co < s.co; ++co)
        for (int ci = 0; ci < s.ci; ++ci)
          for (int t = 0; t < 9; ++t) {
            const float v = w[((size_t)co * s.ci + ci) * 9 + t];
            wp[((size_t)t * s.ci + ci) * cop + co] = v;
            wt[((size_t)t * s.co + co) * cip + ci] = v;
          }
      for (int c = 0; c < s.co; ++c) {
        const float sc = g[c] / std::sqrt(rv[c] + 1e-5f);
        es[c] = sc; ef[c] = (b[c] - rm[c]) * sc + be[c];
        bp[c] = b[c]; gp[c] = g[c]; bep[c] = be[c];
      }
      DownW& d = h->down[l];
      d.w = h->wmem.upload(wp); d.wt = h->wmem.upload(wt); d.bias = h->wmem.upload(bp);
      d.gamma = h->wmem.upload(gp); d.beta = h->wmem.upload(bep);
      std::vector<float> rmp(cop, 0.f), rvp(cop, 1.f);
      std::copy(rm.begin(), rm.end(), rmp.begin()); std::copy(rv.begin(), rv.end(), rvp.begin());
      d.rmean = h->wmem.upload(rmp); d.rvar = h->wmem.upload(rvp);
      d.escale = h->wmem.upload(es); d.eshift = h->wmem.upload(ef);
      d.a = a[0];
      d.a_dev = h->wmem.upload(std::vector<float>(1, a[0]));
    }
    for (int i = 0; i < 5; ++i) {
      const UpSpec& s = kUp[i];
      const std::string p = "up_blocks." + std::to_string(i) + ".conv_transpose.0.";
      const auto& w = sd.get(p + "weight", {s.ci, s.co, 3, 3});
      const auto& b = sd.get(p + "bias", {s.co});
      const int cop = pad4(s.co), cip = pad4(s.ci);
      std::vector<float> wp((size_t)9 * s.ci * cop, 0.f), wt((size_t)9 * s.co * cip, 0.f), bp(cop, 0.f);
      for (int ci = 0; ci < s.ci; ++ci)
        for (int co = 0; co < s.co; ++co)
          for (int t = 0; t < 9; ++t) {
            const float v = w[((size_t)ci * s.co + co) * 9 + t];
            wp[((size_t)t * s.ci + ci) * cop + co] = v;
            wt[((size_t)t * s.co + co) * cip + ci] = v;
          }
      std::copy(b.begin(), b.end(), bp.begin());
      UpW& u = h->up[i];
      u.w = h->wmem.upload(wp); u.wt = h->wmem.upload(wt); u.bias = h->wmem.upload(bp);
    }
    // flat PyTorch-layout copy of every parameter, in state_dict order (trainer / export)
    {
      std::vector<float> flat;
      auto add = [&](const std::string& name, const std::vector<float>& v, int kind, int ci, int co, float* d0, float* d1) {
        PmParam q;
        q.name = name; q.off = (long long)flat.size(); q.n = (long long)v.size(); q.kind = kind; q.ci = ci; q.co = co; q.d0 = d0; q.d1 = d1;
        flat.insert(flat.end(), v.begin(), v.end());
        h->params.push_back(q);
      };
      for (int l = 0; l < 7; ++l) {
        const DownSpec& sp = kDown[l];
        const std::string p = "down_blocks." + std::to_string(l) + ".conv.";
        DownW& d = h->down[l];
        add(p + "1.weight", sd.t[p + "1.weight"], 1, sp.ci, sp.co, d.w, d.wt);
        add(p + "1.bias", sd.t[p + "1.bias"], 0, 0, 0, d.bias, nullptr);
        add(p + "2.weight", sd.t[p + "2.weight"], 0, 0, 0, d.gamma, nullptr);
        add(p + "2.bias", sd.t[p + "2.bias"], 0, 0, 0, d.beta, nullptr);
        add(p + "3.weight", sd.t[p + "3.weight"], 0, 0, 0, d.a_dev, nullptr);
      }
      for (int i = 0; i < 5; ++i) {
        const UpSpec& sp = kUp[i];
        const std::string p = "up_blocks." + std::to_string(i) + ".conv_transpose.0.";
        UpW& u = h->up[i];
        add(p + "weight", sd.t[p + "weight"], 2, sp.ci, sp.co, u.w, u.wt);
        add(p + "bias", sd.t[p + "bias"], 0, 0, 0, u.bias, nullptr);
      }
      h->n_params = (long long)flat.size();
      h->P = h->wmem.upload(flat);
      std::vector<PmParamDev> tab;
      for (const PmParam& q : h->params) tab.push_back(PmParamDev{q.off, q.n, q.kind, q.ci, q.co, pad4(q.co), pad4(q.ci), q.d0, q.d1});
      if (tab.size() > 64) fail(AVC_ERR_INVALID, "parameter table too large");
      h->pdev = h->wmem.raw<PmParamDev>(tab.size());
      CK(cudaMemcpy(h->pdev, tab.data(), tab.size() * sizeof(PmParamDev), cudaMemcpyHostToDevice));
    }
    CK(cudaDeviceSynchronize());
    h->have_weights = true;
  });
}

int avc_pm_out_shape(int32_t H, int32_t W, int32_t* Ho, int32_t* Wo) {
  if (!Ho || !Wo || H < 2 || W < 2) return AVC_ERR_INVALID;
  try {
    PmActs A;
    pm_shapes(A, 1, H, W);
    *Ho = A.Hu[5]; *Wo = A.Wu[5];
    return AVC_OK;
  } catch (...) {
    return AVC_ERR_INVALID;
  }
}

int avc_pm_forward(avc_pm_handle* h, const float* x, float* out, int32_t B, int32_t H, int32_t W, int32_t training, void* stream) {
  if (!h) return AVC_ERR_INVALID;
  return pm_guarded(h, [&] {
    if (!h->have_weights) fail(AVC_ERR_STATE, "avc_pm_load_weights must be called first");
    if (!x || !out || B <= 0) fail(AVC_ERR_INVALID, "bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    Arena& mem = pm_workspace(h, st, 0, B, H, W, (training ? 1 : 0) | (h->world << 8));   // zero fills (first call of a shape) ordered on the caller's stream
    PmActs A;
    pm_shapes(A, B, H, W);
    if (!training) pm_refresh_eval(h, st);
    pm_forward(h, mem, A, x, out, training != 0, nullptr, nullptr, st);
    CK(cudaStreamSynchronize(st));
  });
}

// One training forward/backward of loss = mean(out^2).  grads: views named like the state_dict entries
// (weights, biases, BatchNorm weight/bias, PReLU weight), each `data` a device buffer of the PyTorch shape
// that receives the gradient; unknown names are an error, missing ones are simply not written.
// new_stats: optional views "down_blocks.l.conv.2.running_mean|running_var" receiving the updated statistics.
// With avc_pm_set_allreduce(world > 1): BatchNorm runs over the global batch, the loss is this rank's share of the
// global mean and the gradients are this rank's share (the caller sums them over the ranks).
int avc_pm_train_step(avc_pm_handle* h, const float* x, int32_t B, int32_t H, int32_t W, float* out, float* loss,
                      float* grad_x, const avc_weight_view* grads, int32_t n_grads, void* stream) {
  if (!h) return AVC_ERR_INVALID;
  return pm_guarded(h, [&] {
    if (!h->have_weights) fail(AVC_ERR_STATE, "avc_pm_load_weights must be called first");
    if (!x || B <= 0 || !loss) fail(AVC_ERR_INVALID, "bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    std::map<std::string, float*> gout;
    for (int i = 0; i < n_grads; ++i) {
      if (!grads[i].name || !grads[i].data) fail(AVC_ERR_INVALID, "bad gradient view %d", i);
      gout[grads[i].name] = const_cast<float*>(grads[i].data);
    }
    auto want = [&](const std::string& k) -> float* { auto it = gout.find(k); return it == gout.end() ? nullptr : it->second; };
    Arena& mem = pm_workspace(h, st, 1, B, H, W, (out ? 1 : 0) | (grad_x ? 2 : 0) | (n_grads << 2) | (h->world << 16));
    PmActs A;
    pm_shapes(A, B, H, W);
    float* nm[7]; float* nv[7];
    for (int l = 0; l < 7; ++l) {
      const std::string p = "down_blocks." + std::to_string(l) + ".conv.2.";
      nm[l] = want(p + "running_mean"); nv[l] = want(p + "running_var");
      if ((nm[l] == nullptr) != (nv[l] == nullptr)) fail(AVC_ERR_INVALID, "running_mean and running_var outputs must come together");
    }
    pm_forward(h, mem, A, x, out, true, nm, nv, st);
    const float* o = A.u[4];
    const long long n_out = (long long)B * A.Hu[5] * A.Wu[5];
    const float inv_n = 1.f / (float)(n_out * h->world);
    // ---- loss and gradient w.r.t. the last block's pre-activation ----
    float* g = mem.f((size_t)n_out);
    const int LG = (int)ew_grid(n_out, h->sm_count);
    float* lpart = mem.f(LG);
    pm_loss_bwd_kernel<<<LG, 256, 0, st>>>(o, g, n_out, inv_n, lpart);
    CK(cudaGetLastError());
    pm_loss_final_kernel<<<1, 32, 0, st>>>(lpart, LG, inv_n, loss);
    CK(cudaGetLastError());
    h->launches += 2;
    pm_backward(h, mem, A, g, want, grad_x, st);
    CK(cudaStreamSynchronize(st));
  });
}

int avc_pm_set_allreduce(avc_pm_handle* h, avc_allreduce_fn fn, void* ctx, float* comm, int64_t comm_floats, int32_t world_size) {
  if (!h) return AVC_ERR_INVALID;
  return pm_guarded(h, [&] {
    if (world_size < 1) fail(AVC_ERR_INVALID, "world_size %d", world_size);
    if (world_size > 1 && (!fn || !comm || comm_floats < 3 * 512)) fail(AVC_ERR_INVALID, "a callback and a buffer of at least 1536 floats are needed for world_size > 1");
    h->ar = world_size > 1 ? fn : nullptr; h->ar_ctx = ctx; h->comm = comm; h->comm_cap = comm_floats; h->world = world_size;
  });
}

int64_t avc_pm_param_count(const avc_pm_handle* h) { return h ? h->n_params : -1; }

int avc_pm_export_weights(avc_pm_handle* h, const avc_weight_view* tensors, int32_t n, void* stream) {
  if (!h) return AVC_ERR_INVALID;
  return pm_guarded(h, [&] {
    if (!h->have_weights) fail(AVC_ERR_STATE, "avc_pm_load_weights must be called first");
    if (!tensors || n <= 0) fail(AVC_ERR_INVALID, "no tensors");
    cudaStream_t st = (cudaStream_t)stream;
    std::map<std::string, std::pair<const float*, long long>> src;
    for (const PmParam& q : h->params) src[q.name] = {h->P + q.off, q.n};
    for (int l = 0; l < 7; ++l) {
      const std::string p = "down_blocks." + std::to_string(l) + ".conv.2.";
      src[p + "running_mean"] = {h->down[l].rmean, kDown[l].co};
      src[p + "running_var"] = {h->down[l].rvar, kDown[l].co};
    }
    for (int i = 0; i < n; ++i) {
      const avc_weight_view& v = tensors[i];
      if (!v.name || !v.data) fail(AVC_ERR_INVALID, "bad view %d", i);
      auto it = src.find(v.name);
      if (it == src.end()) fail(AVC_ERR_WEIGHTS, "unknown tensor '%s'", v.name);
      long long cnt = 1;
      for (int d = 0; d < v.ndim; ++d) cnt *= v.shape[d];
      if (cnt != it->second.second) fail(AVC_ERR_WEIGHTS, "tensor '%s' has %lld elements, expected %lld", v.name, cnt, it->second.second);
      CK(cudaMemcpyAsync(const_cast<float*>(v.data), it->second.first, (size_t)cnt * sizeof(float), cudaMemcpyDeviceToDevice, st));
    }
    CK(cudaStreamSynchronize(st));
  });
}

int avc_pm_trainer_begin(avc_pm_handle* pm, avc_handle* se, const avc_pm_trainer_args* a, void* stream, avc_pm_trainer** out) {
  if (!pm || !out) return AVC_ERR_INVALID;
  *out = nullptr;
  return pm_guarded(pm, [&] {
    if (!pm->have_weights) fail(AVC_ERR_STATE, "avc_pm_load_weights must be called first");
    if (!se || !a) fail(AVC_ERR_INVALID, "null argument");
    if (a->B <= 0 || a->F < 2 || a->T < 2 || a->future_steps < 0) fail(AVC_ERR_INVALID, "bad B/F/T/future_steps");
    if (!(a->beta1 >= 0.f && a->beta1 < 1.f && a->beta2 >= 0.f && a->beta2 < 1.f && a->adam_eps > 0.f)) fail(AVC_ERR_INVALID, "bad Adam constants");
    cudaStream_t st = (cudaStream_t)stream;
    std::unique_ptr<avc_pm_trainer> t(new avc_pm_trainer());
    t->pm = pm; t->se = se; t->a = *a;
    PmActs A;
    pm_shapes(A, a->B, a->F, a->T);
    t->Fp = A.Hu[5]; t->Tp = A.Wu[5];
    if (t->Fp < a->F) fail(AVC_ERR_INVALID, "the prediction has %d mel rows, fewer than the %d of the input", t->Fp, a->F);
    if (a->future_steps >= a->T) fail(AVC_ERR_INVALID, "future_steps %d leaves nothing to perturb in %d frames (the reference skips such batches, train_predictive.py:98)", a->future_steps, a->T);
    t->mem.reset(new Arena(&pm->pool, st));
    Arena& m = *t->mem;
    const size_t np = (size_t)pm->n_params, nel = (size_t)a->B * a->F * a->T;
    t->G = m.f(np); t->M = m.f(np); t->V = m.f(np);
    t->src = nullptr; t->tgt = nullptr;
    t->perturbed = m.f(nel); t->gmel = m.f(nel); t->loss = m.f(1);
    t->src = m.f(nel); t->tgt = m.f(nel);
    // speaker-embedding loss service: reads the trainer's three mel buffers every step, writes d loss / d perturbed
    avc_spk_grad_args g{};
    const int64_t cs[3] = {(int64_t)a->F * a->T, (int64_t)a->T, 1};
    g.perturbed = t->perturbed; g.source = t->src; g.target = t->tgt; g.grad_out = t->gmel; g.loss_out = t->loss;
    for (int i = 0; i < 3; ++i) { g.p_stride[i] = cs[i]; g.s_stride[i] = cs[i]; g.t_stride[i] = cs[i]; g.g_stride[i] = cs[i]; }
    g.B = a->B; g.T = a->T; g.T_tgt = a->T; g.lambda = a->lambda; g.inv_norm = a->inv_norm; g.use_graph = 1;
    const int rc = avc_spk_grad_begin(se, &g, stream, &t->spk);
    if (rc != AVC_OK) fail(rc, "speaker encoder: %s", avc_last_error(se));
    *out = t.release();
  });
}

int avc_pm_trainer_step(avc_pm_trainer* t, const float* source, const float* target, float lr, float* loss_out, void* stream) {
  if (!t) return AVC_ERR_INVALID;
  avc_pm_handle* h = t->pm;
  return pm_guarded(h, [&] {
    if (!source || !target || !(lr > 0.f)) fail(AVC_ERR_INVALID, "bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    const avc_pm_trainer_args& a = t->a;
    const size_t nel = (size_t)a.B * a.F * a.T;
    CK(cudaMemcpyAsync(t->src, source, nel * sizeof(float), cudaMemcpyDeviceToDevice, st));
    CK(cudaMemcpyAsync(t->tgt, target, nel * sizeof(float), cudaMemcpyDeviceToDevice, st));
    Arena& mem = pm_workspace(h, st, 2, a.B, a.F, a.T, h->world);
    PmActs A;
    pm_shapes(A, a.B, a.F, a.T);
    // model.train(); predicted_perturbation = model(source_mels): running statistics updated in place (:64,92)
    float* nm[7]; float* nv[7];
    for (int l = 0; l < 7; ++l) { nm[l] = h->down[l].rmean; nv[l] = h->down[l].rvar; }
    pm_forward(h, mem, A, t->src, nullptr, true, nm, nv, st);
    VsmaskArgs v{};
    v.src = t->src; v.pert = A.u[4]; v.B = a.B; v.F = a.F; v.T = a.T; v.Fp = t->Fp; v.Tp = t->Tp;
    v.fs = a.future_steps; v.fe = std::min(a.future_steps + t->Tp, a.T);
    v.lo_end = (int)(a.F * 0.3); v.hi_start = (int)(a.F * 0.7);     // utils/audio.py:96-97
    v.e1 = a.eps1; v.e2 = a.eps2; v.e3 = a.eps3;
    vsmask_apply_kernel<<<ew_grid((long long)nel, h->sm_count), 256, 0, st>>>(v, t->perturbed);
    CK(cudaGetLastError());
    h->launches++;
    // three speaker-encoder forwards, the loss, backward to perturbed_mels (:113-123)
    int rc = avc_spk_grad_step(t->spk, stream);
    if (rc != AVC_OK) fail(rc, "speaker encoder: %s", avc_last_error(t->se));
    if (loss_out) CK(cudaMemcpyAsync(loss_out, t->loss, sizeof(float), cudaMemcpyDeviceToDevice, st));
    const long long n_out = (long long)a.B * t->Fp * t->Tp;
    float* g = mem.f((size_t)n_out);
    vsmask_grad_kernel<<<ew_grid(n_out, h->sm_count), 256, 0, st>>>(v, t->gmel, g);
    CK(cudaGetLastError());
    h->launches++;
    // optimizer.zero_grad(); loss.backward()
    std::map<std::string, float*> gout;
    for (const PmParam& q : h->params) gout[q.name] = t->G + q.off;
    auto want = [&](const std::string& k) -> float* { auto it = gout.find(k); return it == gout.end() ? nullptr : it->second; };
    pm_backward(h, mem, A, g, want, nullptr, st);
    if (h->world > 1) {
      // data parallel: the parameter gradients of the global batch are the sum of the ranks' shares
      if (h->n_params > h->comm_cap) fail(AVC_ERR_STATE, "the all-reduce buffer holds %lld floats, the gradient has %lld", h->comm_cap, h->n_params);
      CK(cudaMemcpyAsync(h->comm, t->G, (size_t)h->n_params * sizeof(float), cudaMemcpyDeviceToDevice, st));
      pm_allreduce(h, h->n_params, st);
      CK(cudaMemcpyAsync(t->G, h->comm, (size_t)h->n_params * sizeof(float), cudaMemcpyDeviceToDevice, st));
    }
    // optimizer.step(): bias corrections in fp64 on the host like torch/optim/adam.py
    t->step += 1;
    PmAdamArgs ad{};
    ad.P = h->P; ad.G = t->G; ad.M = t->M; ad.V = t->V; ad.tab = h->pdev; ad.n_tab = (int)h->params.size(); ad.n = h->n_params;
    ad.step_size = (float)((double)lr / (1.0 - std::pow((double)a.beta1, (double)t->step)));
    ad.bc2s = (float)std::sqrt(1.0 - std::pow((double)a.beta2, (double)t->step));
    ad.b1w = 1.f - a.beta1; ad.b2 = a.beta2; ad.b2w = 1.f - a.beta2; ad.eps = a.adam_eps;
    pm_adam_kernel<<<ew_grid(h->n_params, h->sm_count), 256, 0, st>>>(ad);
    CK(cudaGetLastError());
    h->launches++;
    h->eval_stale = true;
    h->planes_stale = true;
    // the step's activations go back to the pool when `mem` dies: the stream must have consumed them
    CK(cudaStreamSynchronize(st));
  });
}

int avc_pm_trainer_grads(avc_pm_trainer* t, const avc_weight_view* grads, int32_t n, void* stream) {
  if (!t) return AVC_ERR_INVALID;
  avc_pm_handle* h = t->pm;
  return pm_guarded(h, [&] {
    if (!grads || n <= 0) fail(AVC_ERR_INVALID, "no tensors");
    cudaStream_t st = (cudaStream_t)stream;
    for (int i = 0; i < n; ++i) {
      const avc_weight_view& v = grads[i];
      if (!v.name || !v.data) fail(AVC_ERR_INVALID, "bad view %d", i);
      const PmParam* q = nullptr;
      for (const PmParam& c : h->params) if (c.name == v.name) q = &c;
      if (!q) fail(AVC_ERR_WEIGHTS, "unknown parameter '%s'", v.name);
      long long cnt = 1;
      for (int d = 0; d < v.ndim; ++d) cnt *= v.shape[d];
      if (cnt != q->n) fail(AVC_ERR_WEIGHTS, "parameter '%s' has %lld elements, expected %lld", v.name, cnt, q->n);
      CK(cudaMemcpyAsync(const_cast<float*>(v.data), t->G + q->off, (size_t)cnt * sizeof(float), cudaMemcpyDeviceToDevice, st));
    }
    CK(cudaStreamSynchronize(st));
  });
}

int avc_pm_trainer_end(avc_pm_trainer* t) {
  if (!t) return AVC_ERR_INVALID;
  avc_pm_handle* h = t->pm;
  const int rc = pm_guarded(h, [&] {
    CK(cudaDeviceSynchronize());
    if (t->spk) { avc_attack_end(t->spk, nullptr); t->spk = nullptr; }
  });
  delete t;
  return rc;
}

}  // extern "C"
