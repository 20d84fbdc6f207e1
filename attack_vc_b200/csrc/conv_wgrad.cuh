// conv_wgrad.cuh -- gradient of pad_layer + nn.Conv1d (models.py:10-30) w.r.t. weight and bias.
//
//   dW[co][ci][j] = sum_{b,t} dy[b,t,co] * xpad[b, t*stride + j, ci]        xpad = reflect pad (k/2 left, k/2 - (k even) right)
//   db[co]        = sum_{b,t} dy[b,t,co]
//
// The attack loop never needs these (it optimises the input, SURVEY.md §8 "wgrad note"): the reference
// accumulates them into model.*.grad as an unread side effect of loss.backward().  The entry point exists so
// that a caller who does want that side effect -- or trains the network -- has the third conv kernel next to
// forward and dgrad.  Exact fp32 on the CUDA cores, fixed summation order (bit-reproducible): per tap a
// [c_out x rows] x [rows x c_in] GEMM, 64 x 64 output tiles, the row (b,t) axis split over CTAs, partial tiles
// summed by a second kernel in split order.
#pragma once
#include "common.cuh"

namespace avc {

struct WgradArgs {
  const float* x; int T; int c_in;        // [B,T,c_in] time-major
  const float* dy; int To; int c_out;     // [B,To,c_out]
  int B, k, stride, pl;
  int splits, rows_per_split;             // rows = B*To
  float* partial;                         // [splits][k][c_out][c_in]
};

constexpr int kWgTile = 64, kWgRows = 16;

__global__ void __launch_bounds__(256) conv_wgrad_kernel(const WgradArgs p) {
  __shared__ __align__(16) float sD[kWgRows][kWgTile];
  __shared__ __align__(16) float sX[kWgRows][kWgTile];
  const int co0 = blockIdx.x * kWgTile, ci0 = blockIdx.y * kWgTile;
  const int j = blockIdx.z % p.k, sp = blockIdx.z / p.k;
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int lr = tid >> 4, lc = (tid & 15) * 4;          // loader: row of the chunk, first of 4 channels
  const long long rows = (long long)p.B * p.To;
  const long long r0 = (long long)sp * p.rows_per_split, r1 = min(rows, r0 + p.rows_per_split);
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  for (long long rc = r0; rc < r1; rc += kWgRows) {
    const long long r = rc + lr;
    float4 d = f4zero(), v = f4zero();
    if (r < r1) {
      const int b = (int)(r / p.To), t = (int)(r - (long long)b * p.To);
      if (co0 + lc < p.c_out) d = ld4(p.dy + ((long long)b * p.To + t) * p.c_out + co0 + lc);
      int q = t * p.stride + j - p.pl;                    // reflect padding by index (edge not repeated)
      q = q < 0 ? -q : q;
      if (q >= p.T) q = 2 * (p.T - 1) - q;
      if (ci0 + lc < p.c_in) v = ld4(p.x + ((long long)b * p.T + q) * p.c_in + ci0 + lc);
    }
    __syncthreads();
    st4(&sD[lr][lc], d);
    st4(&sX[lr][lc], v);
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kWgRows; ++kk) {
      const float4 a = ld4(&sD[kk][ty * 4]), b = ld4(&sX[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int n = 0; n < 4; ++n) acc[m][n] = fmaf(av[m], bv[n], acc[m][n]);
    }
  }
  float* out = p.partial + (((long long)sp * p.k + j) * p.c_out) * p.c_in;
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    const int co = co0 + ty * 4 + m, ci = ci0 + tx * 4;
    if (co < p.c_out && ci < p.c_in) st4(out + (long long)co * p.c_in + ci, make_float4(acc[m][0], acc[m][1], acc[m][2], acc[m][3]));
  }
}

// dW[co][ci][j] = sum over splits (in order) of partial[s][j][co][ci]
__global__ void conv_wgrad_final_kernel(const float* __restrict__ partial, int splits, int k, int c_out, int c_in, float* __restrict__ dw) {
  const long long n = (long long)k * c_out * c_in;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int q = 0; q < splits; ++q) s += partial[(long long)q * n + i];
    const int ci = (int)(i % c_in);
    const long long t = i / c_in;
    const int co = (int)(t % c_out), j = (int)(t / c_out);
    dw[((long long)co * c_in + ci) * k + j] = s;
  }
}

// db[co] = sum over all rows of dy[r][co]: one CTA per 32 channels, 8 row lanes, fixed-order tree
__global__ void __launch_bounds__(256) conv_bgrad_kernel(const float* __restrict__ dy, long long rows, int c_out, float* __restrict__ db) {
  __shared__ float red[8][32];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31), rl = threadIdx.x >> 5;
  float s = 0.f;
  if (c < c_out)
    for (long long r = rl; r < rows; r += 8) s += dy[r * c_out + c];
  red[rl][threadIdx.x & 31] = s;
  __syncthreads();
  if (rl == 0 && c < c_out) {
    float t = red[0][threadIdx.x];
    for (int q = 1; q < 8; ++q) t += red[q][threadIdx.x];
    db[c] = t;
  }
}

}  // namespace avc
