// common.cuh -- shared device helpers for libavc_b200 (sm_100a).
// Activations are time-major ("channels-last") fp32: element (b, t, c) of a tensor with batch
// stride bs and row stride rs lives at base[b*bs + t*rs + c].  A row stride of 0 broadcasts one
// row over time (used for the pooled-gradient of AdaptiveAvgPool1d, models.py:340).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace avc {

// residual / skip-connection addressing shared by the conv epilogue and the norm kernels
enum ResMode : int {
  RES_NONE = 0,
  RES_SAME = 1,      // r[t]
  RES_POOL = 2,      // avg_pool1d(k=rf, ceil_mode=True) of r            (models.py:205-206, 302-303)
  RES_POOL_BWD = 3,  // backward of RES_POOL: r[t/rf] / count(t/rf)
  RES_UP = 4,        // nearest upsample x rf: r[t/rf]                    (models.py:52-63, 430-431)
  RES_UP_BWD = 5     // backward of RES_UP: sum_{q<rf} r[t*rf+q]
};

struct ResArgs {
  const float* R;
  long long bs;
  int rs;
  int T_r;   // rows of R per utterance
  int mode;
  int rf;    // pooling / upsampling factor
};

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 f4add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 f4scale(float4 a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }
__device__ __forceinline__ float4 f4zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }

// act(x) = x > 0 ? x : slope*x  (ReLU: slope 0; "lrelu": slope 0.01 -- models.py:107-118)
__device__ __forceinline__ float actf(float x, float slope) { return x > 0.f ? x : x * slope; }
// derivative factor recovered from the sign of the activation OUTPUT (or of the pre-activation:
// both have the same sign for slope >= 0)
__device__ __forceinline__ float dactf(float y, float slope) { return y > 0.f ? 1.f : slope; }
__device__ __forceinline__ float4 act4(float4 v, float s) { return make_float4(actf(v.x, s), actf(v.y, s), actf(v.z, s), actf(v.w, s)); }
__device__ __forceinline__ float4 dact4mul(float4 g, float4 y, float s) {
  return make_float4(g.x * dactf(y.x, s), g.y * dactf(y.y, s), g.z * dactf(y.z, s), g.w * dactf(y.w, s));
}

// value the skip path contributes to output row t, channels [c, c+4) of utterance b
__device__ __forceinline__ float4 res_load4(const ResArgs& r, int b, int t, int T_out, int c) {
  const float* base = r.R + (long long)b * r.bs + c;
  switch (r.mode) {
    case RES_SAME:
      return ld4(base + (long long)t * r.rs);
    case RES_POOL: {
      int lo = t * r.rf, hi = min(lo + r.rf, r.T_r);
      float4 s = f4zero();
      for (int q = lo; q < hi; ++q) s = f4add(s, ld4(base + (long long)q * r.rs));
      return f4scale(s, 1.f / (float)(hi - lo));
    }
    case RES_POOL_BWD: {
      int q = t / r.rf;
      int cnt = min(r.rf, T_out - q * r.rf);
      return f4scale(ld4(base + (long long)q * r.rs), 1.f / (float)cnt);
    }
    case RES_UP:
      return ld4(base + (long long)(t / r.rf) * r.rs);
    case RES_UP_BWD: {
      float4 s = f4zero();
      for (int q = 0; q < r.rf; ++q) s = f4add(s, ld4(base + (long long)(t * r.rf + q) * r.rs));
      return s;
    }
    default:
      return f4zero();
  }
}

// Programmatic dependent launch (see launch_k in engine.cu).  launch_dependents: the next kernel of the
// stream may start its prologue now.  wait: block until every predecessor kernel has completed and its
// writes are visible -- must precede the first read of anything a predecessor wrote and the first
// write of anything a predecessor may still read.  Both are no-ops without the launch attribute.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() { pdl_launch_dependents(); pdl_wait(); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace avc
