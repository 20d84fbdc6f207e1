// tcgen05.mma cadence on one SM as a function of the instruction shape, the operand layout, the A source and
// concurrent shared-memory traffic.  Decides the conv_tc redesign (DESIGN.md "Tensor-core path"):
//   - is N = 256 (rows as the N operand, weights as A) cheaper per FLOP than N = 128?
//   - does the 128-byte swizzled K-major layout change the operand fetch cost?
//   - what does A-from-TMEM buy?
//   - how much do the loaders' STS and the weight ring's bulk copies slow the MMA stream down?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_shapes mma_shapes.cu ; run: ./mma_shapes [grid]
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../attack_vc_b200/csrc/conv_tc.cuh"
using namespace avc;

enum { KIND_TF32 = 0, KIND_BF16 = 1 };
enum { LAY_NOSW = 0, LAY_SW128 = 1 };
enum { BG_NONE = 0, BG_STS = 1, BG_BULK = 2, BG_BOTH = 3, BG_LDS = 4 };

struct Variant { int kind, N, layout, a_tmem, bg, rot; const char* name; int pattern = 0, busy = 0, ld = 0, data = 0; };
// pattern 1: per stage 4 tf32 + 4 bf16 MMAs (the conv kernel's order); 2: the same + tcgen05.commit per stage;
// 3: pattern 2 + a second commit every third stage.  busy: ALU-spinning warps 4, 8, 12 on the issuer's scheduler.
// ld: warps 8-11 stream tcgen05.ld of TMEM columns [256, 512).  data: non-zero operands.

__device__ __forceinline__ void mma_ts_tf32(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
               ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc) : "memory");
}
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

constexpr int kOpBytes = 64 * 1024;   // per operand region
constexpr int kBgBytes = 48 * 1024;
constexpr int kNMma = 480;

__global__ void __launch_bounds__(512, 1) k(Variant v, const float* __restrict__ gsrc, long long* out) {
  extern __shared__ __align__(1024) unsigned char sm[];
  unsigned char* Aop = sm;
  unsigned char* Bop = sm + kOpBytes;
  unsigned char* Bg = sm + 2 * kOpBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + 2 * kOpBytes + kBgBytes);
  volatile int* stop = reinterpret_cast<volatile int*>(bars + 10);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 11);
  const uint32_t b0 = smem_u32(bars);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (2 * kOpBytes + kBgBytes) / 4; i += blockDim.x)
    reinterpret_cast<float*>(sm)[i] = v.data ? 1e-3f * (float)((i * 2654435761u) >> 20) - 2.f : 0.f;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(b0 + 8 * i, 1);
    *stop = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = *slot;
  long long bg_count = 0;
  if (warp == 0) {
    if (lane == 0) {
      const uint32_t N = (uint32_t)v.N;
      const uint32_t fmt = v.kind == KIND_TF32 ? 2u : 1u;
      const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t a = smem_u32(Aop), b = smem_u32(Bop);
      // no-swizzle K-major: [k/4 chunk][row][16 B]; LBO = rows*16 (distance between the two K chunks of one MMA), SBO = 128
      // rotation: 4 k-steps x 5 taps like one K block of a k=5 conv (A window of 136 rows: LBO 136*16)
      long long best = 1LL << 60;
      int ph = 0;
      // descriptors of the 20 (tap, k-step) positions are compile-time offsets from two bases: the issue loop is
      // nothing but MMAs (the first version of this bench rebuilt descriptors per MMA and measured its own ALU work)
      const bool sw = v.layout != LAY_NOSW;
      const uint64_t da0 = sw ? desc_sw128(a) : tc_desc(a, 136 * 16, 128);
      const uint64_t db0 = sw ? desc_sw128(b) : tc_desc(b, N * 16, 128);
      const uint64_t a_ks = sw ? 2 : (uint64_t)((2 * 136 * 16) >> 4), b_ks = sw ? 2 : (uint64_t)((2 * N * 16) >> 4);
      const uint64_t a_tap = (sw || !v.rot) ? 0 : 1;
      const uint64_t a_k1 = v.rot ? a_ks : 0, b_k1 = v.rot ? b_ks : 0;
      for (int rep = 0; rep < 4; ++rep) {
        const long long t0 = clock64();
        if (v.pattern) {
          const uint32_t idesc16 = (1u << 4) | (1u << 7) | (1u << 10) | ((N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
          const uint64_t lo = (uint64_t)(32768 >> 4);      // second operand plane 32 KB further
          for (int st = 0; st < kNMma / 8; ++st) {
            const uint64_t tap = (uint64_t)(st % 5);
            if (v.pattern >= 4) { mbar_wait(b0 + 8 * 6, 1); if (v.pattern >= 5) mbar_wait(b0 + 8 * 7, 1); tc_fence_after(); }
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) tc_mma_tf32(tm, da0 + tap + ks * a_k1, db0 + ks * b_k1, idesc, 1);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) tc_mma_bf16(tm, da0 + lo + tap + ks * a_k1, db0 + lo + ks * b_k1, idesc16);
            if (v.pattern >= 2) tc_commit(b0 + 8 * 4);
            if (v.pattern >= 3 && st % 3 == 2) tc_commit(b0 + 8 * 5);
          }
        } else if (v.a_tmem) {
          for (int i = 0; i < kNMma / 20; ++i) {
#pragma unroll
            for (int tap = 0; tap < 5; ++tap)
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) mma_ts_tf32(tm, tm + 256 + (uint32_t)ks * 8, db0 + ks * b_k1, idesc);
          }
        } else if (v.kind == KIND_TF32) {
          for (int i = 0; i < kNMma / 20; ++i) {
#pragma unroll
            for (int tap = 0; tap < 5; ++tap)
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) tc_mma_tf32(tm, da0 + tap * a_tap + ks * a_k1, db0 + ks * b_k1, idesc, 1);
          }
        } else {
          for (int i = 0; i < kNMma / 20; ++i) {
#pragma unroll
            for (int tap = 0; tap < 5; ++tap)
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) tc_mma_bf16(tm, da0 + tap * a_tap + ks * a_k1, db0 + ks * b_k1, idesc);
          }
        }
        tc_commit(b0); mbar_wait(b0, ph); ph ^= 1;
        const long long t1 = clock64();
        if (t1 - t0 < best) best = t1 - t0;
      }
      out[0] = best;
      if (v.pattern == 0 && !v.a_tmem && v.kind == KIND_TF32 && v.bg == BG_NONE && v.rot == 0) {
        for (int kq = 1; kq <= 12; ++kq) {
          const long long t0 = clock64();
          for (int i = 0; i < kq; ++i) tc_mma_tf32(tm, da0, db0, idesc, 1);
          const long long t1 = clock64();
          tc_commit(b0); mbar_wait(b0, ph); ph ^= 1;
          out[8 + kq] = t1 - t0;
        }
      }
      *stop = 1;
    }
  } else if (warp >= 2 && warp < 6 && (v.bg == BG_STS || v.bg == BG_BOTH)) {
    // loader-like traffic: every lane stores 16 B, conflict-free, as fast as the LSU takes them
    float4* dst = reinterpret_cast<float4*>(Bg) + (warp - 2) * 256 + lane;
    const float4 val = make_float4(1.f, 2.f, 3.f, 4.f);
    while (!*stop) {
#pragma unroll
      for (int i = 0; i < 8; ++i) dst[i * 32] = val;
      bg_count += 8 * 512;
    }
  } else if (warp >= 2 && warp < 6 && v.bg == BG_LDS) {
    const float4* src = reinterpret_cast<const float4*>(Bg) + (warp - 2) * 256 + lane;
    float acc = 0.f;
    while (!*stop) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { float4 q = src[i * 32]; acc += q.x + q.w; }
      bg_count += 8 * 512;
    }
    if (acc == 123.f) out[7] = 1;
  } else if (warp == 6 && lane == 0 && (v.bg == BG_BULK || v.bg == BG_BOTH)) {
    // weight-ring-like traffic: 16 KB bulk copies from L2 into two alternating slots
    int ph[2] = {0, 0};
    int s = 0;
    unsigned char* base = Bg + 16 * 1024;
    mbar_expect_tx(b0 + 8 * 1, 16384); bulk_g2s(smem_u32(base), gsrc, 16384, b0 + 8 * 1);
    mbar_expect_tx(b0 + 8 * 2, 16384); bulk_g2s(smem_u32(base + 16384), gsrc + 4096, 16384, b0 + 8 * 2);
    while (!*stop) {
      mbar_wait(b0 + 8 * (1 + s), ph[s]); ph[s] ^= 1;
      bg_count += 16384;
      mbar_expect_tx(b0 + 8 * (1 + s), 16384);
      bulk_g2s(smem_u32(base + s * 16384), gsrc + s * 4096, 16384, b0 + 8 * (1 + s));
      s ^= 1;
    }
    mbar_wait(b0 + 8 * 1, ph[0]); mbar_wait(b0 + 8 * 2, ph[1]);
  }
  if (v.busy && (warp == 4 || warp == 8 || warp == 12)) {
    float x = (float)lane;
    while (!*stop) {
#pragma unroll
      for (int i = 0; i < 64; ++i) x = fmaf(x, 1.0001f, 0.5f);
    }
    if (x == 123.f) out[6] = 1;
  }
  if (v.ld && warp >= 8 && warp < 12 && !(v.busy && warp == 8)) {
    uint32_t r[16];
    float acc = 0.f;
    const uint32_t t0 = tm + ((uint32_t)((warp & 3) * 32) << 16) + 256u;
    while (!*stop) {
#pragma unroll 1
      for (int c = 0; c < 256; c += 16) { tmem_ld16(t0 + c, r); tmem_ld_wait(); acc += __uint_as_float(r[0]) + __uint_as_float(r[15]); }
      bg_count += 32 * 256 * 4;
    }
    if (acc == 123.f) out[6] = 1;
  }
  if (lane == 0 && bg_count) atomicAdd((unsigned long long*)&out[1 + (warp == 6 ? 1 : 0)], (unsigned long long)bg_count);
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
}

int main(int argc, char** argv) {
  const int grid = argc > 1 ? atoi(argv[1]) : 1;
  long long* out; long long h[24];
  float* gsrc;
  cudaMalloc(&out, 192); cudaMalloc(&gsrc, 1 << 20); cudaMemset(gsrc, 0, 1 << 20);
  const int smem = 2 * kOpBytes + kBgBytes + 256;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const Variant vs[] = {
      {KIND_TF32, 128, LAY_NOSW, 0, BG_NONE, 0, "tf32 N128 nosw same-addr"},
      {KIND_TF32, 128, LAY_NOSW, 0, BG_NONE, 1, "tf32 N128 nosw rot"},
      {KIND_TF32, 256, LAY_NOSW, 0, BG_NONE, 1, "tf32 N256 nosw rot"},
      {KIND_TF32, 64, LAY_NOSW, 0, BG_NONE, 1, "tf32 N64  nosw rot"},
      {KIND_TF32, 80, LAY_NOSW, 0, BG_NONE, 1, "tf32 N80  nosw rot"},
      {KIND_BF16, 128, LAY_NOSW, 0, BG_NONE, 1, "bf16 N128 nosw rot"},
      {KIND_BF16, 256, LAY_NOSW, 0, BG_NONE, 1, "bf16 N256 nosw rot"},
      {KIND_TF32, 128, LAY_SW128, 0, BG_NONE, 1, "tf32 N128 sw128 rot"},
      {KIND_TF32, 256, LAY_SW128, 0, BG_NONE, 1, "tf32 N256 sw128 rot"},
      {KIND_TF32, 128, LAY_NOSW, 1, BG_NONE, 1, "tf32 N128 A-tmem"},
      {KIND_TF32, 256, LAY_NOSW, 1, BG_NONE, 1, "tf32 N256 A-tmem"},
      {KIND_TF32, 128, LAY_NOSW, 0, BG_STS, 1, "tf32 N128 nosw + STS"},
      {KIND_TF32, 256, LAY_NOSW, 0, BG_STS, 1, "tf32 N256 nosw + STS"},
      {KIND_TF32, 128, LAY_NOSW, 0, BG_LDS, 1, "tf32 N128 nosw + LDS"},
      {KIND_TF32, 256, LAY_NOSW, 0, BG_LDS, 1, "tf32 N256 nosw + LDS"},
      {KIND_TF32, 128, LAY_NOSW, 0, BG_BULK, 1, "tf32 N128 nosw + bulk"},
      {KIND_TF32, 256, LAY_NOSW, 0, BG_BULK, 1, "tf32 N256 nosw + bulk"},
      {KIND_TF32, 128, LAY_NOSW, 0, BG_BOTH, 1, "tf32 N128 nosw + STS + bulk"},
      {KIND_TF32, 256, LAY_NOSW, 0, BG_BOTH, 1, "tf32 N256 nosw + STS + bulk"},
      {KIND_TF32, 256, LAY_SW128, 0, BG_BOTH, 1, "tf32 N256 sw128 + STS + bulk"},
      {KIND_TF32, 256, LAY_NOSW, 1, BG_BOTH, 1, "tf32 N256 A-tmem + STS + bulk"},
      {KIND_TF32, 128, LAY_NOSW, 0, BG_NONE, 1, "N128 stage 4tf32+4bf16", 1},
      {KIND_TF32, 128, LAY_NOSW, 0, BG_NONE, 1, "N128 stage + commit", 2},
      {KIND_TF32, 128, LAY_NOSW, 0, BG_NONE, 1, "N128 stage + 1.33 commits", 3},
      {KIND_TF32, 128, LAY_NOSW, 0, BG_NONE, 1, "N128 stage + commit, data", 2, 0, 0, 1},
      {KIND_TF32, 128, LAY_NOSW, 0, BG_NONE, 1, "N128 stage + commit, busy", 2, 1, 0, 0},
      {KIND_TF32, 128, LAY_NOSW, 0, BG_NONE, 1, "N128 stage + commit, tmem ld", 2, 0, 1, 0},
      {KIND_TF32, 128, LAY_NOSW, 0, BG_BOTH, 1, "N128 stage+commit,all bg,data", 2, 1, 1, 1},
      {KIND_TF32, 256, LAY_NOSW, 0, BG_NONE, 1, "N256 stage + commit", 2},
      {KIND_TF32, 256, LAY_NOSW, 0, BG_BOTH, 1, "N256 stage+commit,all bg,data", 2, 1, 1, 1},
      {KIND_TF32, 128, LAY_NOSW, 0, BG_NONE, 1, "N128 stage+commit+1 wait", 4},
      {KIND_TF32, 128, LAY_NOSW, 0, BG_NONE, 1, "N128 stage+commit+2 waits", 5},
      {KIND_TF32, 128, LAY_NOSW, 0, BG_BOTH, 1, "N128 stage+commit+2 waits, all bg", 5, 1, 1, 1},
      {KIND_TF32, 256, LAY_NOSW, 0, BG_NONE, 1, "N256 stage+commit+2 waits", 5},
      {KIND_TF32, 128, LAY_NOSW, 0, BG_NONE, 1, "tf32 N128 pure, data", 0, 0, 0, 1},
      {KIND_BF16, 128, LAY_NOSW, 0, BG_NONE, 1, "bf16 N128 pure, data", 0, 0, 0, 1},
  };
  printf("grid %d, %d MMAs per measurement (best of 4)\n", grid, kNMma);
  for (const Variant& v : vs) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaMemset(out, 0, 192);
      k<<<grid, 512, smem>>>(v, gsrc, out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%-32s error %s\n", v.name, cudaGetErrorString(e)); return 1; }
    }
    cudaMemcpy(h, out, 192, cudaMemcpyDeviceToHost);
    const double clk = (double)h[0] / kNMma;
    // background bytes are counted over all 4 repetitions of all CTAs; normalise per CTA and per measured clock
    printf("%-32s %7.1f clk/MMA  (%5.1f clk per 128x128x8-equivalent)  bg: STS/LDS %.1f B/clk, bulk %.1f B/clk\n", v.name, clk,
           clk * 128.0 / v.N, (double)h[1] / grid / (4.0 * h[0]), (double)h[2] / grid / (4.0 * h[0]));
    if (h[9]) { printf("   issue time of k back-to-back MMAs:"); for (int kq = 1; kq <= 12; ++kq) printf(" %lld", h[8 + kq]); printf("\n"); }
  }
  return 0;
}
