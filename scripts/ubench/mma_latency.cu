// tcgen05.mma (kind::tf32, M=128, N=128, K=8) latency / throughput as seen by the issuing thread.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../attack_vc_b200/csrc/conv_tc.cuh"
using namespace avc;

__global__ void k(long long* out) {
  extern __shared__ __align__(128) unsigned char sm[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + 140000);
  const uint32_t b0 = smem_u32(bars);
  const int warp = threadIdx.x >> 5;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 140000 / 4; i += blockDim.x) reinterpret_cast<float*>(sm)[i] = 0.f;
  if (threadIdx.x == 0) { for (int i = 1; i <= 20; ++i) mbar_init(b0 + 8 * i, 1); mbar_init(b0, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t a = smem_u32(sm), b = smem_u32(sm + 70000 / 128 * 128);
    const uint64_t da = tc_desc(a, 136 * 16, 128), db = tc_desc(b, 128 * 16, 128);
    int ph = 0, o = 0;
    const int ns[6] = {1, 3, 12, 36, 120, 480};
    for (int t = 0; t < 6; ++t) {
      const int n = ns[t];
      long long best = 1LL << 60, issue = 0;
      for (int rep = 0; rep < 5; ++rep) {
        long long t0 = clock64();
        for (int i = 0; i < n; ++i) tc_mma_tf32(tm, da, db, idesc, 1);
        long long t1 = clock64();
        tc_commit(b0); mbar_wait(b0, ph); ph ^= 1;
        long long t2 = clock64();
        if (t2 - t0 < best) { best = t2 - t0; issue = t1 - t0; }
      }
      out[o++] = best; out[o++] = issue;
    }
    // 20 groups of 12 MMAs, each followed by a commit (not waited), fence::after_thread_sync in between
    for (int variant = 0; variant < 2; ++variant) {
      long long t0 = clock64();
      for (int g = 0; g < 20; ++g) {
        if (variant) tc_fence_after();
        for (int i = 0; i < 12; ++i) tc_mma_tf32(tm, da, db, idesc, 1);
        if (g < 19) { /* commits to a barrier nobody waits on would break phases; skip */ }
      }
      tc_commit(b0); mbar_wait(b0, ph); ph ^= 1;
      out[o++] = clock64() - t0;
    }
    // A-operand alignment: start offset (tap shift) and LBO (bytes between the two K chunks)
    {
      const uint32_t offs[4] = {0, 16, 64, 0};
      const uint32_t lbos[4] = {136 * 16, 136 * 16, 136 * 16, 137 * 16};
      for (int v = 0; v < 4; ++v) {
        const uint64_t dav = tc_desc(a + offs[v], lbos[v], 128);
        long long t0 = clock64();
        for (int i = 0; i < 240; ++i) tc_mma_tf32(tm, dav, db, idesc, 1);
        tc_commit(b0); mbar_wait(b0, ph); ph ^= 1;
        out[16 + v] = clock64() - t0;
      }
    }
    // same with a tcgen05.commit after every group (to 20 distinct single-use barriers), as the conv kernel does
    {
      long long t0 = clock64();
      for (int g = 0; g < 20; ++g) {
        for (int i = 0; i < 12; ++i) tc_mma_tf32(tm, da, db, idesc, 1);
        tc_commit(b0 + 8 * (1 + g));
      }
      mbar_wait(b0 + 8 * 20, 0);
      out[o++] = clock64() - t0;
    }
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(256u) : "memory");
}

int main() {
  long long* out; long long h[24];
  cudaMalloc(&out, 256);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 141000);
  for (int rep = 0; rep < 2; ++rep) {
    k<<<1, 128, 141000>>>(out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h, out, 160, cudaMemcpyDeviceToHost);
    const int ns[6] = {1, 3, 12, 36, 120, 480};
    for (int t = 0; t < 6; ++t) printf("n=%3d MMAs: total %6lld clk (issue %5lld) -> %.1f clk/MMA\n", ns[t], h[2 * t], h[2 * t + 1], (double)h[2 * t] / ns[t]);
    printf("240 MMAs, A start +0/LBO 2176: %lld clk | +16 B: %lld | +64 B: %lld | +0/LBO 2192: %lld\n", h[16], h[17], h[18], h[19]);
    printf("240 MMAs in 20 groups: no fence %lld clk, fence::after_thread_sync per group %lld clk, commit per group %lld clk\n", h[12], h[13], h[14]);
  }
  return 0;
}
