// Micro-benchmarks for the conv_tc pipeline: latency of (1) tcgen05.commit -> mbarrier -> try_wait by the
// same thread, (2) the two-thread ping-pong commit -> producer wait -> arrive -> consumer wait,
// (3) a 32 KB cp.async.bulk from L2 -> mbarrier complete.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../attack_vc_b200/csrc/conv_tc.cuh"
using namespace avc;

__global__ void k(const float* w, long long* out) {
  extern __shared__ __align__(128) unsigned char sm[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + 65536);
  const uint32_t b0 = smem_u32(bars), b1 = b0 + 8, b2 = b0 + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __shared__ uint32_t slot;
  if (threadIdx.x == 0) { mbar_init(b0, 1); mbar_init(b1, 1); mbar_init(b2, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(32u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const int R = 200;
  if (threadIdx.x == 0) {
    // (1) self round trip
    long long t0 = clock64();
    for (int i = 0; i < R; ++i) { tc_commit(b0); mbar_wait(b0, i & 1); }
    out[0] = (clock64() - t0) / R;
    // (3) 32 KB bulk copy, L2 resident after the first
    t0 = clock64();
    for (int i = 0; i < R; ++i) { mbar_expect_tx(b2, 32768); bulk_g2s(smem_u32(sm), w, 32768, b2); mbar_wait(b2, i & 1); }
    out[2] = (clock64() - t0) / R;
    // plain arrive + wait by the same thread
    t0 = clock64();
    for (int i = 0; i < R; ++i) { mbar_arrive(b2); mbar_wait(b2, (R + i) & 1); }
    out[3] = (clock64() - t0) / R;
  }
  __syncthreads();
  // (2) two-thread chain: thread 0 commits b0, thread 32 waits b0 then arrives b1, thread 0 waits b1
  if (threadIdx.x == 0) {
    long long t0 = clock64();
    for (int i = 0; i < R; ++i) { tc_commit(b0); mbar_wait(b1, i & 1); }
    out[1] = (clock64() - t0) / R;
  } else if (threadIdx.x == 32) {
    for (int i = 0; i < R; ++i) { mbar_wait(b0, (R + i) & 1); mbar_arrive(b1); }
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(32u) : "memory");
}

int main() {
  float* w; long long* out; long long h[4];
  cudaMalloc(&w, 1 << 20); cudaMemset(w, 0, 1 << 20); cudaMalloc(&out, 64);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000);
  for (int rep = 0; rep < 2; ++rep) {
    k<<<1, 64, 70000>>>(w, out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h, out, 32, cudaMemcpyDeviceToHost);
    printf("clk: commit->wait (same thread) %lld | commit->wait->arrive->wait (2 threads) %lld | 32KB bulk L2->smem %lld | arrive->wait %lld\n", h[0], h[1], h[2], h[3]);
  }
  return 0;
}
