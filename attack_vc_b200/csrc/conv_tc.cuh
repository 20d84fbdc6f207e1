// conv_tc.cuh -- tcgen05 / TMEM implicit-GEMM Conv1d (3xTF32 split, fp32 accumulate).  STUB: the
// tensor-core path is not wired yet; every conv runs on the exact-fp32 SIMT kernel.
#pragma once
#include <vector>

#include "conv_simt.cuh"
#include "host_util.h"

namespace avc {

struct TcPack {
  bool ok = false;
  float* img = nullptr;
  int k = 0, kc = 0, n = 0;
};

inline void tc_pack_conv(Arena&, TcPack& p, const std::vector<float>&, int k, int kc, int n) {
  p.ok = false; p.k = k; p.kc = kc; p.n = n;
}
inline bool tc_eligible(const ConvArgs&, const TcPack&, bool) { return false; }
inline void tc_init_attributes() {}
inline void launch_conv_tc(const ConvArgs&, const TcPack&, int, cudaStream_t) {}

}  // namespace avc
