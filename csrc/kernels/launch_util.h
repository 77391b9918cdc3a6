// Host-side launch helper: every hot kernel is launched with the programmatic-dependent-launch
// attribute so that its grid launch, prologue (barrier init, TMEM allocation, tensormap prefetch)
// and CTA scheduling overlap the tail of the previous kernel in the stream (also inside captured
// CUDA graphs: CUDA >= 12.3 records programmatic edges).  The kernels call griddepcontrol.wait
// before they touch any memory produced by earlier kernels, so ordering semantics are unchanged.
// SKY_PDL=0 disables the attribute (A/B testing).
#pragma once
#include <cuda_runtime.h>

#include <cstdlib>
#include <utility>

namespace sky {

inline bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = std::getenv("SKY_PDL");
    v = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem,
                              cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(std::forward<Args>(args))...);
}

}  // namespace sky
