// Launch helper for the small kernels that sit between the GEMMs of a step (BatchNorm passes, ordered
// reductions, gradient reduce, input pack): with programmatic dependent launch the kernel is
// scheduled while its predecessor drains, and its first instruction -- griddepcontrol.wait, see
// pdl_entry() -- holds it until the predecessor's writes are visible.  The gap between two dependent
// kernels shrinks from a full launch latency to the hand-off (~1 us each, dozens per training step).
#pragma once
#include <cuda_runtime.h>
#include <string.h>

namespace vp3d {

bool conv_gemm_pdl_enabled();   // conv_gemm.cu: VP3D_PDL / vp3d_set_pdl

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                              cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  int n = 0;
  if (conv_gemm_pdl_enabled()) {
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    n = 1;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

#ifdef __CUDACC__
// First statement of every kernel launched through launch_pdl: nothing the predecessor wrote may
// be touched before it (no-op for a plain launch).
__device__ __forceinline__ void pdl_entry() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
#endif

}  // namespace vp3d
