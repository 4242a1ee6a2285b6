// ea_internal.h — declarations shared between the translation units of libea_b200.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <string.h>
#include "../../include/editanything_b200.h"

typedef CUresult (*ea_tmap_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                      const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                      const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
// Resolved through cudaGetDriverEntryPoint (no link-time libcuda dependency).
ea_tmap_encode_fn ea_tmap_encode();
void ea_count_launch();

// Programmatic dependent launch (PDL): every kernel of this library calls griddepcontrol.wait before
// it touches global memory, so consecutive launches on a stream may overlap the next kernel's
// prologue (barrier init, TMEM allocation, tensor-map prefetch, block scheduling) with the tail of
// the previous one.  ea_set_pdl(0) / EA_PDL=0 falls back to plain stream serialisation.
int ea_pdl_enabled();
#ifdef __CUDACC__
#include <utility>
template <typename... KArgs, typename... Args>
static inline cudaError_t ea_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                    cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = ea_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(std::forward<Args>(args))...);
}
#endif
