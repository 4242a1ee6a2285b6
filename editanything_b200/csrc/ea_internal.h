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
