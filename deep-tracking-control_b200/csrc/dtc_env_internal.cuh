#pragma once
#include <cuda.h>
#include "dtc_common.cuh"

struct dtc_env {
  dtc_env_config cfg;
  dtc_env_buffers buf;
  bool bound;
  dtc_env_config* d_cfg;  // device copy (tables are too large for kernel parameters)
  int16_t* min3;          // variants 5 / 6: min(H[x][y], H[x+1][y], H[x][y+1]) of the bound heightmap (library-owned)
  size_t min3_bytes;
  CUtensorMap min3_map;   // variant 6: 2-D tensor map of min3 (box = one 42 x 48-cell patch) for the TMA patch loads
  bool min3_map_ok;
  const int64_t* step_base;  // optional device-side counter added to every launch's step argument (dtc_env_set_step_base)
  float4* gtab;              // variant 6 with 7 CTAs / SM: grid-point table in global memory (dtc_foothold.cu: k_v6_gtab)
};
int dtc_env_build_min3(dtc_env* e);  // dtc_foothold.cu
