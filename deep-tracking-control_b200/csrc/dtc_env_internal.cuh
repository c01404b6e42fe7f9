#pragma once
#include <cuda.h>
#include "dtc_common.cuh"

struct dtc_env {
  dtc_env_config cfg;
  dtc_env_buffers buf;
  bool bound;
  dtc_env_config* d_cfg;  // device copy (tables are too large for kernel parameters)
  int16_t* min3;          // variant 5: min(H[x][y], H[x+1][y], H[x][y+1]) of the bound heightmap (library-owned)
  size_t min3_bytes;
};
int dtc_env_build_min3(dtc_env* e);  // dtc_foothold.cu
