#define INVGPU_TILE_DEFINE
#include "tile_launch.cuh"
#include "tile_configs.h"
INVGPU_GJT_F64(INVGPU_GJT_INSTANTIATE)
