#define INVGPU_TILE_DEFINE
#include "tile_launch.cuh"
#include "tile_configs.h"
INVGPU_GJR_F64(INVGPU_GJR_INSTANTIATE)
