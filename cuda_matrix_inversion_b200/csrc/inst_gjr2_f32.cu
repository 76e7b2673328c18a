#define INVGPU_TILE_DEFINE
#include "tile_launch.cuh"
#include "tile_configs.h"
INVGPU_GJR2_F32(INVGPU_GJR2_INSTANTIATE)
