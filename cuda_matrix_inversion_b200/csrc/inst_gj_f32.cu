#define INVGPU_TILE_DEFINE
#include "tile_launch.cuh"
#include "tile_configs.h"
INVGPU_GJ_F32(INVGPU_GJ_INSTANTIATE)
