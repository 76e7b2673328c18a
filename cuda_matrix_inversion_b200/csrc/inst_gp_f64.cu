#define INVGPU_TILE_DEFINE
#include "tile_launch.cuh"
#include "tile_configs.h"
INVGPU_TILE_GP_F64(INVGPU_TILE_INSTANTIATE_GP)
