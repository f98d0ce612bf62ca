// stand-in for the driver API header: only what ffmpeg_b200/csrc/common.h names (the TMA kernels themselves are not emulated)
#pragma once
#include <stdint.h>
typedef struct CUtensorMap_st { alignas(64) uint64_t opaque[16]; } CUtensorMap;
enum { CU_TENSOR_MAP_SWIZZLE_NONE = 0, CU_TENSOR_MAP_SWIZZLE_32B = 1 };
