// Host-side helpers of the tensor-core kernels: CUtensorMap construction with a small cache.
//
// A training step issues ~350 tensor-core launches with 2-5 tensor maps each.  Weights (and, with stable batch shapes, the
// arena activations) come back at the same (address, shape) step after step, so encoded maps are kept in a direct-mapped
// cache keyed by everything that enters the encoding; a hit is one 64-byte compare + one 128-byte copy.
#pragma once
#include "common.cuh"
#include <cuda.h>
#include <stdlib.h>

namespace tc {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)ptr;
  }
  return fn;
}

struct MapKey {
  const void* base; int64_t rows, cols, ld; int box_cols, box_rows, fp32, pad;
  bool operator==(const MapKey& o) const {
    return base == o.base && rows == o.rows && cols == o.cols && ld == o.ld && box_cols == o.box_cols && box_rows == o.box_rows &&
           fp32 == o.fp32;
  }
};
struct MapSlot { MapKey key; CUtensorMap map; bool valid; };
constexpr int MAP_CACHE_SLOTS = 4096;

// 2-D tensor [rows][cols] (cols contiguous, row stride ld elements); box = {box_cols, box_rows}; 128-byte swizzle
static int make_map(CUtensorMap* m, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_cols, int box_rows,
                    bool fp32 = false) {
  // one cache per translation unit and host thread (the C ABI is re-entrant per stream; SURVEY §8b: one thread calls the model)
  static thread_local MapSlot* cache = nullptr;
  if (!cache) cache = (MapSlot*)calloc(MAP_CACHE_SLOTS, sizeof(MapSlot));
  MapKey key{base, rows, cols, ld, box_cols, box_rows, fp32 ? 1 : 0, 0};
  uint64_t h = (uint64_t)(uintptr_t)base * 0x9E3779B97F4A7C15ull ^ (uint64_t)rows * 0xC2B2AE3D27D4EB4Full ^ (uint64_t)cols * 0x165667B19E3779F9ull ^
               (uint64_t)ld * 0x27D4EB2F165667C5ull ^ (uint64_t)(box_cols * 131 + box_rows * 7 + (fp32 ? 1 : 0));
  MapSlot* slot = cache ? &cache[(h >> 40) & (MAP_CACHE_SLOTS - 1)] : nullptr;
  if (slot && slot->valid && slot->key == key) { *m = slot->map; return 0; }
  EncodeTiledFn enc = get_encode();
  if (!enc) { lotus_set_error("cuTensorMapEncodeTiled unavailable"); return 1; }
  int es = fp32 ? 4 : 2;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * es};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, fp32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)base, dims, strides, box,
                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    lotus_set_error("cuTensorMapEncodeTiled failed (%d): base=%p rows=%lld cols=%lld ld=%lld box=%dx%d fp32=%d (base must be 16-byte "
                    "aligned, row stride a multiple of 16 bytes)", (int)r, base, (long long)rows, (long long)cols, (long long)ld,
                    box_cols, box_rows, (int)fp32);
    return 1;
  }
  if (slot) { slot->key = key; slot->map = *m; slot->valid = true; }
  return 0;
}

}  // namespace tc
