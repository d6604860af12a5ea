#include "tma_desc.h"

#include <unordered_map>

namespace {

typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_tmapEncodeTiled get_encode_fn() {
  static PFN_tmapEncodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<PFN_tmapEncodeTiled>(p);
  }
  return fn;
}

struct TmapKey {
  const void* ptr;
  int64_t ld1, ld2;
  int d0, d1, d2, box0, box1;
  bool operator==(const TmapKey& o) const {
    return ptr == o.ptr && ld1 == o.ld1 && ld2 == o.ld2 && d0 == o.d0 && d1 == o.d1 && d2 == o.d2 && box0 == o.box0 &&
           box1 == o.box1;
  }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr);
    auto mix = [&](size_t v) { h ^= v * 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2); };
    mix(static_cast<size_t>(k.ld1));
    mix(static_cast<size_t>(k.ld2));
    mix((static_cast<size_t>(k.d0) << 32) | static_cast<uint32_t>(k.d1));
    mix((static_cast<size_t>(k.d2) << 32) | (static_cast<uint32_t>(k.box0) << 16) | static_cast<uint32_t>(k.box1));
    return h;
  }
};

}  // namespace

int make_tmap_bf16(const bf16* ptr, int d0, int d1, int d2, int64_t ld1, int64_t ld2, int box0, int box1, CUtensorMap* out) {
  static thread_local std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
  TmapKey key{ptr, ld1, ld2, d0, d1, d2, box0, box1};
  auto it = cache.find(key);
  if (it != cache.end()) {
    *out = it->second;
    return 0;
  }
  PFN_tmapEncodeTiled enc = get_encode_fn();
  VLA_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled driver entry point unavailable");
  VLA_REQUIRE(box0 * 2 <= 128 && box1 <= 256, "tensor map box %d x %d too large for the 128B swizzle", box0, box1);
  const cuuint32_t rank = d2 > 1 ? 3 : 2;
  cuuint64_t gdim[3] = {static_cast<cuuint64_t>(d0), static_cast<cuuint64_t>(d1), static_cast<cuuint64_t>(d2)};
  cuuint64_t gstride[2] = {static_cast<cuuint64_t>(ld1) * 2, static_cast<cuuint64_t>(ld2) * 2};
  cuuint32_t box[3] = {static_cast<cuuint32_t>(box0), static_cast<cuuint32_t>(box1), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<bf16*>(ptr), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  VLA_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed: CUresult %d (ptr %p dims %d x %d x %d ld %lld %lld)", (int)r,
              (const void*)ptr, d0, d1, d2, (long long)ld1, (long long)ld2);
  if (cache.size() > 16384) cache.clear();
  cache.emplace(key, *out);
  return 0;
}

// 4-D variant: bf16 tensor [d3][d2][d1][d0] (d0 contiguous; strides ld1, ld2, ld3 in ELEMENTS), box = box0 x 1 x box2 x 1.
// Used by the attention kernels with d0 = head dim, d1 = head index, d2 = sequence position, d3 = batch: columns beyond
// the head dim (72 -> 128) and rows beyond the sequence are zero filled by the TMA unit.
int make_tmap_bf16_4d(const bf16* ptr, int d0, int d1, int d2, int d3, int64_t ld1, int64_t ld2, int64_t ld3, int box0, int box2,
                      CUtensorMap* out) {
  struct Key4 {
    const void* ptr;
    int64_t ld1, ld2, ld3;
    int d0, d1, d2, d3, box0, box2;
    bool operator==(const Key4& o) const {
      return ptr == o.ptr && ld1 == o.ld1 && ld2 == o.ld2 && ld3 == o.ld3 && d0 == o.d0 && d1 == o.d1 && d2 == o.d2 && d3 == o.d3 &&
             box0 == o.box0 && box2 == o.box2;
    }
  };
  struct Key4Hash {
    size_t operator()(const Key4& k) const {
      size_t h = reinterpret_cast<size_t>(k.ptr);
      auto mix = [&](size_t v) { h ^= v * 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2); };
      mix(static_cast<size_t>(k.ld1));
      mix(static_cast<size_t>(k.ld2));
      mix(static_cast<size_t>(k.ld3));
      mix((static_cast<size_t>(k.d0) << 32) | static_cast<uint32_t>(k.d1));
      mix((static_cast<size_t>(k.d2) << 32) | static_cast<uint32_t>(k.d3));
      mix((static_cast<size_t>(k.box0) << 32) | static_cast<uint32_t>(k.box2));
      return h;
    }
  };
  static thread_local std::unordered_map<Key4, CUtensorMap, Key4Hash> cache;
  Key4 key{ptr, ld1, ld2, ld3, d0, d1, d2, d3, box0, box2};
  auto it = cache.find(key);
  if (it != cache.end()) {
    *out = it->second;
    return 0;
  }
  PFN_tmapEncodeTiled enc = get_encode_fn();
  VLA_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled driver entry point unavailable");
  VLA_REQUIRE(box0 * 2 <= 128 && box2 <= 256, "tensor map box %d x %d too large for the 128B swizzle", box0, box2);
  VLA_REQUIRE((ld1 * 2) % 16 == 0 && (ld2 * 2) % 16 == 0 && (ld3 * 2) % 16 == 0, "tensor map strides must be multiples of 16 bytes");
  cuuint64_t gdim[4] = {static_cast<cuuint64_t>(d0), static_cast<cuuint64_t>(d1), static_cast<cuuint64_t>(d2),
                        static_cast<cuuint64_t>(d3)};
  cuuint64_t gstride[3] = {static_cast<cuuint64_t>(ld1) * 2, static_cast<cuuint64_t>(ld2) * 2, static_cast<cuuint64_t>(ld3) * 2};
  cuuint32_t box[4] = {static_cast<cuuint32_t>(box0), 1, static_cast<cuuint32_t>(box2), 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<bf16*>(ptr), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  VLA_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (4-D) failed: CUresult %d (ptr %p dims %d x %d x %d x %d)", (int)r,
              (const void*)ptr, d0, d1, d2, d3);
  if (cache.size() > 16384) cache.clear();
  cache.emplace(key, *out);
  return 0;
}
