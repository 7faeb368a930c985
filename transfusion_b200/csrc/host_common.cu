#include "host_common.cuh"
#include "ptx.cuh"

#include <mutex>
#include <string.h>
#include <unordered_map>

#include "../../include/xfusion.h"

namespace xf {

thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, []() {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeTiledFn>(p);
    }
  });
  return fn;
}

// ---- descriptor cache ("lazily created TMA descriptors keyed by (ptr, shape) and guarded by a mutex").
// A tensor map is a pure function of (address, extents, pitches, box, swizzle): a step re-encodes the same ~1300 maps, so
// they are looked up by a 64-bit hash of those fields and verified field by field; the table is dropped when it grows past
// 8192 entries (shapes changed) -- never stale: the same key always encodes the same descriptor.
namespace {
struct TmapKey {
  const void* ptr;
  uint32_t rank, swz;
  uint64_t gdim[4], gstr[3];
  uint32_t box[4];
  bool operator==(const TmapKey& o) const { return memcmp(this, &o, sizeof(TmapKey)) == 0; }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    const uint64_t* w = reinterpret_cast<const uint64_t*>(&k);
    uint64_t h = 0x9E3779B97F4A7C15ull;
    for (size_t i = 0; i < sizeof(TmapKey) / 8; ++i) { h ^= w[i]; h *= 0xff51afd7ed558ccdull; h ^= h >> 32; }
    return static_cast<size_t>(h);
  }
};
std::mutex g_tmap_mu;
std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> g_tmap_cache;
std::atomic<int64_t> g_tmap_hits{0}, g_tmap_misses{0};
}  // namespace

static CUresult encode_cached(EncodeTiledFn enc, CUtensorMap* out, cuuint32_t rank, const void* ptr, const cuuint64_t* gdim,
                              const cuuint64_t* gstr, const cuuint32_t* box, const cuuint32_t* estr, CUtensorMapSwizzle swz) {
  TmapKey key;
  memset(&key, 0, sizeof(key));
  key.ptr = ptr; key.rank = rank; key.swz = static_cast<uint32_t>(swz);
  for (cuuint32_t i = 0; i < rank; ++i) { key.gdim[i] = gdim[i]; key.box[i] = box[i]; }
  for (cuuint32_t i = 0; i + 1 < rank; ++i) key.gstr[i] = gstr[i];
  {
    std::lock_guard<std::mutex> lock(g_tmap_mu);
    auto it = g_tmap_cache.find(key);
    if (it != g_tmap_cache.end()) { *out = it->second; g_tmap_hits.fetch_add(1, std::memory_order_relaxed); return CUDA_SUCCESS; }
  }
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(ptr), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r == CUDA_SUCCESS) {
    g_tmap_misses.fetch_add(1, std::memory_order_relaxed);
    std::lock_guard<std::mutex> lock(g_tmap_mu);
    if (g_tmap_cache.size() >= 8192) g_tmap_cache.clear();
    g_tmap_cache.emplace(key, *out);
  }
  return r;
}

int make_tmap_2d_bf16(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld,
                      uint32_t box_cols, uint32_t box_rows, int swizzle_bytes) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return fail(-10, "cuTensorMapEncodeTiled not available (no CUDA driver?)");
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0) return fail(-11, "TMA base pointer not 16-byte aligned");
  if ((ld * 2) % 16 != 0) return fail(-12, "TMA row pitch %llu B not a multiple of 16", (unsigned long long)(ld * 2));
  if ((swizzle_bytes > 0 && static_cast<int>(box_cols * 2) > swizzle_bytes) || box_cols > 256 || box_rows > 256)
    return fail(-13, "bad TMA box %u x %u", box_cols, box_rows);
  const CUtensorMapSwizzle swz2 = swizzle_bytes == 0 ? CU_TENSOR_MAP_SWIZZLE_NONE
                                : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstr[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode_cached(enc, out, 2, ptr, gdim, gstr, box, estr, swz2);
  if (r != CUDA_SUCCESS) return fail(-14, "cuTensorMapEncodeTiled failed: %d (rows=%llu cols=%llu ld=%llu)", (int)r,
                                     (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld);
  return 0;
}

int make_tmap_3d_bf16(CUtensorMap* out, const void* ptr, uint64_t batch, uint64_t rows, uint64_t cols, uint64_t ld,
                      uint32_t box_cols, uint32_t box_rows, int swizzle_bytes) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return fail(-10, "cuTensorMapEncodeTiled not available (no CUDA driver?)");
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0) return fail(-11, "TMA base pointer not 16-byte aligned");
  if ((ld * 2) % 16 != 0) return fail(-12, "TMA row pitch %llu B not a multiple of 16", (unsigned long long)(ld * 2));
  if (static_cast<int>(box_cols * 2) > swizzle_bytes || box_rows > 256) return fail(-13, "bad TMA box %u x %u", box_cols, box_rows);
  const CUtensorMapSwizzle swz = swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
  cuuint64_t gdim[3] = {cols, rows, batch};
  cuuint64_t gstr[2] = {ld * 2, rows * ld * 2};
  cuuint32_t box[3] = {box_cols, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = encode_cached(enc, out, 3, ptr, gdim, gstr, box, estr, swz);
  if (r != CUDA_SUCCESS) return fail(-14, "cuTensorMapEncodeTiled(3d) failed: %d", (int)r);
  return 0;
}

int make_tmap_4d_bf16(CUtensorMap* out, const void* ptr, uint64_t nb1, uint64_t nb2, uint64_t rows, uint64_t cols, uint64_t ld,
                      uint64_t bs1, uint64_t bs2, uint32_t box_cols, uint32_t box_rows, int swizzle_bytes) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return fail(-10, "cuTensorMapEncodeTiled not available (no CUDA driver?)");
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0) return fail(-11, "TMA base pointer not 16-byte aligned");
  if (nb1 == 0) nb1 = 1;
  if (nb2 == 0) nb2 = 1;
  if (nb1 == 1 && bs1 == 0) bs1 = 8;   // a stride of an extent-1 dimension is never used but must be valid
  if (nb2 == 1 && bs2 == 0) bs2 = 8;
  if ((ld * 2) % 16 != 0 || (bs1 * 2) % 16 != 0 || (bs2 * 2) % 16 != 0)
    return fail(-12, "TMA pitches (row %llu, batch %llu / %llu elements) must be multiples of 8 elements", (unsigned long long)ld,
                (unsigned long long)bs1, (unsigned long long)bs2);
  if ((swizzle_bytes > 0 && static_cast<int>(box_cols * 2) > swizzle_bytes) || box_cols > 256 || box_rows > 256)
    return fail(-13, "bad TMA box %u x %u", box_cols, box_rows);
  const CUtensorMapSwizzle swz = swizzle_bytes == 0 ? CU_TENSOR_MAP_SWIZZLE_NONE
                               : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
  cuuint64_t gdim[4] = {cols, rows, nb2, nb1};
  cuuint64_t gstr[3] = {ld * 2, bs2 * 2, bs1 * 2};
  cuuint32_t box[4] = {box_cols, box_rows, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = encode_cached(enc, out, 4, ptr, gdim, gstr, box, estr, swz);
  if (r != CUDA_SUCCESS) return fail(-14, "cuTensorMapEncodeTiled(4d batched) failed: %d", (int)r);
  return 0;
}

int make_tmap_chunks_bf16(CUtensorMap* out, const void* ptr, uint64_t batch, uint64_t rows, uint64_t cols, uint64_t ld,
                          uint32_t box_rows, uint32_t box_chunks) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return fail(-10, "cuTensorMapEncodeTiled not available (no CUDA driver?)");
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0) return fail(-11, "TMA base pointer not 16-byte aligned");
  if ((ld * 2) % 16 != 0) return fail(-12, "TMA row pitch %llu B not a multiple of 16", (unsigned long long)(ld * 2));
  if (cols % 32 != 0 || box_rows > 256 || box_chunks > 256 || box_chunks == 0) return fail(-13, "bad chunked TMA box");
  cuuint64_t gdim[4] = {32, rows, cols / 32, batch};
  cuuint64_t gstr[3] = {ld * 2, 64, rows * ld * 2};
  cuuint32_t box[4] = {32, box_rows, box_chunks, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = encode_cached(enc, out, 4, ptr, gdim, gstr, box, estr, CU_TENSOR_MAP_SWIZZLE_64B);
  if (r != CUDA_SUCCESS) return fail(-14, "cuTensorMapEncodeTiled(4d chunks) failed: %d", (int)r);
  return 0;
}

int make_tmap_rowblock_bf16(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t bc,
                            uint32_t box_rows) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return fail(-10, "cuTensorMapEncodeTiled not available (no CUDA driver?)");
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0) return fail(-11, "TMA base pointer not 16-byte aligned");
  if ((ld * 2) % 16 != 0) return fail(-12, "TMA row pitch %llu B not a multiple of 16", (unsigned long long)(ld * 2));
  if (bc == 0 || bc > 256 || bc % 8 || cols % bc || cols / bc > 256 || box_rows > 256) return fail(-13, "bad row-block TMA box");
  cuuint64_t gdim[3] = {bc, rows, cols / bc};
  cuuint64_t gstr[2] = {ld * 2, static_cast<cuuint64_t>(bc) * 2};
  cuuint32_t box[3] = {bc, box_rows, static_cast<cuuint32_t>(cols / bc)};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = encode_cached(enc, out, 3, ptr, gdim, gstr, box, estr, CU_TENSOR_MAP_SWIZZLE_NONE);
  if (r != CUDA_SUCCESS) return fail(-14, "cuTensorMapEncodeTiled(row block) failed: %d", (int)r);
  return 0;
}

int sm_count() {
  static std::atomic<int> n[64];   // zero-initialised; one slot per device
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  int v = n[dev].load(std::memory_order_relaxed);
  if (v == 0) {
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) return 148;
    n[dev].store(v, std::memory_order_relaxed);
  }
  return v;
}

// Odd column hashes of the GEMM / LayerNorm dropout sites (ptx.cuh:drop_colodd), one table per device, filled
// from the host on first use (synchronous copy: the first call must not happen inside a stream capture).
const uint32_t* drop_col_table() {
  static const uint32_t* tabs[64] = {nullptr};
  static std::mutex mu;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  std::lock_guard<std::mutex> lock(mu);
  if (!tabs[dev]) {
    static uint32_t host_tab[XF_DROP_TABLE_COLS];
    for (uint32_t c = 0; c < XF_DROP_TABLE_COLS; ++c) host_tab[c] = drop_colodd(c);
    uint32_t* d = nullptr;
    if (cudaMalloc(&d, sizeof(host_tab)) != cudaSuccess) return nullptr;
    if (cudaMemcpy(d, host_tab, sizeof(host_tab), cudaMemcpyHostToDevice) != cudaSuccess) { cudaFree(d); return nullptr; }
    tabs[dev] = d;
  }
  return tabs[dev];
}

}  // namespace xf

extern "C" {
int xf_version(void) { return XF_ABI_VERSION; }
const char* xf_last_error(void) { return xf::g_err; }
int64_t xf_launch_count(void) { return xf::g_launches.load(); }
int64_t xf_tmap_cache_stats(int which) { return which == 0 ? xf::g_tmap_hits.load() : which == 1 ? xf::g_tmap_misses.load() : 0; }
}
