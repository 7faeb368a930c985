// Host-side helpers shared by the C-ABI entry points: error string, launch counter,
// lazily resolved cuTensorMapEncodeTiled (no link-time dependency on libcuda, so the library
// loads on a CPU-only box), TMA descriptor construction.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <mutex>

namespace xf {

extern thread_local char g_err[512];
extern std::atomic<int64_t> g_launches;

inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
inline int cuda_fail(cudaError_t e, const char* what) {
  snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
  return static_cast<int>(e);
}

#define XF_CUDA(call)                                   \
  do {                                                  \
    cudaError_t _e = (call);                            \
    if (_e != cudaSuccess) return xf::cuda_fail(_e, #call); \
  } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_tiled();

// 2-D bf16 row-major tensor [rows, cols] with leading dimension ld (elements); box = box_cols x box_rows,
// SWIZZLE_128B (box_cols * 2 bytes must be <= 128), out-of-bounds elements read as zero.
int make_tmap_2d_bf16(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld,
                      uint32_t box_cols, uint32_t box_rows, int swizzle_bytes = 128);

// 3-D bf16 tensor [batch, rows, cols] (cols contiguous, row pitch ld, batch pitch rows*ld);
// box = box_cols x box_rows x 1, SWIZZLE_128B, out-of-bounds reads as zero (per batch).
int make_tmap_3d_bf16(CUtensorMap* out, const void* ptr, uint64_t batch, uint64_t rows, uint64_t cols, uint64_t ld,
                      uint32_t box_cols, uint32_t box_rows, int swizzle_bytes = 128);

// 3-D view of a [rows, cols] bf16 matrix as (bc-column chunk, row, chunk index), no swizzle: one TMA instruction moves
// a [box_rows x cols] row block into shared memory laid out [chunk][box_rows][bc].  cols % bc == 0, bc <= 256.
int make_tmap_rowblock_bf16(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t bc,
                            uint32_t box_rows);

// 4-D view of the same [batch, rows, cols] bf16 tensor as (32-column chunk, row, chunk index, batch): one TMA
// instruction moves `box_chunks` SWIZZLE_64B chunks of [box_rows x 64 B], laid out chunk-major in shared memory
// (the layout the attention kernels' 32-column-chunk descriptors expect).  cols % 32 == 0.
int make_tmap_chunks_bf16(CUtensorMap* out, const void* ptr, uint64_t batch, uint64_t rows, uint64_t cols, uint64_t ld,
                          uint32_t box_rows, uint32_t box_chunks);

// Batched 4-D view (inner columns, rows, batch2, batch1) of bf16 matrices [rows, cols] that sit at element offsets
// b1 * bs1 + b2 * bs2 from ptr (row pitch ld); box = box_cols x box_rows x 1 x 1.  Out-of-bounds rows / columns are
// zero-filled on loads and clipped on stores PER BATCH ENTRY.  bs1, bs2, ld must be multiples of 8 elements.
int make_tmap_4d_bf16(CUtensorMap* out, const void* ptr, uint64_t nb1, uint64_t nb2, uint64_t rows, uint64_t cols, uint64_t ld,
                      uint64_t bs1, uint64_t bs2, uint32_t box_cols, uint32_t box_rows, int swizzle_bytes = 128);

int sm_count();   // of the CURRENT device (cached per device)

// Per-device one-time setup (cudaFuncSetAttribute is per device / context): runs f() the first time it is reached
// on each device, under a mutex, and only marks the device done when f() returned 0.  Usage:
//   static DeviceOnce once;  if (int rc = once.run([&] { XF_CUDA(cudaFuncSetAttribute(...)); return 0; })) return rc;
struct DeviceOnce {
  std::atomic<uint64_t> done{0};
  template <typename F>
  int run(F&& f) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
    const uint64_t bit = 1ull << (dev & 63);
    if (done.load(std::memory_order_acquire) & bit) return 0;
    std::lock_guard<std::mutex> lock(mu);
    if (done.load(std::memory_order_relaxed) & bit) return 0;
    const int rc = f();
    if (rc == 0) done.fetch_or(bit, std::memory_order_release);
    return rc;
  }
  std::mutex mu;
};

// device table of drop_colodd(0 .. XF_DROP_TABLE_COLS-1) for the current device (nullptr on failure)
constexpr uint32_t XF_DROP_TABLE_COLS = 16384;
const uint32_t* drop_col_table();

}  // namespace xf
