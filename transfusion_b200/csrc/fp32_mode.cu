// fp32-tolerance mode (north_star: "about 1e-5 relative in fp32"; the reference's Ego4Dv2 config trains with precision 32,
// runner/nao/configs/ego_nao_res50_ego4dv2.yml:124).  The tensor cores stay the engine: an fp32 operand is split into three
// bf16 terms a = a0 + a1 + a2 (24 mantissa bits) and the product a.b ~ sum_{i+j<=2} a_i b_j is obtained from the ORDINARY
// bf16 GEMM kernel by concatenating the terms along K:
//     A' = [a0 | a0 | a0 | a1 | a1 | a2]   (rows x 6K),     B' = [b0 | b1 | b2 | b0 | b1 | b0]   (rows x 6K)
// (products of bf16 numbers are exact in the fp32 accumulator; the three dropped cross terms are below 2^-24 relative).
// This file holds the two element-wise helpers the mode needs: the 3-way split and an fp32 row softmax with key padding.
#include <string.h>

#include "../../include/xfusion.h"
#include "host_common.cuh"
#include "ptx.cuh"

namespace xf {

__device__ __forceinline__ void split3(float a, __nv_bfloat16& t0, __nv_bfloat16& t1, __nv_bfloat16& t2) {
  t0 = __float2bfloat16(a);
  const float r1 = a - __bfloat162float(t0);      // exact in fp32
  t1 = __float2bfloat16(r1);
  t2 = __float2bfloat16(r1 - __bfloat162float(t1));
}

// `extra` = 8: the row gets 8 more columns that carry a bias through the GEMM itself (so the split-K reduction epilogue, which
// cannot add one, is enough): A side (1, 1, 1, 0 ...), B side (bias[r]_0, bias[r]_1, bias[r]_2, 0 ...).  `act` = 1 applies
// the exact GELU (erf) to the source first (the A operand of linear2, torch18_adapters.py:111).
__global__ void __launch_bounds__(256) split3_kernel(const float* __restrict__ src, long long lds, int rows, int cols,
                                                     __nv_bfloat16* __restrict__ dst, int pattern, int act, int extra,
                                                     const float* __restrict__ bias) {
  const long long total = static_cast<long long>(rows) * cols;
  const long long ldd = 6ll * cols + extra;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / cols;
    const int c = static_cast<int>(i - r * cols);
    float a = src[r * lds + c];
    if (act == 1) a = 0.5f * a * (1.f + erff(a * 0.70710678118654752f));
    __nv_bfloat16 t0, t1, t2;
    split3(a, t0, t1, t2);
    __nv_bfloat16* d = dst + r * ldd + c;
    if (pattern == 0) { d[0] = t0; d[cols] = t0; d[2 * cols] = t0; d[3 * cols] = t1; d[4 * cols] = t1; d[5 * cols] = t2; }
    else              { d[0] = t0; d[cols] = t1; d[2 * cols] = t2; d[3 * cols] = t0; d[4 * cols] = t1; d[5 * cols] = t0; }
    if (extra && c == 0) {
      __nv_bfloat16* e = dst + r * ldd + 6ll * cols;
      const __nv_bfloat16 zero = __float2bfloat16(0.f), one = __float2bfloat16(1.f);
      __nv_bfloat16 b0 = zero, b1 = zero, b2 = zero;
      if (pattern == 1 && bias) split3(bias[r], b0, b1, b2);
      e[0] = pattern == 0 ? one : b0; e[1] = pattern == 0 ? one : b1; e[2] = pattern == 0 ? one : b2;
      for (int k = 3; k < extra; ++k) e[k] = zero;
    }
  }
}

// s[bh, q, :] <- softmax over k < Sk of (scale * s[bh, q, k]) with keys masked by kpm[b, k] != 0 (-inf); columns k >= Sk of
// the padded row (pitch Sp) are written as 0.  One warp per row (torch18_adapters.py:578-597,789-798 in fp32).
__global__ void __launch_bounds__(256) softmax_rows_f32_kernel(float* __restrict__ s, long long rows, int Sq, int Sk, int Sp, int H,
                                                               const uint8_t* __restrict__ kpm, float scale) {
  const int lane = threadIdx.x & 31;
  const long long row = blockIdx.x * 8ll + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int b = static_cast<int>(row / (static_cast<long long>(Sq) * H));
  float* p = s + row * Sp;
  const uint8_t* m = kpm ? kpm + static_cast<long long>(b) * Sk : nullptr;
  float mx = -INFINITY;
  for (int k = lane; k < Sk; k += 32)
    if (!(m && m[k])) mx = fmaxf(mx, p[k] * scale);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
  for (int k = lane; k < Sk; k += 32) {
    const float e = (m && m[k]) ? 0.f : expf(p[k] * scale - mx);
    p[k] = e;
    sum += e;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float inv = 1.f / sum;
  for (int k = lane; k < Sp; k += 32) p[k] = k < Sk ? p[k] * inv : 0.f;
}

}  // namespace xf

using namespace xf;

extern "C" int xf_split3(const float* src, int64_t lds, int rows, int cols, void* dst_bf16, int pattern, int act, int bias_cols,
                         const float* bias, xf_stream_t s) {
  if (!src || !dst_bf16) return fail(-1, "xf_split3: null pointer");
  if (pattern != 0 && pattern != 1) return fail(-2, "xf_split3: pattern must be 0 (A side) or 1 (B side)");
  if (bias_cols != 0 && bias_cols != 8) return fail(-3, "xf_split3: bias_cols must be 0 or 8");
  if (act != 0 && act != 1) return fail(-4, "xf_split3: act must be 0 or 1 (GELU erf)");
  if (rows <= 0 || cols <= 0) return 0;
  const long long total = static_cast<long long>(rows) * cols;
  long long ctas = (total + 255) / 256;
  const long long cap = 16ll * sm_count();
  if (ctas > cap) ctas = cap;
  split3_kernel<<<static_cast<int>(ctas), 256, 0, reinterpret_cast<cudaStream_t>(s)>>>(
      src, lds, rows, cols, reinterpret_cast<__nv_bfloat16*>(dst_bf16), pattern, act, bias_cols, bias);
  g_launches.fetch_add(1);
  XF_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int xf_softmax_rows_f32(float* s, int B, int H, int Sq, int Sk, int Sp, const uint8_t* kpm, float scale, xf_stream_t st) {
  if (!s) return fail(-1, "xf_softmax_rows_f32: null pointer");
  if (Sp < Sk || B <= 0 || H <= 0 || Sq <= 0 || Sk <= 0) return fail(-2, "xf_softmax_rows_f32: bad shape");
  const long long rows = static_cast<long long>(B) * H * Sq;
  softmax_rows_f32_kernel<<<static_cast<int>((rows + 7) / 8), 256, 0, reinterpret_cast<cudaStream_t>(st)>>>(s, rows, Sq, Sk, Sp, H, kpm, scale);
  g_launches.fetch_add(1);
  XF_CUDA(cudaGetLastError());
  return 0;
}
