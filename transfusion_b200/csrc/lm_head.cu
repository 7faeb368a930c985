// LM head: PoolPredictor of modeling/cross_fusion/ego_fusion/lm_layers.py:30-81 --
// mask-multiply + mean / max pooling over the L language tokens, LayerNorm, optional GELU + Linear, noun / verb
// Linears -- forward and backward, in fp32 (the head is [B, D] -> a few hundred logits: launch-bound, no tensor
// cores; what matters is that it runs on the same stream without leaving the library).
#include <string.h>

#include "../../include/xfusion.h"
#include "host_common.cuh"

namespace xf {

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float gelu_exact(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_exact_grad(float x) {
  return 0.5f * (1.0f + erff(x * 0.70710678118654752f)) + x * 0.39894228040143268f * expf(-0.5f * x * x);
}

// ---- pooling over L: pooled[b, d] = mean_l / max_l (tok[b, l, d] * mask[b, l])   (lm_layers.py:60-66; the mean
// divides by the padded length L, masked rows enter the max as zeros -- both as in the reference)
__global__ void lm_pool_fwd_kernel(const float* __restrict__ tok, const uint8_t* __restrict__ mask, int B, int L, int D, int type,
                                   float* __restrict__ pooled, int* __restrict__ argmax) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * D) return;
  const int b = i / D, d = i - b * D;
  float acc = type == 0 ? 0.f : -INFINITY;
  int am = 0;
  for (int l = 0; l < L; ++l) {
    const float m = mask ? static_cast<float>(mask[b * L + l]) : 1.f;
    const float v = tok[(static_cast<long long>(b) * L + l) * D + d] * m;
    if (type == 0) acc += v;
    else if (v > acc) { acc = v; am = l; }
  }
  pooled[i] = type == 0 ? acc / L : acc;
  if (argmax) argmax[i] = am;
}
__global__ void lm_pool_bwd_kernel(const float* __restrict__ dpooled, const uint8_t* __restrict__ mask, const int* __restrict__ argmax,
                                   int B, int L, int D, int type, float* __restrict__ dtok) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= static_cast<long long>(B) * L * D) return;
  const int d = i % D;
  const long long bl = i / D;
  const int l = bl % L, b = bl / L;
  const float m = mask ? static_cast<float>(mask[b * L + l]) : 1.f;
  const float g = dpooled[b * D + d];
  dtok[i] = type == 0 ? g * m / L : (argmax[b * D + d] == l ? g * m : 0.f);
}

// ---- fp32 LayerNorm over the rows of a small [R, D] matrix: one CTA per row (nn.LayerNorm, eps 1e-5; lm_layers.py:68-69)
__global__ void __launch_bounds__(256) rowln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, int D, float eps, float* __restrict__ y,
                                                        float* __restrict__ mean_out, float* __restrict__ rstd_out) {
  __shared__ float red[2][8];
  const int r = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* xr = x + static_cast<long long>(r) * D;
  float s = 0.f;
  for (int d = threadIdx.x; d < D; d += 256) s += xr[d];
  s = warp_sum_f(s);
  if (lane == 0) red[0][warp] = s;
  __syncthreads();
  float tot = 0.f;
  for (int w = 0; w < 8; ++w) tot += red[0][w];
  const float mean = tot / D;
  float v = 0.f;
  for (int d = threadIdx.x; d < D; d += 256) { const float c = xr[d] - mean; v += c * c; }
  v = warp_sum_f(v);
  if (lane == 0) red[1][warp] = v;
  __syncthreads();
  float var = 0.f;
  for (int w = 0; w < 8; ++w) var += red[1][w];
  const float rstd = rsqrtf(var / D + eps);
  for (int d = threadIdx.x; d < D; d += 256) y[static_cast<long long>(r) * D + d] = (xr[d] - mean) * rstd * gamma[d] + beta[d];
  if (threadIdx.x == 0) { mean_out[r] = mean; rstd_out[r] = rstd; }
}
// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)), g = dy * gamma; dgamma += dy * xhat, dbeta += dy (atomics, R small)
__global__ void __launch_bounds__(256) rowln_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                        const float* __restrict__ gamma, const float* __restrict__ mean_in,
                                                        const float* __restrict__ rstd_in, int D, float* __restrict__ dx,
                                                        float* __restrict__ dgamma, float* __restrict__ dbeta) {
  __shared__ float red[2][8];
  const int r = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float mean = mean_in[r], rstd = rstd_in[r];
  const float* xr = x + static_cast<long long>(r) * D;
  const float* dyr = dy + static_cast<long long>(r) * D;
  float s1 = 0.f, s2 = 0.f;
  for (int d = threadIdx.x; d < D; d += 256) {
    const float g = dyr[d] * gamma[d], xh = (xr[d] - mean) * rstd;
    s1 += g; s2 += g * xh;
  }
  s1 = warp_sum_f(s1); s2 = warp_sum_f(s2);
  if (lane == 0) { red[0][warp] = s1; red[1][warp] = s2; }
  __syncthreads();
  float t1 = 0.f, t2 = 0.f;
  for (int w = 0; w < 8; ++w) { t1 += red[0][w]; t2 += red[1][w]; }
  t1 /= D; t2 /= D;
  for (int d = threadIdx.x; d < D; d += 256) {
    const float xh = (xr[d] - mean) * rstd;
    dx[static_cast<long long>(r) * D + d] = rstd * (dyr[d] * gamma[d] - t1 - xh * t2);
    atomicAdd(dgamma + d, dyr[d] * xh);
    atomicAdd(dbeta + d, dyr[d]);
  }
}

// ---- small fp32 linear: y[r, c] = sum_d act(x[r, d]) * W[c, d] + bias[c], act = identity | GELU(erf)
// (nn.Linear of lm_layers.py:47-55; the optional nn.Sequential(GELU, Linear) of :43-45).  One warp per output.
__global__ void __launch_bounds__(256) small_linear_fwd_kernel(const float* __restrict__ x, const float* __restrict__ W,
                                                               const float* __restrict__ bias, int R, int C, int D, int act,
                                                               float* __restrict__ y) {
  const int o = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (o >= R * C) return;
  const int r = o / C, c = o - r * C;
  const float* xr = x + static_cast<long long>(r) * D;
  const float* wc = W + static_cast<long long>(c) * D;
  float s = 0.f;
  for (int d = lane; d < D; d += 32) s += (act ? gelu_exact(xr[d]) : xr[d]) * wc[d];
  s = warp_sum_f(s);
  if (lane == 0) y[o] = s + (bias ? bias[c] : 0.f);
}
// dx[r, d] = act'(x[r, d]) * sum_c dy[r, c] W[c, d]
__global__ void small_linear_dx_kernel(const float* __restrict__ dy, const float* __restrict__ W, const float* __restrict__ x, int R,
                                       int C, int D, int act, float* __restrict__ dx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * D) return;
  const int r = i / D, d = i - r * D;
  float s = 0.f;
  for (int c = 0; c < C; ++c) s += dy[r * C + c] * W[static_cast<long long>(c) * D + d];
  dx[i] = act ? s * gelu_exact_grad(x[i]) : s;
}
// dW[c, d] += sum_r dy[r, c] act(x[r, d]);  dbias[c] += sum_r dy[r, c]
__global__ void small_linear_dw_kernel(const float* __restrict__ dy, const float* __restrict__ x, int R, int C, int D, int act,
                                       float* __restrict__ dW, float* __restrict__ dbias) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= static_cast<long long>(C) * D) return;
  const int c = i / D, d = i - static_cast<long long>(c) * D;
  float s = 0.f, sb = 0.f;
  for (int r = 0; r < R; ++r) {
    const float g = dy[r * C + c];
    const float xv = x[static_cast<long long>(r) * D + d];
    s += g * (act ? gelu_exact(xv) : xv);
    sb += g;
  }
  dW[i] += s;
  if (d == 0 && dbias) dbias[c] += sb;
}

}  // namespace xf

using namespace xf;

extern "C" int xf_lm_pool_fwd(const float* tok, const uint8_t* mask, int B, int L, int D, int type, float* pooled, int32_t* argmax,
                              xf_stream_t s) {
  if (!tok || !pooled) return fail(-1, "xf_lm_pool_fwd: null pointer");
  if (B <= 0 || L <= 0 || D <= 0 || (type != 0 && type != 1)) return fail(-2, "xf_lm_pool_fwd: bad shape / pooling type");
  if (type == 1 && !argmax) return fail(-3, "xf_lm_pool_fwd: max pooling needs the argmax buffer");
  lm_pool_fwd_kernel<<<(B * D + 255) / 256, 256, 0, reinterpret_cast<cudaStream_t>(s)>>>(tok, mask, B, L, D, type, pooled, argmax);
  g_launches.fetch_add(1);
  XF_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int xf_lm_pool_bwd(const float* dpooled, const uint8_t* mask, const int32_t* argmax, int B, int L, int D, int type,
                              float* dtok, xf_stream_t s) {
  if (!dpooled || !dtok) return fail(-1, "xf_lm_pool_bwd: null pointer");
  if (B <= 0 || L <= 0 || D <= 0 || (type != 0 && type != 1)) return fail(-2, "xf_lm_pool_bwd: bad shape / pooling type");
  if (type == 1 && !argmax) return fail(-3, "xf_lm_pool_bwd: max pooling needs the argmax buffer");
  const long long total = static_cast<long long>(B) * L * D;
  lm_pool_bwd_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(s)>>>(dpooled, mask, argmax, B, L,
                                                                                                               D, type, dtok);
  g_launches.fetch_add(1);
  XF_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int xf_rowln_fwd(const float* x, const float* gamma, const float* beta, int rows, int D, float eps, float* y, float* mean,
                            float* rstd, xf_stream_t s) {
  if (!x || !gamma || !beta || !y || !mean || !rstd) return fail(-1, "xf_rowln_fwd: null pointer");
  if (rows <= 0 || D <= 0) return fail(-2, "xf_rowln_fwd: bad shape");
  rowln_fwd_kernel<<<rows, 256, 0, reinterpret_cast<cudaStream_t>(s)>>>(x, gamma, beta, D, eps, y, mean, rstd);
  g_launches.fetch_add(1);
  XF_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int xf_rowln_bwd(const float* dy, const float* x, const float* gamma, const float* mean, const float* rstd, int rows, int D,
                            float* dx, float* dgamma, float* dbeta, xf_stream_t s) {
  if (!dy || !x || !gamma || !mean || !rstd || !dx || !dgamma || !dbeta) return fail(-1, "xf_rowln_bwd: null pointer");
  if (rows <= 0 || D <= 0) return fail(-2, "xf_rowln_bwd: bad shape");
  rowln_bwd_kernel<<<rows, 256, 0, reinterpret_cast<cudaStream_t>(s)>>>(dy, x, gamma, mean, rstd, D, dx, dgamma, dbeta);
  g_launches.fetch_add(1);
  XF_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int xf_small_linear_fwd(const float* x, const float* W, const float* bias, int rows, int C, int D, int act, float* y,
                                   xf_stream_t s) {
  if (!x || !W || !y) return fail(-1, "xf_small_linear_fwd: null pointer");
  if (rows <= 0 || C <= 0 || D <= 0) return fail(-2, "xf_small_linear_fwd: bad shape");
  small_linear_fwd_kernel<<<(rows * C + 7) / 8, 256, 0, reinterpret_cast<cudaStream_t>(s)>>>(x, W, bias, rows, C, D, act, y);
  g_launches.fetch_add(1);
  XF_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int xf_small_linear_bwd(const float* dy, const float* x, const float* W, int rows, int C, int D, int act, float* dx,
                                   float* dW, float* dbias, xf_stream_t s) {
  if (!dy || !x || !W) return fail(-1, "xf_small_linear_bwd: null pointer");
  if (rows <= 0 || C <= 0 || D <= 0) return fail(-2, "xf_small_linear_bwd: bad shape");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(s);
  if (dx) {
    small_linear_dx_kernel<<<(rows * D + 255) / 256, 256, 0, st>>>(dy, W, x, rows, C, D, act, dx);
    g_launches.fetch_add(1);
  }
  if (dW) {
    const long long total = static_cast<long long>(C) * D;
    small_linear_dw_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(dy, x, rows, C, D, act, dW, dbias);
    g_launches.fetch_add(1);
  }
  XF_CUDA(cudaGetLastError());
  return 0;
}
