// HBM-bound kernels of the cross_fusion path: patchify / fold layout passes, language-row
// scatter, LayerNorm forward / backward, column sums (bias grads), casts with head padding.
// All are vectorised (128-bit where the layout allows), coalesced on both the read and the write
// side (shared-memory staging for the layout permutations), and sized as multiples of the SM count.
#include <string.h>

#include "../../include/xfusion.h"
#include "host_common.cuh"
#include "ptx.cuh"

namespace xf {

// ------------------------------------------------------------------------------------------
// patchify / fold.  Token matrix T[(b,i,j), (c,u,v)]  <->  feature map F[b,c,i*p+u,j*p+v]
// (reference: nn.Conv2d(k=stride=p) im2col, cross_f_box_wrapper.py:268-274 + utils.py:35-39;
//  inverse: utils.py:42-46 regroup_patches / F.fold).
// A CTA moves a tile of 32 tokens (along j) x 128 columns through shared memory so that the
// NCHW side is accessed in runs along w and the token side in runs along k.
// ------------------------------------------------------------------------------------------
constexpr int PF_TJ = 32;
constexpr int PF_TK = 128;

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16(v); }

template <typename FT, int P, bool FOLD, bool ACC>
__global__ void __launch_bounds__(256)
patchify_fold_kernel(FT* __restrict__ feat, __nv_bfloat16* __restrict__ tok, int B, int Cc, int H, int W, long long tok_ld) {
  __shared__ float tile[PF_TJ][PF_TK + 1];
  constexpr int PP = P * P;
  constexpr int WRUN = PF_TJ * P;  // contiguous elements along w per (c,u)
  const int gh = H / P, gw = W / P;
  const int K = Cc * PP;
  const int jt = (gw + PF_TJ - 1) / PF_TJ;
  int bid = blockIdx.x;
  const int j0 = (bid % jt) * PF_TJ; bid /= jt;
  const int i = bid % gh; bid /= gh;
  const int b = bid;
  const int k0 = blockIdx.y * PF_TK;
  const int c0 = k0 / PP;
  const long long row0 = (static_cast<long long>(b) * gh + i) * gw + j0;
  constexpr int TOTAL = PF_TJ * PF_TK;

  if (!FOLD) {
#pragma unroll 4
    for (int t = threadIdx.x; t < TOTAL; t += 256) {
      const int wl = t % WRUN;
      const int cu = t / WRUN;
      const int u = cu % P, cl = cu / P;
      const int jl = wl / P, v = wl % P;
      const int c = c0 + cl, j = j0 + jl;
      float x = 0.f;
      if (c < Cc && j < gw) x = to_f32<FT>(feat[((static_cast<long long>(b) * Cc + c) * H + i * P + u) * W + j * P + v]);
      tile[jl][cl * PP + u * P + v] = x;
    }
    __syncthreads();
#pragma unroll 4
    for (int t = threadIdx.x; t < TOTAL / 2; t += 256) {
      const int kl = (t % (PF_TK / 2)) * 2, jl = t / (PF_TK / 2);
      if (j0 + jl < gw && k0 + kl < K)
        *reinterpret_cast<uint32_t*>(tok + (row0 + jl) * tok_ld + k0 + kl) = pack_bf16(tile[jl][kl], tile[jl][kl + 1]);
    }
  } else {
#pragma unroll 4
    for (int t = threadIdx.x; t < TOTAL / 2; t += 256) {
      const int kl = (t % (PF_TK / 2)) * 2, jl = t / (PF_TK / 2);
      uint32_t q = 0;
      if (j0 + jl < gw && k0 + kl < K) q = *reinterpret_cast<const uint32_t*>(tok + (row0 + jl) * tok_ld + k0 + kl);
      tile[jl][kl] = bf16_lo(q);
      tile[jl][kl + 1] = bf16_hi(q);
    }
    __syncthreads();
#pragma unroll 4
    for (int t = threadIdx.x; t < TOTAL; t += 256) {
      const int wl = t % WRUN;
      const int cu = t / WRUN;
      const int u = cu % P, cl = cu / P;
      const int jl = wl / P, v = wl % P;
      const int c = c0 + cl, j = j0 + jl;
      if (c < Cc && j < gw) {
        FT* dst = feat + ((static_cast<long long>(b) * Cc + c) * H + i * P + u) * W + j * P + v;
        const float x = tile[jl][cl * PP + u * P + v];
        if (ACC) *dst = from_f32<FT>(to_f32<FT>(*dst) + x);
        else *dst = from_f32<FT>(x);
      }
    }
  }
}

// 128-bit variant for fp32 feature maps with P >= 4 (the two fine FPN levels hold 94 % of the elements): the scalar
// kernel above spends ~15 instructions per element on index arithmetic and 32-bit accesses.  Same tile; the
// feature side moves float4 runs along w (one instruction = 4 consecutive pixels of one patch row), the token
// side 4 bf16 per thread (a warp writes 256 contiguous bytes of one token row).  Tile rows are 132 floats apart
// so that both access patterns are 16-byte aligned and (for P = 4) bank-conflict free.
template <int P, bool FOLD, bool ACC>
__global__ void __launch_bounds__(256)
patchify_fold_vec4_kernel(float* __restrict__ feat, __nv_bfloat16* __restrict__ tok, int B, int Cc, int H, int W, long long tok_ld) {
  constexpr int LDT = PF_TK + 4;
  __shared__ __align__(16) float tile[PF_TJ][LDT];
  constexpr int PP = P * P;
  constexpr int WRUN4 = PF_TJ * P / 4;  // float4 per (c, u) run along w
  const int gh = H / P, gw = W / P;
  const int K = Cc * PP;
  const int jt = (gw + PF_TJ - 1) / PF_TJ;
  int bid = blockIdx.x;
  const int j0 = (bid % jt) * PF_TJ; bid /= jt;
  const int i = bid % gh; bid /= gh;
  const int b = bid;
  const int k0 = blockIdx.y * PF_TK;
  const int c0 = k0 / PP;
  const long long row0 = (static_cast<long long>(b) * gh + i) * gw + j0;
  constexpr int UNITS = PF_TJ * PF_TK / 4;   // 1024 float4 units per tile

  auto feat_unit = [&](int t, int& jl, int& col, long long& off, bool& ok) {
    const int w4 = t % WRUN4, cu = t / WRUN4;
    const int u = cu % P, cl = cu / P;
    jl = (w4 * 4) / P;
    const int v0 = (w4 * 4) % P;
    col = cl * PP + u * P + v0;
    const int c = c0 + cl, j = j0 + jl;
    ok = c < Cc && j < gw;
    off = ((static_cast<long long>(b) * Cc + c) * H + i * P + u) * W + j * P + v0;
  };
  if (!FOLD) {
#pragma unroll
    for (int t = threadIdx.x; t < UNITS; t += 256) {
      int jl, col; long long off; bool ok;
      feat_unit(t, jl, col, off, ok);
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ok) x = __ldg(reinterpret_cast<const float4*>(feat + off));
      *reinterpret_cast<float4*>(&tile[jl][col]) = x;
    }
    __syncthreads();
#pragma unroll
    for (int t = threadIdx.x; t < UNITS; t += 256) {
      const int k4 = (t % (PF_TK / 4)) * 4, jl = t / (PF_TK / 4);
      if (j0 + jl < gw && k0 + k4 < K) {
        const float4 x = *reinterpret_cast<const float4*>(&tile[jl][k4]);
        *reinterpret_cast<uint2*>(tok + (row0 + jl) * tok_ld + k0 + k4) = make_uint2(pack_bf16(x.x, x.y), pack_bf16(x.z, x.w));
      }
    }
  } else {
#pragma unroll
    for (int t = threadIdx.x; t < UNITS; t += 256) {
      const int k4 = (t % (PF_TK / 4)) * 4, jl = t / (PF_TK / 4);
      uint2 q = make_uint2(0u, 0u);
      if (j0 + jl < gw && k0 + k4 < K) q = *reinterpret_cast<const uint2*>(tok + (row0 + jl) * tok_ld + k0 + k4);
      *reinterpret_cast<float4*>(&tile[jl][k4]) = make_float4(bf16_lo(q.x), bf16_hi(q.x), bf16_lo(q.y), bf16_hi(q.y));
    }
    __syncthreads();
#pragma unroll
    for (int t = threadIdx.x; t < UNITS; t += 256) {
      int jl, col; long long off; bool ok;
      feat_unit(t, jl, col, off, ok);
      if (ok) {
        float4 x = *reinterpret_cast<const float4*>(&tile[jl][col]);
        float4* dst = reinterpret_cast<float4*>(feat + off);
        if (ACC) { const float4 o = *dst; x.x += o.x; x.y += o.y; x.z += o.z; x.w += o.w; }
        *dst = x;
      }
    }
  }
}

// bf16 feature maps, P >= 4: the same tile with 128-bit accesses on the feature side (8 consecutive pixels along w = 8 / P
// patch rows) and 64-bit on the token side; the scalar kernel moves 2 bytes per instruction on the feature side.
template <int P, bool FOLD, bool ACC>
__global__ void __launch_bounds__(256)
patchify_fold_vec8h_kernel(__nv_bfloat16* __restrict__ feat, __nv_bfloat16* __restrict__ tok, int B, int Cc, int H, int W, long long tok_ld) {
  constexpr int LDT = PF_TK + 4;
  __shared__ __align__(16) float tile[PF_TJ][LDT];
  constexpr int PP = P * P;
  constexpr int WRUN8 = PF_TJ * P / 8;  // 8-pixel units per (c, u) run along w
  constexpr int TPU = 8 / P;            // tokens touched by one unit (2 for P = 4, 1 for P = 8)
  const int gh = H / P, gw = W / P;
  const int K = Cc * PP;
  const int jt = (gw + PF_TJ - 1) / PF_TJ;
  int bid = blockIdx.x;
  const int j0 = (bid % jt) * PF_TJ; bid /= jt;
  const int i = bid % gh; bid /= gh;
  const int b = bid;
  const int k0 = blockIdx.y * PF_TK;
  const int c0 = k0 / PP;
  const long long row0 = (static_cast<long long>(b) * gh + i) * gw + j0;
  constexpr int UNITS8 = PF_TJ * PF_TK / 8;   // 512 feature-side units per tile
  constexpr int UNITS4 = PF_TJ * PF_TK / 4;   // 1024 token-side units per tile

  auto feat_unit = [&](int t, int& jl, int& col, long long& off, int& ntok) {
    const int w8 = t % WRUN8, cu = t / WRUN8;
    const int u = cu % P, cl = cu / P;
    jl = (w8 * 8) / P;
    const int v0 = (w8 * 8) % P;            // 0 for P = 4 and P = 8
    col = cl * PP + u * P + v0;
    const int c = c0 + cl;
    ntok = c < Cc ? min(TPU, gw - (j0 + jl)) : 0;   // tokens of this unit inside the grid (gw * P % 8 == 0: whole units)
    off = ((static_cast<long long>(b) * Cc + c) * H + i * P + u) * W + (j0 + jl) * P + v0;
  };
  if (!FOLD) {
#pragma unroll
    for (int t = threadIdx.x; t < UNITS8; t += 256) {
      int jl, col, ntok; long long off;
      feat_unit(t, jl, col, off, ntok);
      uint4 q = make_uint4(0u, 0u, 0u, 0u);
      if (ntok > 0) q = __ldg(reinterpret_cast<const uint4*>(feat + off));
      if (P == 4) {
        *reinterpret_cast<float4*>(&tile[jl][col]) = make_float4(bf16_lo(q.x), bf16_hi(q.x), bf16_lo(q.y), bf16_hi(q.y));
        if (jl + 1 < PF_TJ) *reinterpret_cast<float4*>(&tile[jl + 1][col]) = make_float4(bf16_lo(q.z), bf16_hi(q.z), bf16_lo(q.w), bf16_hi(q.w));
      } else {
        *reinterpret_cast<float4*>(&tile[jl][col]) = make_float4(bf16_lo(q.x), bf16_hi(q.x), bf16_lo(q.y), bf16_hi(q.y));
        *reinterpret_cast<float4*>(&tile[jl][col + 4]) = make_float4(bf16_lo(q.z), bf16_hi(q.z), bf16_lo(q.w), bf16_hi(q.w));
      }
    }
    __syncthreads();
#pragma unroll
    for (int t = threadIdx.x; t < UNITS4; t += 256) {
      const int k4 = (t % (PF_TK / 4)) * 4, jl = t / (PF_TK / 4);
      if (j0 + jl < gw && k0 + k4 < K) {
        const float4 x = *reinterpret_cast<const float4*>(&tile[jl][k4]);
        *reinterpret_cast<uint2*>(tok + (row0 + jl) * tok_ld + k0 + k4) = make_uint2(pack_bf16(x.x, x.y), pack_bf16(x.z, x.w));
      }
    }
  } else {
#pragma unroll
    for (int t = threadIdx.x; t < UNITS4; t += 256) {
      const int k4 = (t % (PF_TK / 4)) * 4, jl = t / (PF_TK / 4);
      uint2 q = make_uint2(0u, 0u);
      if (j0 + jl < gw && k0 + k4 < K) q = *reinterpret_cast<const uint2*>(tok + (row0 + jl) * tok_ld + k0 + k4);
      *reinterpret_cast<float4*>(&tile[jl][k4]) = make_float4(bf16_lo(q.x), bf16_hi(q.x), bf16_lo(q.y), bf16_hi(q.y));
    }
    __syncthreads();
#pragma unroll
    for (int t = threadIdx.x; t < UNITS8; t += 256) {
      int jl, col, ntok; long long off;
      feat_unit(t, jl, col, off, ntok);
      if (ntok > 0) {
        float4 a = *reinterpret_cast<const float4*>(&tile[jl][col]);
        float4 c = P == 4 ? (jl + 1 < PF_TJ ? *reinterpret_cast<const float4*>(&tile[jl + 1][col]) : make_float4(0.f, 0.f, 0.f, 0.f))
                          : *reinterpret_cast<const float4*>(&tile[jl][col + 4]);
        uint4* dst = reinterpret_cast<uint4*>(feat + off);
        if (ACC) {
          const uint4 o = *dst;
          a.x += bf16_lo(o.x); a.y += bf16_hi(o.x); a.z += bf16_lo(o.y); a.w += bf16_hi(o.y);
          c.x += bf16_lo(o.z); c.y += bf16_hi(o.z); c.z += bf16_lo(o.w); c.w += bf16_hi(o.w);
        }
        *dst = make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(c.x, c.y), pack_bf16(c.z, c.w));
      }
    }
  }
}

// channels_last (NHWC memory) feature maps -- SURVEY 8f N3: a channels_last backbone feeds the patch-embed without an
// NCHW conversion pass, and the fused map goes back out in channels_last for the detector's cuDNN convolutions.
//   tok[(b,i,j), c*p*p + u*p + v]  <->  F_nhwc[b, i*p+u, j*p+v, c]
// A CTA moves 32 tokens (along j) x 32 channels x p*p positions through shared memory: the feature side is accessed in
// runs along c, the token side in runs of 32*p*p consecutive k.
template <typename FT, int P, bool FOLD>
__global__ void __launch_bounds__(256)
patchify_fold_nhwc_kernel(FT* __restrict__ feat, __nv_bfloat16* __restrict__ tok, int B, int Cc, int H, int W, long long tok_ld) {
  constexpr int PP = P * P;
  constexpr int TC = 32;                      // channels per tile
  constexpr int TJ = P >= 4 ? 16 : 32;        // tokens per tile (static shared memory stays under 48 KB)
  __shared__ float tile[TJ][TC][PP + 1];      // +1: conflict-free transposed access
  const int gh = H / P, gw = W / P;
  const int jt = (gw + TJ - 1) / TJ;
  int bid = blockIdx.x;
  const int j0 = (bid % jt) * TJ; bid /= jt;
  const int i = bid % gh; bid /= gh;
  const int b = bid;
  const int c0 = blockIdx.y * TC;
  const long long row0 = (static_cast<long long>(b) * gh + i) * gw + j0;
  constexpr int TOTAL = TJ * TC * PP;
  auto feat_at = [&](int jl, int cl, int uv) -> FT* {
    const int u = uv / P, v = uv % P;
    return feat + ((static_cast<long long>(b) * H + i * P + u) * W + (j0 + jl) * P + v) * Cc + c0 + cl;
  };
  if (!FOLD) {
    for (int t = threadIdx.x; t < TOTAL; t += 256) {      // c fastest: coalesced along the channel axis
      const int cl = t % TC, uv = (t / TC) % PP, jl = t / (TC * PP);
      float x = 0.f;
      if (j0 + jl < gw && c0 + cl < Cc) x = to_f32<FT>(*feat_at(jl, cl, uv));
      tile[jl][cl][uv] = x;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < TOTAL; t += 256) {      // k fastest: 32 * p*p consecutive columns of a token row
      const int kl = t % (TC * PP), jl = t / (TC * PP);
      const int cl = kl / PP, uv = kl % PP;
      if (j0 + jl < gw && c0 + cl < Cc) tok[(row0 + jl) * tok_ld + static_cast<long long>(c0) * PP + kl] = __float2bfloat16(tile[jl][cl][uv]);
    }
  } else {
    for (int t = threadIdx.x; t < TOTAL; t += 256) {
      const int kl = t % (TC * PP), jl = t / (TC * PP);
      const int cl = kl / PP, uv = kl % PP;
      float x = 0.f;
      if (j0 + jl < gw && c0 + cl < Cc) x = __bfloat162float(tok[(row0 + jl) * tok_ld + static_cast<long long>(c0) * PP + kl]);
      tile[jl][cl][uv] = x;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < TOTAL; t += 256) {
      const int cl = t % TC, uv = (t / TC) % PP, jl = t / (TC * PP);
      if (j0 + jl < gw && c0 + cl < Cc) *feat_at(jl, cl, uv) = from_f32<FT>(tile[jl][cl][uv]);
    }
  }
}

template <typename FT, bool FOLD>
static int launch_patchify_fold_nhwc(FT* feat, __nv_bfloat16* tok, int B, int C, int H, int W, int p, long long tok_ld, cudaStream_t st) {
  const int gh = H / p, gw = W / p;
  const int tj = p >= 4 ? 16 : 32;
  dim3 grid(B * gh * ((gw + tj - 1) / tj), (C + 31) / 32);
  switch (p) {
    case 1: patchify_fold_nhwc_kernel<FT, 1, FOLD><<<grid, 256, 0, st>>>(feat, tok, B, C, H, W, tok_ld); break;
    case 2: patchify_fold_nhwc_kernel<FT, 2, FOLD><<<grid, 256, 0, st>>>(feat, tok, B, C, H, W, tok_ld); break;
    case 4: patchify_fold_nhwc_kernel<FT, 4, FOLD><<<grid, 256, 0, st>>>(feat, tok, B, C, H, W, tok_ld); break;
    default: return fail(-2, "channels_last patchify / fold supports patch sizes 1, 2, 4");
  }
  return 0;
}

template <typename FT, bool FOLD, bool ACC>
static void launch_patchify_fold(FT* feat, __nv_bfloat16* tok, int B, int C, int H, int W, int p, long long tok_ld, cudaStream_t st) {
  const int gh = H / p, gw = W / p, K = C * p * p;
  dim3 grid(B * gh * ((gw + PF_TJ - 1) / PF_TJ), (K + PF_TK - 1) / PF_TK);
  if constexpr (sizeof(FT) == 4) {
    const bool vec = (p == 4 || p == 8) && W % 4 == 0 && (reinterpret_cast<uintptr_t>(feat) & 15) == 0 && tok_ld % 4 == 0 &&
                     (reinterpret_cast<uintptr_t>(tok) & 7) == 0 && K % 4 == 0;
    if (vec) {
      if (p == 4) patchify_fold_vec4_kernel<4, FOLD, ACC><<<grid, 256, 0, st>>>(reinterpret_cast<float*>(feat), tok, B, C, H, W, tok_ld);
      else patchify_fold_vec4_kernel<8, FOLD, ACC><<<grid, 256, 0, st>>>(reinterpret_cast<float*>(feat), tok, B, C, H, W, tok_ld);
      return;
    }
  } else {
    // W % 8 == 0 keeps every 8-pixel unit inside one image row and 16-byte aligned; gw * p % 8 == 0 follows
    const bool vec = (p == 4 || p == 8) && W % 8 == 0 && (reinterpret_cast<uintptr_t>(feat) & 15) == 0 && tok_ld % 4 == 0 &&
                     (reinterpret_cast<uintptr_t>(tok) & 7) == 0 && K % 4 == 0;
    if (vec) {
      if (p == 4) patchify_fold_vec8h_kernel<4, FOLD, ACC><<<grid, 256, 0, st>>>(reinterpret_cast<__nv_bfloat16*>(feat), tok, B, C, H, W, tok_ld);
      else patchify_fold_vec8h_kernel<8, FOLD, ACC><<<grid, 256, 0, st>>>(reinterpret_cast<__nv_bfloat16*>(feat), tok, B, C, H, W, tok_ld);
      return;
    }
  }
  switch (p) {
    case 1: patchify_fold_kernel<FT, 1, FOLD, ACC><<<grid, 256, 0, st>>>(feat, tok, B, C, H, W, tok_ld); break;
    case 2: patchify_fold_kernel<FT, 2, FOLD, ACC><<<grid, 256, 0, st>>>(feat, tok, B, C, H, W, tok_ld); break;
    case 4: patchify_fold_kernel<FT, 4, FOLD, ACC><<<grid, 256, 0, st>>>(feat, tok, B, C, H, W, tok_ld); break;
    default: patchify_fold_kernel<FT, 8, FOLD, ACC><<<grid, 256, 0, st>>>(feat, tok, B, C, H, W, tok_ld); break;
  }
}

// ------------------------------------------------------------------------------------------
// language rows: z[b, n + j, :] = bf16(lang[b,j,:] + kind[:])   (cross_f_box_layers.py:76,86)
// backward:      dlang[b,j,:] += dz[b, n+j, :] ;  dkind[:] += sum_{b,j} dz[b,n+j,:]
// ------------------------------------------------------------------------------------------
__global__ void lang_rows_fwd_kernel(const float* __restrict__ lang, const float* __restrict__ kind,
                                     __nv_bfloat16* __restrict__ z, int B, int L, int D, int n, int S) {
  const long long total = static_cast<long long>(B) * L * (D / 2);
  for (long long t = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; t < total;
       t += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int e2 = t % (D / 2);
    const long long r = t / (D / 2);
    const int j = r % L, b = r / L;
    const float2 x = *reinterpret_cast<const float2*>(lang + r * D + 2 * e2);
    const float2 k = *reinterpret_cast<const float2*>(kind + 2 * e2);
    *reinterpret_cast<uint32_t*>(z + (static_cast<long long>(b) * S + n + j) * D + 2 * e2) = pack_bf16(x.x + k.x, x.y + k.y);
  }
}

// grid.x: blocks of 128 column pairs, grid.y: chunks of rows; every (row, column pair) has one owner thread, the
// kind-embedding gradient is reduced per thread over its row chunk, then one atomic pair per thread.
__global__ void __launch_bounds__(128)
lang_rows_bwd_kernel(const __nv_bfloat16* __restrict__ dz, float* __restrict__ dlang, float* __restrict__ dkind, int B, int L,
                     int D, int n, int S, int rows_per_cta) {
  const int e2 = blockIdx.x * blockDim.x + threadIdx.x;
  if (e2 >= D / 2) return;
  const int r0 = blockIdx.y * rows_per_cta;
  const int r1 = min(B * L, r0 + rows_per_cta);
  float s0 = 0.f, s1 = 0.f;
  for (int r = r0; r < r1; ++r) {
    const int j = r % L, b = r / L;
    const uint32_t q = *reinterpret_cast<const uint32_t*>(dz + (static_cast<long long>(b) * S + n + j) * D + 2 * e2);
    const float g0 = bf16_lo(q), g1 = bf16_hi(q);
    if (dlang) {
      float2* d = reinterpret_cast<float2*>(dlang + static_cast<long long>(r) * D + 2 * e2);
      float2 cur = *d;
      cur.x += g0; cur.y += g1;
      *d = cur;
    }
    s0 += g0; s1 += g1;
  }
  atomicAdd(dkind + 2 * e2, s0);
  atomicAdd(dkind + 2 * e2 + 1, s1);
}

// ------------------------------------------------------------------------------------------
// LayerNorm over the last dim (eps 1e-5, affine), one warp per row, row cached in registers.
//   y[orow] = (x[irow] - mean) * rstd * gamma + beta ;  irow / orow support the same block remap
//   as the GEMM epilogue so the final LN reads only the visual rows of z (cross_f_box_layers.py:
//   104-107) and writes a compact [B*n, D] matrix.
// ------------------------------------------------------------------------------------------
constexpr int LN_MAXV = 8;  // up to 8 x (32 lanes x 8 elements) = 2048 columns

struct LnParams {
  const __nv_bfloat16* x; long long ldx;
  __nv_bfloat16* y; long long ldy;
  const float* gamma; const float* beta;
  float* mean; float* rstd;   // [rows], indexed by logical row
  int rows, D;
  int in_rows_in, in_rows_out, in_row_off;     // logical row r -> x row
  int out_rows_in, out_rows_out, out_row_off;  // logical row r -> y row
  float eps;
  // dropout on the output (backproj_dropout, utils.py:115)
  float drop_p; uint32_t drop_seed, drop_stream, drop_thresh; float drop_scale;
  const uint32_t* coltab;   // xf::drop_col_table()
};

__device__ __forceinline__ long long remap_row(int r, int rin, int rout, int off) {
  if (rin <= 0) return r;
  const int g = r / rin;
  return static_cast<long long>(g) * rout + (r - g * rin) + off;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ void unpack8(const uint4& q, float (&v)[8]) {
  v[0] = bf16_lo(q.x); v[1] = bf16_hi(q.x); v[2] = bf16_lo(q.y); v[3] = bf16_hi(q.y);
  v[4] = bf16_lo(q.z); v[5] = bf16_hi(q.z); v[6] = bf16_lo(q.w); v[7] = bf16_hi(q.w);
}
__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
  return make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
}
__device__ __forceinline__ void load8f(const float* p, float (&v)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
// dropout on 8 consecutive columns col0.. (col0 % 8 == 0) of the row with hash rh; tab = xf::drop_col_table()
__device__ __forceinline__ void drop8(float (&v)[8], uint32_t rh, uint32_t col0, uint32_t t32, float scale, const uint32_t* __restrict__ tab) {
  const uint4 c0 = __ldg(reinterpret_cast<const uint4*>(tab + col0));
  const uint4 c1 = __ldg(reinterpret_cast<const uint4*>(tab + col0) + 1);
  const uint32_t ch[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
  for (int k = 0; k < 8; ++k) v[k] = drop_keep_rc(rh, ch[k], t32) ? v[k] * scale : 0.f;
}

// NV = 16-byte vectors per lane (row cached in registers as packed bf16)
template <int NV>
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const LnParams p) {
  const int warps_per_cta = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nvec = p.D >> 3;
  for (int r = blockIdx.x * warps_per_cta + (threadIdx.x >> 5); r < p.rows; r += gridDim.x * warps_per_cta) {
    const __nv_bfloat16* xr = p.x + remap_row(r, p.in_rows_in, p.in_rows_out, p.in_row_off) * p.ldx;
    uint4 q[NV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int vidx = lane + 32 * i;
      if (vidx < nvec) {
        q[i] = __ldg(reinterpret_cast<const uint4*>(xr) + vidx);
        float v[8]; unpack8(q[i], v);
#pragma unroll
        for (int k = 0; k < 8; ++k) s += v[k];
      }
    }
    const float mean = warp_sum(s) / p.D;
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int vidx = lane + 32 * i;
      if (vidx < nvec) {
        float v[8]; unpack8(q[i], v);
#pragma unroll
        for (int k = 0; k < 8; ++k) { const float a = v[k] - mean; ss += a * a; }
      }
    }
    const float rstd = rsqrtf(warp_sum(ss) / p.D + p.eps);
    if (lane == 0 && p.mean) { p.mean[r] = mean; p.rstd[r] = rstd; }
    const long long orow = remap_row(r, p.out_rows_in, p.out_rows_out, p.out_row_off);
    __nv_bfloat16* yr = p.y + orow * p.ldy;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int vidx = lane + 32 * i;
      if (vidx < nvec) {
        float v[8], g[8], b[8];
        unpack8(q[i], v);
        load8f(p.gamma + vidx * 8, g);
        load8f(p.beta + vidx * 8, b);
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = (v[k] - mean) * rstd * g[k] + b[k];
        if (p.drop_p > 0.f) drop8(v, drop_rowhash(p.drop_seed, static_cast<uint64_t>(orow)), vidx * 8, p.drop_thresh, p.drop_scale, p.coltab);
        *(reinterpret_cast<uint4*>(yr) + vidx) = pack8(v);
      }
    }
  }
}

// fp32 += of 4 consecutive columns: a 128-bit vector reduction when the address is 16-byte aligned, else scalar atomics
__device__ __forceinline__ void red_add4(float* dst, float2 a, float2 b) {
  if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y) : "memory");
  } else {
    atomicAdd(dst, a.x); atomicAdd(dst + 1, a.y); atomicAdd(dst + 2, b.x); atomicAdd(dst + 3, b.y);
  }
}

// LayerNorm backward.  dx = rstd * (g - mean(g) - xhat * mean(g*xhat)),  g = dy * gamma.
// Also: dgamma += sum dy*xhat, dbeta += sum dy (fp32) and, optionally, colsum of the gradient that
// enters the producing linear (= its bias grad) and a dropout-masked copy dx2 of dx (the gradient of
// that linear's output when dropout1/2 are on).  Rows are cached as packed bf16; per-lane column
// accumulators are combined through shared memory, then one global atomic per column per CTA.
struct LnBwdParams {
  const __nv_bfloat16* dy; long long lddy;     // indexed by "out" remap of the forward
  const __nv_bfloat16* x; long long ldx;       // forward input (pre-LN), "in" remap
  const float* gamma; const float* mean; const float* rstd;
  __nv_bfloat16* dx; long long lddx;           // "in" remap
  __nv_bfloat16* dx2;                          // optional, same indexing as dx
  float* dgamma; float* dbeta; float* dbias;   // [D] fp32, atomically accumulated; dbias optional
  int rows, D;
  int in_rows_in, in_rows_out, in_row_off;
  int out_rows_in, out_rows_out, out_row_off;
  float dy_drop_p; uint32_t dy_seed, dy_stream, dy_thresh; float dy_scale;
  float dx2_drop_p; uint32_t dx2_seed, dx2_stream, dx2_thresh; float dx2_scale;
  const uint32_t* coltab;   // xf::drop_col_table()
};

// Phase 1 (row-wise, one warp per row, row cached as packed bf16): dx (+ dropout-masked copy dx2).
// Phase 2 (column-wise, one thread per 4 columns): the CTA re-reads the rows it just processed (L2-hot)
// and accumulates dgamma / dbeta / dbias column sums in registers, then one atomic per column per CTA.
// Splitting the phases keeps the register count low (the fused version needed 165 regs and ran at
// 1.1 TB/s) so several CTAs per SM keep enough loads in flight.
constexpr int LNB_ROWS = 32;  // rows per CTA pass

// packed helpers: 8 bf16 (uint4) <-> 4 float2
__device__ __forceinline__ void unpack8x2(const uint4& q, float2 (&v)[4]) {
  v[0] = make_float2(bf16_lo(q.x), bf16_hi(q.x)); v[1] = make_float2(bf16_lo(q.y), bf16_hi(q.y));
  v[2] = make_float2(bf16_lo(q.z), bf16_hi(q.z)); v[3] = make_float2(bf16_lo(q.w), bf16_hi(q.w));
}
__device__ __forceinline__ uint4 pack8x2(const float2 (&v)[4]) {
  return make_uint4(pack_bf16(v[0].x, v[0].y), pack_bf16(v[1].x, v[1].y), pack_bf16(v[2].x, v[2].y), pack_bf16(v[3].x, v[3].y));
}
// dropout mask (no scaling) on 8 consecutive columns col0..; tab = xf::drop_col_table()
__device__ __forceinline__ void dropmask8x2(float2 (&v)[4], uint32_t rh, uint32_t col0, uint32_t t32, const uint32_t* __restrict__ tab) {
  const uint4 c0 = __ldg(reinterpret_cast<const uint4*>(tab + col0));
  const uint4 c1 = __ldg(reinterpret_cast<const uint4*>(tab + col0) + 1);
  if (!drop_keep_rc(rh, c0.x, t32)) v[0].x = 0.f;
  if (!drop_keep_rc(rh, c0.y, t32)) v[0].y = 0.f;
  if (!drop_keep_rc(rh, c0.z, t32)) v[1].x = 0.f;
  if (!drop_keep_rc(rh, c0.w, t32)) v[1].y = 0.f;
  if (!drop_keep_rc(rh, c1.x, t32)) v[2].x = 0.f;
  if (!drop_keep_rc(rh, c1.y, t32)) v[2].y = 0.f;
  if (!drop_keep_rc(rh, c1.z, t32)) v[3].x = 0.f;
  if (!drop_keep_rc(rh, c1.w, t32)) v[3].y = 0.f;
}

// The kernel is instruction-bound (it ran at 43 instructions per element), so both phases use packed fp32 math
// and the row phase uses the affine form of the gradient:
//   g = dy * gamma,  s1 = mean_c(g),  s2 = mean_c(g * xhat) = rstd * (mean_c(g * x) - mean * s1)
//   dx = rstd * (g - s1 - xhat * s2) = dy * (rstd * gamma) + x * b + c,   b = -rstd^2 s2,  c = -rstd s1 - b mean
template <int NV>
__global__ void __launch_bounds__(256, 2) layernorm_bwd_kernel(const LnBwdParams p) {
  const int warps_per_cta = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nvec = p.D >> 3;
  const int c4 = threadIdx.x * 4;                 // phase-2 columns of this thread (+ 1024 * j)
  constexpr int NJ = (NV * 256 + 1023) / 1024;    // column slots of 1024 per thread
  float2 cg[NJ][2], cb[NJ][2], cbias[NJ][2];
#pragma unroll
  for (int j = 0; j < NJ; ++j)
#pragma unroll
    for (int k = 0; k < 2; ++k) cg[j][k] = cb[j][k] = cbias[j][k] = make_float2(0.f, 0.f);
  const float inv_d = 1.f / p.D;
  const bool dy_drop = p.dy_drop_p > 0.f, dx2_drop = p.dx2_drop_p > 0.f;

  for (int base = blockIdx.x * LNB_ROWS; base < p.rows; base += gridDim.x * LNB_ROWS) {
    const int rend = min(p.rows, base + LNB_ROWS);
    // ---------------- phase 1: one warp per row, row cached as packed bf16
    for (int r = base + warp; r < rend; r += warps_per_cta) {
      const long long irow = remap_row(r, p.in_rows_in, p.in_rows_out, p.in_row_off);
      const long long orow = remap_row(r, p.out_rows_in, p.out_rows_out, p.out_row_off);
      const __nv_bfloat16* xr = p.x + irow * p.ldx;
      const __nv_bfloat16* dyr = p.dy + orow * p.lddy;
      const float mean = p.mean[r], rstd = p.rstd[r];
      const uint32_t rh_dy = dy_drop ? drop_rowhash(p.dy_seed, static_cast<uint64_t>(orow)) : 0u;
      uint4 qx[NV], qd[NV];
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int vidx = lane + 32 * i;
        if (vidx < nvec) {
          qx[i] = __ldg(reinterpret_cast<const uint4*>(xr) + vidx);
          qd[i] = __ldg(reinterpret_cast<const uint4*>(dyr) + vidx);
        }
      }
      float2 s1v = make_float2(0.f, 0.f), tv = make_float2(0.f, 0.f);
      const float2 dys = make_float2(p.dy_scale, p.dy_scale);
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int vidx = lane + 32 * i;
        if (vidx < nvec) {
          float2 xv[4], dv[4];
          unpack8x2(qx[i], xv); unpack8x2(qd[i], dv);
          if (dy_drop) {   // the incoming gradient passes the forward's dropout mask first; keep the masked copy
            dropmask8x2(dv, rh_dy, vidx * 8, p.dy_thresh, p.coltab);
#pragma unroll
            for (int k = 0; k < 4; ++k) dv[k] = __fmul2_rn(dv[k], dys);
            qd[i] = pack8x2(dv);
          }
          const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.gamma + vidx * 8));
          const float4 g1 = __ldg(reinterpret_cast<const float4*>(p.gamma + vidx * 8) + 1);
          const float2 gm[4] = {make_float2(g0.x, g0.y), make_float2(g0.z, g0.w), make_float2(g1.x, g1.y), make_float2(g1.z, g1.w)};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float2 g = __fmul2_rn(dv[k], gm[k]);
            s1v = __fadd2_rn(s1v, g);
            tv = __ffma2_rn(g, xv[k], tv);
          }
        }
      }
      const float s1 = warp_sum(s1v.x + s1v.y) * inv_d;
      const float t = warp_sum(tv.x + tv.y) * inv_d;
      const float s2 = rstd * (t - mean * s1);
      const float bcoef = -rstd * rstd * s2, ccoef = -rstd * s1 - bcoef * mean;
      const float2 b2 = make_float2(bcoef, bcoef), c2 = make_float2(ccoef, ccoef), r2 = make_float2(rstd, rstd);
      __nv_bfloat16* dxr = p.dx + irow * p.lddx;
      const uint32_t rh_dx2 = dx2_drop ? drop_rowhash(p.dx2_seed, static_cast<uint64_t>(irow)) : 0u;
      const float2 dx2s = make_float2(p.dx2_scale, p.dx2_scale);
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int vidx = lane + 32 * i;
        if (vidx < nvec) {
          float2 xv[4], dv[4], o[4];
          unpack8x2(qx[i], xv); unpack8x2(qd[i], dv);
          const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.gamma + vidx * 8));
          const float4 g1 = __ldg(reinterpret_cast<const float4*>(p.gamma + vidx * 8) + 1);
          const float2 gm[4] = {make_float2(g0.x, g0.y), make_float2(g0.z, g0.w), make_float2(g1.x, g1.y), make_float2(g1.z, g1.w)};
#pragma unroll
          for (int k = 0; k < 4; ++k) o[k] = __ffma2_rn(dv[k], __fmul2_rn(gm[k], r2), __ffma2_rn(xv[k], b2, c2));
          *(reinterpret_cast<uint4*>(dxr) + vidx) = pack8x2(o);
          if (p.dx2) {
            if (dx2_drop) {
              dropmask8x2(o, rh_dx2, vidx * 8, p.dx2_thresh, p.coltab);
#pragma unroll
              for (int k = 0; k < 4; ++k) o[k] = __fmul2_rn(o[k], dx2s);
            }
            *(reinterpret_cast<uint4*>(p.dx2 + irow * p.lddx) + vidx) = pack8x2(o);
          }
        }
      }
    }
    __syncthreads();   // dx / dx2 of this row block are written (visible CTA-wide after the barrier)
    // ---------------- phase 2: column sums over rows [base, rend); the rows are re-read (L1 / L2 hits)
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int col = c4 + 1024 * j;
      if (col < p.D) {
        uint4 ch = make_uint4(0u, 0u, 0u, 0u);
        if (dy_drop) ch = __ldg(reinterpret_cast<const uint4*>(p.coltab + col));
        for (int rb = base; rb < rend; rb += 4) {   // 4 rows (12 independent 64-bit loads) in flight per thread
          uint2 qx[4], qd[4], qo[4];
          float rs[4], nmr[4];
          long long orow[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int r = min(rb + u, rend - 1);
            const long long irow = remap_row(r, p.in_rows_in, p.in_rows_out, p.in_row_off);
            orow[u] = remap_row(r, p.out_rows_in, p.out_rows_out, p.out_row_off);
            rs[u] = p.rstd[r]; nmr[u] = -p.mean[r] * rs[u];
            qx[u] = *reinterpret_cast<const uint2*>(p.x + irow * p.ldx + col);
            qd[u] = *reinterpret_cast<const uint2*>(p.dy + orow[u] * p.lddy + col);
            if (p.dbias) qo[u] = *reinterpret_cast<const uint2*>((p.dx2 ? p.dx2 : p.dx) + irow * p.lddx + col);
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (rb + u < rend) {
              float2 xv[2] = {make_float2(bf16_lo(qx[u].x), bf16_hi(qx[u].x)), make_float2(bf16_lo(qx[u].y), bf16_hi(qx[u].y))};
              float2 dv[2] = {make_float2(bf16_lo(qd[u].x), bf16_hi(qd[u].x)), make_float2(bf16_lo(qd[u].y), bf16_hi(qd[u].y))};
              if (dy_drop) {
                const uint32_t rh = drop_rowhash(p.dy_seed, static_cast<uint64_t>(orow[u]));
                if (!drop_keep_rc(rh, ch.x, p.dy_thresh)) dv[0].x = 0.f;
                if (!drop_keep_rc(rh, ch.y, p.dy_thresh)) dv[0].y = 0.f;
                if (!drop_keep_rc(rh, ch.z, p.dy_thresh)) dv[1].x = 0.f;
                if (!drop_keep_rc(rh, ch.w, p.dy_thresh)) dv[1].y = 0.f;
                dv[0] = __fmul2_rn(dv[0], make_float2(p.dy_scale, p.dy_scale));
                dv[1] = __fmul2_rn(dv[1], make_float2(p.dy_scale, p.dy_scale));
              }
              const float2 r2 = make_float2(rs[u], rs[u]), n2 = make_float2(nmr[u], nmr[u]);
#pragma unroll
              for (int k = 0; k < 2; ++k) {
                cg[j][k] = __ffma2_rn(dv[k], __ffma2_rn(xv[k], r2, n2), cg[j][k]);   // dy * xhat
                cb[j][k] = __fadd2_rn(cb[j][k], dv[k]);
              }
              if (p.dbias) {
                cbias[j][0] = __fadd2_rn(cbias[j][0], make_float2(bf16_lo(qo[u].x), bf16_hi(qo[u].x)));
                cbias[j][1] = __fadd2_rn(cbias[j][1], make_float2(bf16_lo(qo[u].y), bf16_hi(qo[u].y)));
              }
            }
          }
        }
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const int col = c4 + 1024 * j;
    if (col < p.D) {
      // one 128-bit reduction per array instead of four scalar atomics (hundreds of CTAs hit the same D columns)
      red_add4(p.dgamma + col, cg[j][0], cg[j][1]);
      red_add4(p.dbeta + col, cb[j][0], cb[j][1]);
      if (p.dbias) red_add4(p.dbias + col, cbias[j][0], cbias[j][1]);
    }
  }
}

// ---- TMA-pipelined variant (no row remap) --------------------------------------------------------------------
// The kernel above is latency-bound: each warp serialises load -> reduce -> compute -> store for its row and the
// column phase re-reads the rows from L2.  Here blocks of 8 rows (x and dy) arrive in a 3-stage shared-memory
// ring by TMA, two blocks ahead of the math; both phases read shared memory, dx2 (the operand of the bias-gradient
// column sum) is staged in shared memory too, and global memory is touched exactly once per element.
// Stage layout = TMA boxes [chunk j][8 rows][bc columns] (bc = box width, a divisor of D).
constexpr int LNT_ROWS = 8, LNT_STAGES = 3;

// explicit shared-space accesses (32-bit shared addresses): the compiler cannot prove the address space of the
// aligned dynamic shared-memory pointer and would emit generic loads with 64-bit address arithmetic
__device__ __forceinline__ uint4 lds128(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ uint2 lds64(uint32_t a) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ float lds32f(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t a, const uint4& v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

template <int NV>
__global__ void __launch_bounds__(256, 2)
layernorm_bwd_tma_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_dy,
                         const __grid_constant__ LnBwdParams p, const int bc) {
  extern __shared__ uint8_t lnt_smem_raw[];
  const uint32_t smem = (smem_u32(lnt_smem_raw) + 127u) & ~127u;   // shared-space address of the aligned block
  const uint32_t bar0 = smem;                                       // full[3]
  const uint32_t s_rstd = smem + 32, s_nmr = smem + 64;             // [8] each; nmr = -mean * rstd
  const uint32_t row_bytes = p.D * 2u, half_stage = LNT_ROWS * row_bytes, stage_bytes = 2 * half_stage;
  const uint32_t ring = smem + 128;
  const uint32_t s_dx2 = ring + LNT_STAGES * stage_bytes;           // [8][D] bf16, row-major

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nvec = p.D >> 3, vpc = bc >> 3;
  const int nblk = (p.rows + LNT_ROWS - 1) / LNT_ROWS;
  const int nmine = blockIdx.x < nblk ? (nblk - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_x); tma_prefetch_desc(&tmap_dy);
    for (int s = 0; s < LNT_STAGES; ++s) mbar_init(bar0 + 8 * s, 1);
    fence_mbar_init();
  }
  __syncthreads();
  auto issue = [&](int k) {   // thread 0: block k of this CTA -> stage k % 3
    const int st = k % LNT_STAGES, r0 = (blockIdx.x + k * gridDim.x) * LNT_ROWS;
    const uint32_t bar = bar0 + 8 * st, dst = ring + st * stage_bytes;
    mbar_expect_tx(bar, stage_bytes);
    tma_load_3d(dst, &tmap_x, bar, 0, r0, 0);                  // one instruction per tensor: all column chunks of the 8 rows
    tma_load_3d(dst + half_stage, &tmap_dy, bar, 0, r0, 0);   // (TMA issue costs the issuing thread ~70 cycles per instruction)
  };
  if (threadIdx.x == 0) {
    if (nmine > 0) issue(0);
    if (nmine > 1) issue(1);
  }
  // this lane's vectors (8 columns each) of row `warp` inside a half stage, and this thread's phase-2 columns
  uint32_t voff[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int v = lane + 32 * i, j = v / vpc, q = v - j * vpc;
    voff[i] = ((j * LNT_ROWS + warp) * bc + q * 8) * 2;
  }
  constexpr int NJ = (NV * 256 + 1023) / 1024;
  uint32_t coff[NJ];
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const int col = threadIdx.x * 4 + 1024 * j, cj = col / bc, cq = col - cj * bc;
    coff[j] = (cj * LNT_ROWS * bc + cq) * 2;
  }
  float2 cg[NJ][2], cb[NJ][2], cbias[NJ][2];
#pragma unroll
  for (int j = 0; j < NJ; ++j)
#pragma unroll
    for (int k = 0; k < 2; ++k) cg[j][k] = cb[j][k] = cbias[j][k] = make_float2(0.f, 0.f);
  const float inv_d = 1.f / p.D;
  const bool dy_drop = p.dy_drop_p > 0.f, dx2_drop = p.dx2_drop_p > 0.f;
  const uint32_t bc2 = bc * 2;

  for (int k = 0; k < nmine; ++k) {
    const int st = k % LNT_STAGES, r0 = (blockIdx.x + k * gridDim.x) * LNT_ROWS;
    // stage (k+2) % 3 was read by block k-1: every thread passed that block's trailing barrier
    if (threadIdx.x == 0 && k + 2 < nmine) {
      fence_proxy_async_smem();   // generic-proxy accesses of that stage are ordered before the TMA writes
      issue(k + 2);
    }
    mbar_wait(bar0 + 8 * st, (k / LNT_STAGES) & 1);
    const uint32_t sx = ring + st * stage_bytes, sdy = sx + half_stage;
    const int r = r0 + warp;
    // ---------------- phase 1: warp = row; the unpacked row stays in registers for the second pass
    if (r < p.rows) {
      const float mean = p.mean[r], rstd = p.rstd[r];
      const uint32_t rh_dy = dy_drop ? drop_rowhash(p.dy_seed, static_cast<uint64_t>(r)) : 0u;
      float2 xv[NV][4], dv[NV][4];
      float2 s1v = make_float2(0.f, 0.f), tv = make_float2(0.f, 0.f);
      const float2 dys = make_float2(p.dy_scale, p.dy_scale);
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int vidx = lane + 32 * i;
        if (vidx < nvec) {
          unpack8x2(lds128(sx + voff[i]), xv[i]);
          unpack8x2(lds128(sdy + voff[i]), dv[i]);
          if (dy_drop) {   // the incoming gradient passes the forward's dropout mask first; phase 2 reads the masked copy
            dropmask8x2(dv[i], rh_dy, vidx * 8, p.dy_thresh, p.coltab);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) dv[i][kk] = __fmul2_rn(dv[i][kk], dys);
            sts128(sdy + voff[i], pack8x2(dv[i]));
          }
          const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.gamma + vidx * 8));
          const float4 g1 = __ldg(reinterpret_cast<const float4*>(p.gamma + vidx * 8) + 1);
          const float2 gm[4] = {make_float2(g0.x, g0.y), make_float2(g0.z, g0.w), make_float2(g1.x, g1.y), make_float2(g1.z, g1.w)};
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            dv[i][kk] = __fmul2_rn(dv[i][kk], gm[kk]);   // g = dy * gamma (dy itself is not needed again)
            s1v = __fadd2_rn(s1v, dv[i][kk]);
            tv = __ffma2_rn(dv[i][kk], xv[i][kk], tv);
          }
        }
      }
      const float s1 = warp_sum(s1v.x + s1v.y) * inv_d;
      const float t = warp_sum(tv.x + tv.y) * inv_d;
      const float s2 = rstd * (t - mean * s1);
      if (lane == 0) {
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(s_rstd + 4 * warp), "f"(rstd) : "memory");
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(s_nmr + 4 * warp), "f"(-mean * rstd) : "memory");
      }
      // dx = rstd * g + x * b + c
      const float bcoef = -rstd * rstd * s2, ccoef = -rstd * s1 - bcoef * mean;
      const float2 b2 = make_float2(bcoef, bcoef), c2 = make_float2(ccoef, ccoef), r2 = make_float2(rstd, rstd);
      __nv_bfloat16* dxr = p.dx + static_cast<long long>(r) * p.lddx;
      const uint32_t rh_dx2 = dx2_drop ? drop_rowhash(p.dx2_seed, static_cast<uint64_t>(r)) : 0u;
      const float2 dx2s = make_float2(p.dx2_scale, p.dx2_scale);
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int vidx = lane + 32 * i;
        if (vidx < nvec) {
          float2 o[4];
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) o[kk] = __ffma2_rn(dv[i][kk], r2, __ffma2_rn(xv[i][kk], b2, c2));
          uint4 q = pack8x2(o);
          *(reinterpret_cast<uint4*>(dxr) + vidx) = q;
          if (p.dx2) {
            if (dx2_drop) {
              dropmask8x2(o, rh_dx2, vidx * 8, p.dx2_thresh, p.coltab);
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) o[kk] = __fmul2_rn(o[kk], dx2s);
              q = pack8x2(o);
            }
            *(reinterpret_cast<uint4*>(p.dx2 + static_cast<long long>(r) * p.lddx) + vidx) = q;
          }
          if (p.dbias) sts128(s_dx2 + warp * row_bytes + vidx * 16, q);
        }
      }
    }
    __syncthreads();
    // ---------------- phase 2: thread = 4 columns, all rows of the block from shared memory (full blocks take the
    // branch-free unrolled path)
    const int nr = min(LNT_ROWS, p.rows - r0);
    auto col_row = [&](int j, uint32_t ax, uint32_t ad, uint32_t ao, int u) {
      const uint2 qx = lds64(ax), qd = lds64(ad);
      const float2 xv[2] = {make_float2(bf16_lo(qx.x), bf16_hi(qx.x)), make_float2(bf16_lo(qx.y), bf16_hi(qx.y))};
      const float2 dv[2] = {make_float2(bf16_lo(qd.x), bf16_hi(qd.x)), make_float2(bf16_lo(qd.y), bf16_hi(qd.y))};
      const float rs = lds32f(s_rstd + 4 * u), nm = lds32f(s_nmr + 4 * u);
      const float2 r2 = make_float2(rs, rs), n2 = make_float2(nm, nm);
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        cg[j][kk] = __ffma2_rn(dv[kk], __ffma2_rn(xv[kk], r2, n2), cg[j][kk]);   // dy * xhat
        cb[j][kk] = __fadd2_rn(cb[j][kk], dv[kk]);
      }
      if (p.dbias) {
        const uint2 qo = lds64(ao);
        cbias[j][0] = __fadd2_rn(cbias[j][0], make_float2(bf16_lo(qo.x), bf16_hi(qo.x)));
        cbias[j][1] = __fadd2_rn(cbias[j][1], make_float2(bf16_lo(qo.y), bf16_hi(qo.y)));
      }
    };
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int col = threadIdx.x * 4 + 1024 * j;
      if (col < p.D) {
        const uint32_t ax = sx + coff[j], ad = sdy + coff[j], ao = s_dx2 + col * 2;
        if (nr == LNT_ROWS) {
#pragma unroll
          for (int u = 0; u < LNT_ROWS; ++u) col_row(j, ax + u * bc2, ad + u * bc2, ao + u * row_bytes, u);
        } else {
          for (int u = 0; u < nr; ++u) col_row(j, ax + u * bc2, ad + u * bc2, ao + u * row_bytes, u);
        }
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const int col = threadIdx.x * 4 + 1024 * j;
    if (col < p.D) {
      // one 128-bit reduction per array instead of four scalar atomics (hundreds of CTAs hit the same D columns)
      red_add4(p.dgamma + col, cg[j][0], cg[j][1]);
      red_add4(p.dbeta + col, cb[j][0], cb[j][1]);
      if (p.dbias) red_add4(p.dbias + col, cbias[j][0], cbias[j][1]);
    }
  }
}

template <int NV>
static int launch_ln_bwd_tma(const LnBwdParams& p, int bc, cudaStream_t st) {
  CUtensorMap tx, td;
  int rc;
  if ((rc = make_tmap_rowblock_bf16(&tx, p.x, p.rows, p.D, p.ldx, bc, LNT_ROWS))) return rc;
  if ((rc = make_tmap_rowblock_bf16(&td, p.dy, p.rows, p.D, p.lddy, bc, LNT_ROWS))) return rc;
  const int smem = 128 + 128 + (2 * LNT_STAGES + 1) * LNT_ROWS * p.D * 2;
  static DeviceOnce once;
  if (int rc = once.run([] { XF_CUDA(cudaFuncSetAttribute(layernorm_bwd_tma_kernel<NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024)); return 0; }))
    return rc;
  const int nblk = (p.rows + LNT_ROWS - 1) / LNT_ROWS;
  int ctas = 2 * sm_count();
  if (ctas > nblk) ctas = nblk;
  layernorm_bwd_tma_kernel<NV><<<ctas, 256, smem, st>>>(tx, td, p, bc);
  return 0;
}

// ------------------------------------------------------------------------------------------
// column sums: out[n] += sum_m x[m, n]  (bias gradients of in_proj and linear1)
// ------------------------------------------------------------------------------------------
// CTA = 64 column-vectors (512 columns) x 4 row lanes; each thread keeps 4 independent 128-bit loads in
// flight; the 4 row lanes are combined through shared memory, then one atomic per column per CTA.
__global__ void __launch_bounds__(256)
colsum_kernel(const __nv_bfloat16* __restrict__ x, long long ld, int rows, int cols, int rows_per_cta, float* __restrict__ out) {
  __shared__ float red[4][64 * 8 + 8];
  const int vl = threadIdx.x & 63, rl = threadIdx.x >> 6;
  const int vc = blockIdx.x * 64 + vl;  // 8-column vector index
  const bool active = vc * 8 < cols;
  const int r0 = blockIdx.y * rows_per_cta;
  const int r1 = min(rows, r0 + rows_per_cta);
  float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (active) {
    int r = r0 + rl;
    for (; r + 12 < r1; r += 16) {
      uint4 q[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) q[u] = __ldg(reinterpret_cast<const uint4*>(x + static_cast<long long>(r + 4 * u) * ld) + vc);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        s[0] += bf16_lo(q[u].x); s[1] += bf16_hi(q[u].x); s[2] += bf16_lo(q[u].y); s[3] += bf16_hi(q[u].y);
        s[4] += bf16_lo(q[u].z); s[5] += bf16_hi(q[u].z); s[6] += bf16_lo(q[u].w); s[7] += bf16_hi(q[u].w);
      }
    }
    for (; r < r1; r += 4) {
      const uint4 q = __ldg(reinterpret_cast<const uint4*>(x + static_cast<long long>(r) * ld) + vc);
      s[0] += bf16_lo(q.x); s[1] += bf16_hi(q.x); s[2] += bf16_lo(q.y); s[3] += bf16_hi(q.y);
      s[4] += bf16_lo(q.z); s[5] += bf16_hi(q.z); s[6] += bf16_lo(q.w); s[7] += bf16_hi(q.w);
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) red[rl][vl * 8 + k] = s[k];
  __syncthreads();
  for (int c = threadIdx.x; c < 512; c += 256) {
    const int col = blockIdx.x * 512 + c;
    if (col < cols) atomicAdd(out + col, red[0][c] + red[1][c] + red[2][c] + red[3][c]);
  }
}

// ------------------------------------------------------------------------------------------
// cast fp32 [rows, cols] -> bf16 with optional row / column block padding
// (src block of `rin` rows -> dst block of `rout` rows; same for columns), and the inverse
// (fp32 padded -> fp32 compact, accumulate) used to un-pad weight gradients.
// ------------------------------------------------------------------------------------------
__global__ void cast_pad_kernel(const float* __restrict__ src, long long lds, __nv_bfloat16* __restrict__ dst, long long ldd,
                                int rows, int cols, int rin, int rout, int cin, int cout) {
  const long long total = static_cast<long long>(rows) * cols;
  for (long long t = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; t < total;
       t += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = t % cols;
    const int r = t / cols;
    const long long dr = rin > 0 ? static_cast<long long>(r / rin) * rout + r % rin : r;
    const long long dc = cin > 0 ? static_cast<long long>(c / cin) * cout + c % cin : c;
    dst[dr * ldd + dc] = __float2bfloat16(src[static_cast<long long>(r) * lds + c]);
  }
}

// contiguous, unpadded fast path: 8 elements per thread (2 x 128-bit loads, 1 x 128-bit store)
__global__ void __launch_bounds__(256) cast_vec8_kernel(const float4* __restrict__ src, uint4* __restrict__ dst, long long nvec) {
  for (long long t = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; t < nvec;
       t += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 a = __ldg(src + 2 * t), b = __ldg(src + 2 * t + 1);
    dst[t] = make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(b.x, b.y), pack_bf16(b.z, b.w));
  }
}

// all jobs of one call in a single launch: work unit = 8 consecutive source elements of a row
struct CastJobs {
  int n;
  long long unit_start[XF_CAST_MAX_JOBS + 1];   // prefix sums of units per job
  XfCastJob job[XF_CAST_MAX_JOBS];
};
__global__ void __launch_bounds__(256) cast_multi_kernel(const __grid_constant__ CastJobs js) {
  const long long total = js.unit_start[js.n];
  int j = 0;
  for (long long t = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; t < total;
       t += static_cast<long long>(gridDim.x) * blockDim.x) {
    while (t >= js.unit_start[j + 1]) ++j;   // t only grows
    const XfCastJob& J = js.job[j];
    const long long u = t - js.unit_start[j];
    const int upr = J.cols >> 3;              // units per row (cols % 8 == 0)
    const int r = static_cast<int>(u / upr), c = static_cast<int>(u - static_cast<long long>(r) * upr) * 8;
    const float4 a = __ldg(reinterpret_cast<const float4*>(J.src + static_cast<long long>(r) * J.lds + c));
    const float4 b = __ldg(reinterpret_cast<const float4*>(J.src + static_cast<long long>(r) * J.lds + c) + 1);
    const long long dr = J.rin > 0 ? static_cast<long long>(r / J.rin) * J.rout + r % J.rin : r;
    __nv_bfloat16* drow = reinterpret_cast<__nv_bfloat16*>(J.dst_bf16) + dr * J.ldd;
    if (J.cin > 0) {   // column blocks are padded: element-wise destination columns
      const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
      for (int k = 0; k < 8; ++k) drow[static_cast<long long>((c + k) / J.cin) * J.cout + (c + k) % J.cin] = __float2bfloat16(v[k]);
    } else {
      *reinterpret_cast<uint4*>(drow + c) = make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(b.x, b.y), pack_bf16(b.z, b.w));
    }
  }
}

__global__ void unpad_add_kernel(const float* __restrict__ src, long long lds, float* __restrict__ dst, long long ldd,
                                 int rows, int cols, int rin, int rout, int cin, int cout) {
  // dst is compact [rows, cols]; src is padded
  const long long total = static_cast<long long>(rows) * cols;
  for (long long t = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; t < total;
       t += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = t % cols;
    const int r = t / cols;
    const long long sr = rin > 0 ? static_cast<long long>(r / rin) * rout + r % rin : r;
    const long long sc = cin > 0 ? static_cast<long long>(c / cin) * cout + c % cin : c;
    dst[static_cast<long long>(r) * ldd + c] += src[sr * lds + sc];
  }
}

// delta[b, h, s] = sum_e O[b*S+s, h*dp + e] * dO[b*S+s, h*dp + e]   (attention backward pre-pass)
__global__ void __launch_bounds__(256)
attn_delta_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ d_o, long long ld, int B, int S,
                  int heads, int dp, int stat_stride, float* __restrict__ delta) {
  const int warps_per_cta = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const long long total = static_cast<long long>(B) * S * heads;
  for (long long w = blockIdx.x * static_cast<long long>(warps_per_cta) + (threadIdx.x >> 5); w < total;
       w += static_cast<long long>(gridDim.x) * warps_per_cta) {
    const int h = w % heads;
    const long long r = w / heads;   // token row b*S + s
    const int sidx = r % S;
    const int b = r / S;
    const __nv_bfloat16* po = o + r * ld + h * dp;
    const __nv_bfloat16* pd = d_o + r * ld + h * dp;
    float s = 0.f;
    for (int e = lane * 2; e < dp; e += 64) {
      const uint32_t a = *reinterpret_cast<const uint32_t*>(po + e);
      const uint32_t bb = *reinterpret_cast<const uint32_t*>(pd + e);
      s += bf16_lo(a) * bf16_lo(bb) + bf16_hi(a) * bf16_hi(bb);
    }
    s = warp_sum(s);
    if (lane == 0) delta[(static_cast<long long>(b) * heads + h) * stat_stride + sidx] = s;
  }
}

// 128-bit variant: one warp per token row, one head at a time (dp / 8 <= 32 vectors per head), 32-bit index math.
// The scalar kernel above spends most of its instructions on 64-bit div / mod per (row, head).
__global__ void __launch_bounds__(256)
attn_delta_vec_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ d_o, long long ld, int B, int S,
                      int heads, int dp, int stat_stride, float* __restrict__ delta) {
  const int warps_per_cta = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rows = B * S, vph = dp >> 3;
  for (int r = blockIdx.x * warps_per_cta + (threadIdx.x >> 5); r < rows; r += gridDim.x * warps_per_cta) {
    const int b = r / S, sidx = r - b * S;
    const uint4* po = reinterpret_cast<const uint4*>(o + static_cast<long long>(r) * ld);
    const uint4* pd = reinterpret_cast<const uint4*>(d_o + static_cast<long long>(r) * ld);
    for (int h = 0; h < heads; ++h) {
      float2 acc = make_float2(0.f, 0.f);
      if (lane < vph) {
        const uint4 a = __ldg(po + h * vph + lane), c = __ldg(pd + h * vph + lane);
        acc = __ffma2_rn(make_float2(bf16_lo(a.x), bf16_hi(a.x)), make_float2(bf16_lo(c.x), bf16_hi(c.x)), acc);
        acc = __ffma2_rn(make_float2(bf16_lo(a.y), bf16_hi(a.y)), make_float2(bf16_lo(c.y), bf16_hi(c.y)), acc);
        acc = __ffma2_rn(make_float2(bf16_lo(a.z), bf16_hi(a.z)), make_float2(bf16_lo(c.z), bf16_hi(c.z)), acc);
        acc = __ffma2_rn(make_float2(bf16_lo(a.w), bf16_hi(a.w)), make_float2(bf16_lo(c.w), bf16_hi(c.w)), acc);
      }
      const float sum = warp_sum(acc.x + acc.y);
      if (lane == 0) delta[(static_cast<long long>(b) * heads + h) * stat_stride + sidx] = sum;
    }
  }
}

// ------------------------------------------------------------------------------------------
// rows gather: out[r, :] = in[remap(r), :] (optionally dropout-masked with the mask of the forward
// write at that source position) and colsum += sum_r out[r, :].  Used for the visual rows of dz0
// (patch-embed backward: cross_f_box_layers.py:72-74 backward).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
rows_gather_kernel(const __nv_bfloat16* __restrict__ in, long long ldi, __nv_bfloat16* __restrict__ out, long long ldo, int rows,
                   int D, int rin, int rout, int roff, int rows_per_cta, float* __restrict__ colsum, float drop_p,
                   uint32_t seed, uint32_t stream, uint32_t thresh, float drop_scale, const uint32_t* __restrict__ coltab) {
  const int vc = blockIdx.x * blockDim.x + threadIdx.x;
  if (vc * 8 >= D) return;
  const int r0 = blockIdx.y * rows_per_cta;
  const int r1 = min(rows, r0 + rows_per_cta);
  float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int rb = r0; rb < r1; rb += 4) {   // four independent 128-bit loads in flight per thread
    uint4 q[4];
    long long ir[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      ir[u] = remap_row(min(rb + u, r1 - 1), rin, rout, roff);
      q[u] = __ldg(reinterpret_cast<const uint4*>(in + ir[u] * ldi) + vc);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int r = rb + u;
      if (r < r1) {
        float v[8] = {bf16_lo(q[u].x), bf16_hi(q[u].x), bf16_lo(q[u].y), bf16_hi(q[u].y),
                      bf16_lo(q[u].z), bf16_hi(q[u].z), bf16_lo(q[u].w), bf16_hi(q[u].w)};
        if (drop_p > 0.f) drop8(v, drop_rowhash(seed, static_cast<uint64_t>(ir[u])), vc * 8, thresh, drop_scale, coltab);
#pragma unroll
        for (int k = 0; k < 8; ++k) s[k] += v[k];
        *(reinterpret_cast<uint4*>(out + static_cast<long long>(r) * ldo) + vc) =
            make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
      }
    }
  }
  if (colsum) {
#pragma unroll
    for (int k = 0; k < 8; ++k) atomicAdd(colsum + vc * 8 + k, s[k]);
  }
}

static inline int grid_for(long long work_items, int per_cta, int waves = 8) {
  long long ctas = (work_items + per_cta - 1) / per_cta;
  long long cap = static_cast<long long>(sm_count()) * waves;
  if (ctas > cap) ctas = cap;
  if (ctas < 1) ctas = 1;
  return static_cast<int>(ctas);
}

}  // namespace xf

using namespace xf;

extern "C" int xf_patchify(const void* feat, int feat_dtype, void* tok, int64_t tok_ld, int B, int C, int H, int W, int p,
                           xf_stream_t s) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(s);
  if (!feat || !tok) return fail(-1, "xf_patchify: null pointer");
  if (!(p == 1 || p == 2 || p == 4 || p == 8) || H % p || W % p) return fail(-2, "xf_patchify: unsupported patch %d for %dx%d", p, H, W);
  if (tok_ld % 2 || (C * p * p) % 2) return fail(-4, "xf_patchify: token matrix width must be even");
  __nv_bfloat16* t = reinterpret_cast<__nv_bfloat16*>(tok);
  if (feat_dtype & 4) {   // channels_last (NHWC) memory
    int rc = (feat_dtype & 3) == 1 ? launch_patchify_fold_nhwc<float, false>(const_cast<float*>(reinterpret_cast<const float*>(feat)), t, B, C, H, W, p, tok_ld, stream)
                                   : launch_patchify_fold_nhwc<__nv_bfloat16, false>(const_cast<__nv_bfloat16*>(reinterpret_cast<const __nv_bfloat16*>(feat)), t, B, C, H, W, p, tok_ld, stream);
    if (rc) return rc;
    g_launches.fetch_add(1);
    XF_CUDA(cudaGetLastError());
    return 0;
  }
  if (feat_dtype == 1) launch_patchify_fold<float, false, false>(const_cast<float*>(reinterpret_cast<const float*>(feat)), t, B, C, H, W, p, tok_ld, stream);
  else if (feat_dtype == 0) launch_patchify_fold<__nv_bfloat16, false, false>(const_cast<__nv_bfloat16*>(reinterpret_cast<const __nv_bfloat16*>(feat)), t, B, C, H, W, p, tok_ld, stream);
  else return fail(-3, "xf_patchify: bad dtype");
  g_launches.fetch_add(1);
  XF_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int xf_fold(const void* tok, int64_t tok_ld, void* feat, int feat_dtype, int accumulate, int B, int C, int H, int W,
                       int p, xf_stream_t s) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(s);
  if (!feat || !tok) return fail(-1, "xf_fold: null pointer");
  if (!(p == 1 || p == 2 || p == 4 || p == 8) || H % p || W % p) return fail(-2, "xf_fold: unsupported patch %d for %dx%d", p, H, W);
  if (tok_ld % 2 || (C * p * p) % 2) return fail(-4, "xf_fold: token matrix width must be even");
  __nv_bfloat16* t = const_cast<__nv_bfloat16*>(reinterpret_cast<const __nv_bfloat16*>(tok));
  if (feat_dtype & 4) {   // channels_last (NHWC) memory
    if (accumulate) return fail(-5, "xf_fold: accumulate is not supported for channels_last maps");
    int rc = (feat_dtype & 3) == 1 ? launch_patchify_fold_nhwc<float, true>(reinterpret_cast<float*>(feat), t, B, C, H, W, p, tok_ld, stream)
                                   : launch_patchify_fold_nhwc<__nv_bfloat16, true>(reinterpret_cast<__nv_bfloat16*>(feat), t, B, C, H, W, p, tok_ld, stream);
    if (rc) return rc;
    g_launches.fetch_add(1);
    XF_CUDA(cudaGetLastError());
    return 0;
  }
  if (feat_dtype == 1) {
    if (accumulate) launch_patchify_fold<float, true, true>(reinterpret_cast<float*>(feat), t, B, C, H, W, p, tok_ld, stream);
    else launch_patchify_fold<float, true, false>(reinterpret_cast<float*>(feat), t, B, C, H, W, p, tok_ld, stream);
  } else if (feat_dtype == 0) {
    if (accumulate) launch_patchify_fold<__nv_bfloat16, true, true>(reinterpret_cast<__nv_bfloat16*>(feat), t, B, C, H, W, p, tok_ld, stream);
    else launch_patchify_fold<__nv_bfloat16, true, false>(reinterpret_cast<__nv_bfloat16*>(feat), t, B, C, H, W, p, tok_ld, stream);
  } else return fail(-3, "xf_fold: bad dtype");
  g_launches.fetch_add(1);
  XF_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int xf_lang_rows_fwd(const float* lang, const float* kind, void* z, int B, int L, int D, int n, int S, xf_stream_t s) {
  if (!lang || !kind || !z) return fail(-1, "xf_lang_rows_fwd: null pointer");
  if (D % 2) return fail(-2, "xf_lang_rows_fwd: D must be even");
  if (B * L == 0) return 0;
  lang_rows_fwd_kernel<<<grid_for(static_cast<long long>(B) * L * D / 2, 256), 256, 0, reinterpret_cast<cudaStream_t>(s)>>>(
      lang, kind, reinterpret_cast<__nv_bfloat16*>(z), B, L, D, n, S);
  g_launches.fetch_add(1);
  XF_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int xf_lang_rows_bwd(const void* dz, float* dlang, float* dkind, int B, int L, int D, int n, int S, xf_stream_t s) {
  if (!dz || !dkind) return fail(-1, "xf_lang_rows_bwd: null pointer");
  if (D % 2) return fail(-2, "xf_lang_rows_bwd: D must be even");
  if (B * L == 0) return 0;
  const int rows = B * L;
  const int gx = (D / 2 + 127) / 128;
  int gy = (2 * sm_count() + gx - 1) / gx;
  if (gy > rows) gy = rows;
  const int rows_per_cta = (rows + gy - 1) / gy;
  gy = (rows + rows_per_cta - 1) / rows_per_cta;
  lang_rows_bwd_kernel<<<dim3(gx, gy), 128, 0, reinterpret_cast<cudaStream_t>(s)>>>(reinterpret_cast<const __nv_bfloat16*>(dz), dlang, dkind,
                                                                                B, L, D, n, S, rows_per_cta);
  g_launches.fetch_add(1);
  XF_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int xf_layernorm_fwd(const XfLayerNorm* a, xf_stream_t s) {
  if (!a || !a->x || !a->y || !a->gamma || !a->beta) return fail(-1, "xf_layernorm_fwd: null pointer");
  if (a->D % 8 || a->D > LN_MAXV * 256) return fail(-2, "xf_layernorm_fwd: D=%d must be a multiple of 8 and <= %d", a->D, LN_MAXV * 256);
  if ((a->ldx % 8) || (a->ldy % 8)) return fail(-3, "xf_layernorm_fwd: leading dims must be multiples of 8");
  if (a->rows == 0) return 0;
  LnParams p;
  memset(&p, 0, sizeof(p));
  p.x = reinterpret_cast<const __nv_bfloat16*>(a->x); p.ldx = a->ldx;
  p.y = reinterpret_cast<__nv_bfloat16*>(a->y); p.ldy = a->ldy;
  p.gamma = a->gamma; p.beta = a->beta; p.mean = a->mean; p.rstd = a->rstd;
  p.rows = a->rows; p.D = a->D;
  p.in_rows_in = a->in_rows_in; p.in_rows_out = a->in_rows_out; p.in_row_off = a->in_row_off;
  p.out_rows_in = a->out_rows_in; p.out_rows_out = a->out_rows_out; p.out_row_off = a->out_row_off;
  p.eps = a->eps;
  p.drop_p = a->drop_p; p.drop_seed = drop_key(a->drop_seed, a->drop_stream); p.drop_stream = a->drop_stream;
  p.drop_thresh = drop_thresh32(a->drop_p);
  p.drop_scale = a->drop_p > 0.f ? 1.f / (1.f - a->drop_p) : 1.f;
  if (a->drop_p > 0.f) {
    if (a->D > static_cast<int>(XF_DROP_TABLE_COLS)) return fail(-6, "xf_layernorm_fwd: dropout supports D <= %u", XF_DROP_TABLE_COLS);
    if (!(p.coltab = drop_col_table())) return fail(-7, "xf_layernorm_fwd: dropout column table allocation failed");
  }
  const int nv = (a->D / 8 + 31) / 32;
  const int grid = grid_for(a->rows, 8);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(s);
  if (nv <= 1) layernorm_fwd_kernel<1><<<grid, 256, 0, st>>>(p);
  else if (nv <= 2) layernorm_fwd_kernel<2><<<grid, 256, 0, st>>>(p);
  else if (nv <= 4) layernorm_fwd_kernel<4><<<grid, 256, 0, st>>>(p);
  else layernorm_fwd_kernel<8><<<grid, 256, 0, st>>>(p);
  g_launches.fetch_add(1);
  XF_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int xf_layernorm_bwd(const XfLayerNormBwd* a, xf_stream_t s) {
  if (!a || !a->dy || !a->x || !a->gamma || !a->mean || !a->rstd || !a->dx || !a->dgamma || !a->dbeta)
    return fail(-1, "xf_layernorm_bwd: null pointer");
  if (a->D % 8 || a->D > LN_MAXV * 256) return fail(-2, "xf_layernorm_bwd: D=%d must be a multiple of 8 and <= %d", a->D, LN_MAXV * 256);
  if ((a->ldx % 8) || (a->lddy % 8) || (a->lddx % 8)) return fail(-3, "xf_layernorm_bwd: leading dims must be multiples of 8");
  if (a->rows == 0) return 0;
  LnBwdParams p;
  memset(&p, 0, sizeof(p));
  p.dy = reinterpret_cast<const __nv_bfloat16*>(a->dy); p.lddy = a->lddy;
  p.x = reinterpret_cast<const __nv_bfloat16*>(a->x); p.ldx = a->ldx;
  p.gamma = a->gamma; p.mean = a->mean; p.rstd = a->rstd;
  p.dx = reinterpret_cast<__nv_bfloat16*>(a->dx); p.lddx = a->lddx;
  p.dx2 = reinterpret_cast<__nv_bfloat16*>(a->dx2);
  p.dgamma = a->dgamma; p.dbeta = a->dbeta; p.dbias = a->dbias;
  p.rows = a->rows; p.D = a->D;
  p.in_rows_in = a->in_rows_in; p.in_rows_out = a->in_rows_out; p.in_row_off = a->in_row_off;
  p.out_rows_in = a->out_rows_in; p.out_rows_out = a->out_rows_out; p.out_row_off = a->out_row_off;
  p.dy_drop_p = a->dy_drop_p; p.dy_seed = drop_key(a->dy_drop_seed, a->dy_drop_stream); p.dy_stream = a->dy_drop_stream;
  p.dy_thresh = drop_thresh32(a->dy_drop_p);
  p.dy_scale = a->dy_drop_p > 0.f ? 1.f / (1.f - a->dy_drop_p) : 1.f;
  p.dx2_drop_p = a->dx2_drop_p; p.dx2_seed = drop_key(a->dx2_drop_seed, a->dx2_drop_stream); p.dx2_stream = a->dx2_drop_stream;
  p.dx2_thresh = drop_thresh32(a->dx2_drop_p);
  p.dx2_scale = a->dx2_drop_p > 0.f ? 1.f / (1.f - a->dx2_drop_p) : 1.f;
  if (a->dy_drop_p > 0.f || a->dx2_drop_p > 0.f) {
    if (a->D > static_cast<int>(XF_DROP_TABLE_COLS)) return fail(-6, "xf_layernorm_bwd: dropout supports D <= %u", XF_DROP_TABLE_COLS);
    if (!(p.coltab = drop_col_table())) return fail(-7, "xf_layernorm_bwd: dropout column table allocation failed");
  }
  int ctas = (a->rows + LNB_ROWS - 1) / LNB_ROWS;
  if (ctas > 4 * sm_count()) ctas = 4 * sm_count();
  if (ctas < 1) ctas = 1;
  const int nv = (a->D / 8 + 31) / 32;
  const size_t sh = 0;
  if (a->D % 4) return fail(-4, "xf_layernorm_bwd: D must be a multiple of 4");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(s);
  // TMA-pipelined kernel: contiguous row ranges (no remap), 16-byte aligned operands, stage ring within 113 KB
  {
    int bc = 0;
    for (int c = 256; c >= 8; c -= 8)
      if (a->D % c == 0) { bc = c; break; }
    const bool aligned = ((reinterpret_cast<uintptr_t>(a->x) | reinterpret_cast<uintptr_t>(a->dy) | reinterpret_cast<uintptr_t>(a->dx) |
                           reinterpret_cast<uintptr_t>(a->dx2)) & 15) == 0;
    const int smem = 256 + (2 * LNT_STAGES + 1) * LNT_ROWS * a->D * 2;
    if (a->in_rows_in == 0 && a->out_rows_in == 0 && bc >= 8 && aligned && smem <= 113 * 1024 && nv <= 4) {
      int rc;
      if (nv <= 1) rc = launch_ln_bwd_tma<1>(p, bc, st);
      else if (nv <= 2) rc = launch_ln_bwd_tma<2>(p, bc, st);
      else rc = launch_ln_bwd_tma<4>(p, bc, st);
      if (rc) return rc;
      g_launches.fetch_add(1);
      XF_CUDA(cudaGetLastError());
      return 0;
    }
  }
  if (nv <= 1) layernorm_bwd_kernel<1><<<ctas, 256, sh, st>>>(p);
  else if (nv <= 2) layernorm_bwd_kernel<2><<<ctas, 256, sh, st>>>(p);
  else if (nv <= 4) layernorm_bwd_kernel<4><<<ctas, 256, sh, st>>>(p);
  else layernorm_bwd_kernel<8><<<ctas, 256, sh, st>>>(p);
  g_launches.fetch_add(1);
  XF_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int xf_colsum(const void* x, int64_t ld, int rows, int cols, float* out, xf_stream_t s) {
  if (!x || !out) return fail(-1, "xf_colsum: null pointer");
  if (cols % 8 || ld % 8) return fail(-2, "xf_colsum: cols and ld must be multiples of 8");
  if (rows == 0) return 0;
  const int vcols = cols / 8;
  const int gx = (vcols + 63) / 64;
  int gy = (4 * sm_count() + gx - 1) / gx;
  if (gy * 16 > rows) gy = (rows + 15) / 16;
  if (gy < 1) gy = 1;
  const int rows_per_cta = (rows + gy - 1) / gy;
  gy = (rows + rows_per_cta - 1) / rows_per_cta;
  colsum_kernel<<<dim3(gx, gy), 256, 0, reinterpret_cast<cudaStream_t>(s)>>>(reinterpret_cast<const __nv_bfloat16*>(x), ld, rows, cols,
                                                                            rows_per_cta, out);
  g_launches.fetch_add(1);
  XF_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int xf_cast_pad(const float* src, int64_t lds, void* dst, int64_t ldd, int rows, int cols, int rin, int rout, int cin,
                           int cout, xf_stream_t s) {
  if (!src || !dst) return fail(-1, "xf_cast_pad: null pointer");
  if (rows == 0 || cols == 0) return 0;
  const long long total = static_cast<long long>(rows) * cols;
  if (rin == 0 && cin == 0 && lds == cols && ldd == cols && total % 8 == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
    cast_vec8_kernel<<<grid_for(total / 8, 256), 256, 0, reinterpret_cast<cudaStream_t>(s)>>>(
        reinterpret_cast<const float4*>(src), reinterpret_cast<uint4*>(dst), total / 8);
    g_launches.fetch_add(1);
    XF_CUDA(cudaGetLastError());
    return 0;
  }
  cast_pad_kernel<<<grid_for(static_cast<long long>(rows) * cols, 256 * 4), 256, 0, reinterpret_cast<cudaStream_t>(s)>>>(
      src, lds, reinterpret_cast<__nv_bfloat16*>(dst), ldd, rows, cols, rin, rout, cin, cout);
  g_launches.fetch_add(1);
  XF_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int xf_unpad_add(const float* src, int64_t lds, float* dst, int64_t ldd, int rows, int cols, int rin, int rout, int cin,
                            int cout, xf_stream_t s) {
  if (!src || !dst) return fail(-1, "xf_unpad_add: null pointer");
  if (rows == 0 || cols == 0) return 0;
  unpad_add_kernel<<<grid_for(static_cast<long long>(rows) * cols, 256 * 4), 256, 0, reinterpret_cast<cudaStream_t>(s)>>>(
      src, lds, dst, ldd, rows, cols, rin, rout, cin, cout);
  g_launches.fetch_add(1);
  XF_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int xf_cast_pad_multi(const XfCastJob* jobs, int n_jobs, xf_stream_t s) {
  if (!jobs || n_jobs < 0) return fail(-1, "xf_cast_pad_multi: null pointer");
  if (n_jobs > XF_CAST_MAX_JOBS) return fail(-2, "xf_cast_pad_multi: at most %d jobs per call", XF_CAST_MAX_JOBS);
  CastJobs js;
  memset(&js, 0, sizeof(js));
  long long units = 0;
  for (int i = 0; i < n_jobs; ++i) {
    const XfCastJob& J = jobs[i];
    if (!J.src || !J.dst_bf16) return fail(-1, "xf_cast_pad_multi: null pointer in job %d", i);
    if (J.rows <= 0 || J.cols <= 0) continue;
    const bool ok = J.cols % 8 == 0 && J.lds % 4 == 0 && (reinterpret_cast<uintptr_t>(J.src) & 15) == 0 &&
                    (J.cin > 0 || (J.ldd % 8 == 0 && (reinterpret_cast<uintptr_t>(J.dst_bf16) & 15) == 0));
    if (!ok) {   // odd shapes / alignments: the single-tensor path handles them
      int rc = xf_cast_pad(J.src, J.lds, J.dst_bf16, J.ldd, J.rows, J.cols, J.rin, J.rout, J.cin, J.cout, s);
      if (rc) return rc;
      continue;
    }
    js.job[js.n] = J;
    js.unit_start[js.n] = units;
    units += static_cast<long long>(J.rows) * (J.cols / 8);
    ++js.n;
  }
  js.unit_start[js.n] = units;
  if (units == 0) return 0;
  cast_multi_kernel<<<grid_for(units, 256), 256, 0, reinterpret_cast<cudaStream_t>(s)>>>(js);
  g_launches.fetch_add(1);
  XF_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int xf_attn_delta(const void* o, const void* d_o, int64_t ld, int B, int S, int heads, int dp, int stat_stride,
                             float* delta, xf_stream_t s) {
  if (!o || !d_o || !delta) return fail(-1, "xf_attn_delta: null pointer");
  if (dp % 2) return fail(-2, "xf_attn_delta: dp must be even");
  if (stat_stride < S) return fail(-3, "xf_attn_delta: stat_stride < S");
  if (B * S == 0) return 0;
  const bool vec = dp % 8 == 0 && dp <= 256 && ld % 8 == 0 && ((reinterpret_cast<uintptr_t>(o) | reinterpret_cast<uintptr_t>(d_o)) & 15) == 0 &&
                   static_cast<long long>(B) * S < (1ll << 31);
  if (vec)
    attn_delta_vec_kernel<<<grid_for(static_cast<long long>(B) * S, 8), 256, 0, reinterpret_cast<cudaStream_t>(s)>>>(
        reinterpret_cast<const __nv_bfloat16*>(o), reinterpret_cast<const __nv_bfloat16*>(d_o), ld, B, S, heads, dp, stat_stride, delta);
  else
    attn_delta_kernel<<<grid_for(static_cast<long long>(B) * S * heads, 8), 256, 0, reinterpret_cast<cudaStream_t>(s)>>>(
        reinterpret_cast<const __nv_bfloat16*>(o), reinterpret_cast<const __nv_bfloat16*>(d_o), ld, B, S, heads, dp, stat_stride, delta);
  g_launches.fetch_add(1);
  XF_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int xf_rows_gather(const void* in, int64_t ldi, void* out, int64_t ldo, int rows, int D, int rin, int rout, int roff,
                              float* colsum, float drop_p, uint32_t seed, uint32_t stream_id, xf_stream_t s) {
  if (!in || !out) return fail(-1, "xf_rows_gather: null pointer");
  if (D % 8 || ldi % 8 || ldo % 8) return fail(-2, "xf_rows_gather: D and leading dims must be multiples of 8");
  if (rows == 0) return 0;
  const int gx = (D / 8 + 255) / 256;
  int gy = (4 * sm_count() + gx - 1) / gx;
  if (gy > rows) gy = rows;
  const int rows_per_cta = (rows + gy - 1) / gy;
  gy = (rows + rows_per_cta - 1) / rows_per_cta;
  const uint32_t* coltab = nullptr;
  if (drop_p > 0.f) {
    if (D > static_cast<int>(XF_DROP_TABLE_COLS)) return fail(-6, "xf_rows_gather: dropout supports D <= %u", XF_DROP_TABLE_COLS);
    if (!(coltab = drop_col_table())) return fail(-7, "xf_rows_gather: dropout column table allocation failed");
  }
  rows_gather_kernel<<<dim3(gx, gy), 256, 0, reinterpret_cast<cudaStream_t>(s)>>>(
      reinterpret_cast<const __nv_bfloat16*>(in), ldi, reinterpret_cast<__nv_bfloat16*>(out), ldo, rows, D, rin, rout, roff,
      rows_per_cta, colsum, drop_p, drop_key(seed, stream_id), stream_id, drop_thresh32(drop_p),
      drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f, coltab);
  g_launches.fetch_add(1);
  XF_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------
// bf16 -> fp32 with a scale (the inverse of the weight cast): unpacks a bf16-compressed gradient arena after its
// all-reduce and applies the 1 / world averaging in the same pass.
// ------------------------------------------------------------------------------------------
namespace xf {
__global__ void __launch_bounds__(256) bf16_to_f32_scale_kernel(const uint4* __restrict__ src, float4* __restrict__ dst, long long nvec, float scale) {
  for (long long t = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; t < nvec;
       t += static_cast<long long>(gridDim.x) * blockDim.x) {
    const uint4 q = __ldg(src + t);
    dst[2 * t] = make_float4(bf16_lo(q.x) * scale, bf16_hi(q.x) * scale, bf16_lo(q.y) * scale, bf16_hi(q.y) * scale);
    dst[2 * t + 1] = make_float4(bf16_lo(q.z) * scale, bf16_hi(q.z) * scale, bf16_lo(q.w) * scale, bf16_hi(q.w) * scale);
  }
}
}  // namespace xf

extern "C" int xf_bf16_to_f32(const void* src_bf16, float* dst, int64_t n, float scale, xf_stream_t s) {
  if (!src_bf16 || !dst) return fail(-1, "xf_bf16_to_f32: null pointer");
  if (n % 8 || ((reinterpret_cast<uintptr_t>(src_bf16) | reinterpret_cast<uintptr_t>(dst)) & 15))
    return fail(-2, "xf_bf16_to_f32: n must be a multiple of 8 and the pointers 16-byte aligned");
  if (n == 0) return 0;
  bf16_to_f32_scale_kernel<<<grid_for(n / 8, 256), 256, 0, reinterpret_cast<cudaStream_t>(s)>>>(
      reinterpret_cast<const uint4*>(src_bf16), reinterpret_cast<float4*>(dst), n / 8, scale);
  g_launches.fetch_add(1);
  XF_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------
// Test aid: materialise the counter-based dropout keep masks (ptx.cuh) so a CPU checker can apply the SAME masks.
// These are the canonical definitions every kernel of the library recomputes on the fly.
// ------------------------------------------------------------------------------------------
namespace xf {
__global__ void debug_drop_mask_kernel(uint8_t* __restrict__ out, long long row0, int rows, int cols, uint32_t key, uint32_t t32) {
  const long long total = static_cast<long long>(rows) * cols;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / cols;
    const uint32_t c = static_cast<uint32_t>(i - r * cols);
    out[i] = drop_keep_rc(drop_rowhash(key, static_cast<uint64_t>(row0 + r)), drop_colodd(c), t32) ? 1 : 0;
  }
}
__global__ void debug_attn_drop_mask_kernel(uint8_t* __restrict__ out, int BH, int Sq, int Sk, uint32_t key, uint32_t t32) {
  const long long total = static_cast<long long>(BH) * Sq * Sk;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long rq = i / Sk;   // bh * Sq + q: the row index the attention kernels hash
    const uint32_t k = static_cast<uint32_t>(i - rq * Sk);
    out[i] = drop_keep_rc(drop_rowhash(key, static_cast<uint64_t>(rq)), drop_colhash(key, k), t32) ? 1 : 0;
  }
}
}  // namespace xf

extern "C" int xf_debug_dropout_mask(float p, uint32_t seed, uint32_t stream_id, int64_t row0, int rows, int cols, uint8_t* out,
                                     xf_stream_t s) {
  if (!out) return fail(-1, "xf_debug_dropout_mask: null pointer");
  if (rows <= 0 || cols <= 0) return 0;
  debug_drop_mask_kernel<<<grid_for(static_cast<long long>(rows) * cols, 256), 256, 0, reinterpret_cast<cudaStream_t>(s)>>>(
      out, row0, rows, cols, drop_key(seed, stream_id), drop_thresh32(p));
  g_launches.fetch_add(1);
  XF_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int xf_debug_attn_dropout_mask(float p, uint32_t seed, uint32_t stream_id, int BH, int Sq, int Sk, uint8_t* out,
                                          xf_stream_t s) {
  if (!out) return fail(-1, "xf_debug_attn_dropout_mask: null pointer");
  if (BH <= 0 || Sq <= 0 || Sk <= 0) return 0;
  debug_attn_drop_mask_kernel<<<grid_for(static_cast<long long>(BH) * Sq * Sk, 256), 256, 0, reinterpret_cast<cudaStream_t>(s)>>>(
      out, BH, Sq, Sk, drop_key(seed, stream_id), drop_thresh32(p));
  g_launches.fetch_add(1);
  XF_CUDA(cudaGetLastError());
  return 0;
}
