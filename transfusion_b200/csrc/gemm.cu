// xf_gemm: persistent, warp-specialised tcgen05 GEMM for sm_100a.
//
//   D[M,N] (+)= A(M x K) * B(N x K)^T,  bf16 operands, fp32 accumulation in TMEM.
//
// CTA = 320 threads: warps 0..7 = epilogue (two warps per TMEM lane quadrant, alternating 32-column
// chunks), warp 8 = TMA producer, warp 9 = MMA issuer (+ TMEM alloc): the scheduler favours high warp ids.  Tile = 128 x tile_n x 64; operands are staged by TMA
// into a ring of SWIZZLE_128B shared-memory stages; tcgen05.mma (M=128, N=tile_n, K=16) reads them
// through shared-memory descriptors in either K-major or MN-major form, so forward (x W^T), dgrad
// (dy W) and wgrad (dy^T x) all run on the same kernel without materialised transposes.  The
// accumulator is double-buffered in TMEM (2 x tile_n columns) so the epilogue of tile i overlaps the
// MMAs of tile i+1.  Work items (m-tile, n-tile, k-split) are distributed round-robin over a
// persistent grid of one CTA per SM.
//
// Shared-memory stage layout (all atoms 1024-byte aligned):
//   A: K-major  -> one TMA box {64 k, 128 rows}            = 128 rows x 128 B
//      MN-major -> two TMA boxes {64 m, 64 k}, 8 KB apart  = chunk c holds m in [64c, 64c+64)
//   B: K-major  -> one TMA box {64 k, tile_n rows}
//      MN-major -> ceil(tile_n/64) TMA boxes {64 n, 64 k}, 8 KB apart
#include "../../include/xfusion.h"
#include "host_common.cuh"
#include "ptx.cuh"
#include <string.h>

namespace xf {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_EPI_WARPS = 8;   // two warps per TMEM lane quadrant, each takes every other 32-column chunk
constexpr int GEMM_EPI_GROUPS = GEMM_EPI_WARPS / 4;
constexpr int GEMM_THREADS = 64 + 32 * GEMM_EPI_WARPS;
constexpr int GEMM_A_BYTES = GEMM_BM * GEMM_BK * 2;  // 16 KB
constexpr int GEMM_CTRL_BYTES = 5120;                // barriers, TMEM pointer, per-tile bias / column-hash staging
constexpr int GEMM_IO_BYTES = GEMM_EPI_WARPS * 4096;  // per epilogue warp: two [32 rows x 64 B] SWIZZLE_64B boxes (TMA store / load)

struct GemmParams {
  int M, N, K;
  int tile_n;
  int a_mn, b_mn;
  int num_m_tiles, num_n_tiles, split_k, kb_per_split, total_kb;
  int stages, b_bytes, stage_bytes;
  // batching: items enumerate (batch, m-tile, k-split, n-tile); batch = b1 * nb2 + b2 addresses the operands through the
  // two outer dimensions of 4-D tensor maps (A, B, out, aux) -- M, N, K are per batch entry
  int batched, nb2, items_per_batch;
  long long out_bs1, out_bs2;   // element offsets of the output per batch index (fp32 reduction epilogue)
  // epilogue
  const float* bias;
  const float* pos_table;
  int rows_in, rows_out, row_off;
  int act;
  __nv_bfloat16* preact_out;
  const __nv_bfloat16* dact_in;
  const __nv_bfloat16* residual;
  long long ldr;
  void* out;
  long long ldc;
  int out_f32, accumulate, vec_ok;
  float drop_p;
  uint32_t drop_seed, drop_stream, drop_thresh;
  float drop_scale;
  int drop_first;
};

__device__ __forceinline__ void decode_item(const GemmParams& p, int item, int& mt, int& nt, int& kb0, int& nkb, int& b1, int& b2) {
  b1 = b2 = 0;
  if (p.batched) {
    const int batch = item / p.items_per_batch;
    item -= batch * p.items_per_batch;
    b1 = batch / p.nb2;
    b2 = batch - b1 * p.nb2;
  }
  nt = item % p.num_n_tiles;
  int r = item / p.num_n_tiles;
  int ks = r % p.split_k;
  mt = r / p.split_k;
  kb0 = ks * p.kb_per_split;
  int kb1 = min(p.total_kb, kb0 + p.kb_per_split);
  nkb = max(0, kb1 - kb0);
}

// One 32-column chunk of one accumulator row.
__device__ __forceinline__ void epilogue_chunk(const GemmParams& p, int m, int n0, uint32_t (&acc)[32]) {
  if (m >= p.M) return;
  int m_out = m, m_in = m;
  if (p.rows_in > 0) {
    int g = m / p.rows_in;
    m_in = m - g * p.rows_in;
    m_out = g * p.rows_out + m_in + p.row_off;
  }
  const bool full = (n0 + 32 <= p.N);
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(acc[i]);

  if (p.bias) {
    if (full) {
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + i));
        v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (n0 + i < p.N) v[i] += __ldg(p.bias + n0 + i);
    }
  }
  if (p.pos_table) {
    const float* pt = p.pos_table + static_cast<long long>(m_in) * p.N + n0;
    if (full && p.vec_ok) {
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        float4 b = __ldg(reinterpret_cast<const float4*>(pt + i));
        v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (n0 + i < p.N) v[i] += __ldg(pt + i);
    }
  }
  const long long orow = static_cast<long long>(m_out) * p.ldc + n0;
  if (p.drop_p > 0.f && p.drop_first) {
    const uint32_t rh = drop_rowhash(p.drop_seed, static_cast<uint64_t>(m_out));
#pragma unroll
    for (int i = 0; i < 32; ++i)
      v[i] = drop_keep_rc(rh, drop_colodd(static_cast<uint32_t>(n0 + i)), p.drop_thresh) ? v[i] * p.drop_scale : 0.f;
  }
  if (p.act == 1) {
    if (p.preact_out) {
      __nv_bfloat16* po = p.preact_out + orow;
      if (full && p.vec_ok) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 q = make_uint4(pack_bf16(v[i], v[i + 1]), pack_bf16(v[i + 2], v[i + 3]), pack_bf16(v[i + 4], v[i + 5]),
                               pack_bf16(v[i + 6], v[i + 7]));
          *reinterpret_cast<uint4*>(po + i) = q;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (n0 + i < p.N) po[i] = __float2bfloat16(v[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = gelu_erf(v[i]);
  } else if (p.act == 2 && !p.dact_in) {   // ReLU (TwoMLPHead fc6 / fc7)
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
  }
  if (p.dact_in && p.act == 2) {   // ReLU backward: dact_in holds the forward OUTPUT (> 0 exactly where the input was)
    const __nv_bfloat16* di = p.dact_in + orow;
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (n0 + i < p.N && !(__bfloat162float(di[i]) > 0.f)) v[i] = 0.f;
  } else if (p.dact_in) {
    const __nv_bfloat16* di = p.dact_in + orow;
    if (full && p.vec_ok) {
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        uint4 q = __ldg(reinterpret_cast<const uint4*>(di + i));
        v[i] *= gelu_erf_grad(bf16_lo(q.x)); v[i + 1] *= gelu_erf_grad(bf16_hi(q.x));
        v[i + 2] *= gelu_erf_grad(bf16_lo(q.y)); v[i + 3] *= gelu_erf_grad(bf16_hi(q.y));
        v[i + 4] *= gelu_erf_grad(bf16_lo(q.z)); v[i + 5] *= gelu_erf_grad(bf16_hi(q.z));
        v[i + 6] *= gelu_erf_grad(bf16_lo(q.w)); v[i + 7] *= gelu_erf_grad(bf16_hi(q.w));
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (n0 + i < p.N) v[i] *= gelu_erf_grad(__bfloat162float(di[i]));
    }
  }
  if (p.drop_p > 0.f && !p.drop_first) {
    const uint32_t rh = drop_rowhash(p.drop_seed, static_cast<uint64_t>(m_out));
#pragma unroll
    for (int i = 0; i < 32; ++i)
      v[i] = drop_keep_rc(rh, drop_colodd(static_cast<uint32_t>(n0 + i)), p.drop_thresh) ? v[i] * p.drop_scale : 0.f;
  }
  if (p.residual) {
    const __nv_bfloat16* rs = p.residual + static_cast<long long>(m_out) * p.ldr + n0;
    if (full && p.vec_ok) {
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        uint4 q = __ldg(reinterpret_cast<const uint4*>(rs + i));
        v[i] += bf16_lo(q.x); v[i + 1] += bf16_hi(q.x); v[i + 2] += bf16_lo(q.y); v[i + 3] += bf16_hi(q.y);
        v[i + 4] += bf16_lo(q.z); v[i + 5] += bf16_hi(q.z); v[i + 6] += bf16_lo(q.w); v[i + 7] += bf16_hi(q.w);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (n0 + i < p.N) v[i] += __bfloat162float(rs[i]);
    }
  }
  if (p.out_f32) {
    float* o = reinterpret_cast<float*>(p.out) + orow;
    if (p.accumulate) {
      if (full && p.vec_ok) {
        // 128-bit vector reductions: 8 L2 transactions per 32-column chunk instead of 32
#pragma unroll
        for (int i = 0; i < 32; i += 4)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + i), "f"(v[i]), "f"(v[i + 1]), "f"(v[i + 2]),
                       "f"(v[i + 3])
                       : "memory");
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (n0 + i < p.N) atomicAdd(o + i, v[i]);
      }
    } else if (full && p.vec_ok) {
#pragma unroll
      for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(o + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (n0 + i < p.N) o[i] = v[i];
    }
  } else {
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + orow;
    if (full && p.vec_ok) {
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        uint4 q = make_uint4(pack_bf16(v[i], v[i + 1]), pack_bf16(v[i + 2], v[i + 3]), pack_bf16(v[i + 4], v[i + 5]),
                             pack_bf16(v[i + 6], v[i + 7]));
        *reinterpret_cast<uint4*>(o + i) = q;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (n0 + i < p.N) o[i] = __float2bfloat16(v[i]);
    }
  }
}

// ---- specialised epilogues (the common cases; everything else takes epilogue_chunk above) --------------------
// They exist because the epilogue, not the tensor pipe, paced the K = 896 GEMMs: ~40 instructions per element in
// the generic path.  Here: packed fp32 math (fma/add/mul .f32x2), bias and dropout column hashes staged once per
// tile in shared memory and re-read as 128-bit broadcasts, residual / pre-activation operands requested before
// the accumulator load is waited for, no per-element bounds checks (host guarantees N % 8 == 0 and 16-byte
// alignment), and one kernel instantiation per kind so each fits the instruction cache.
enum { EPI_GENERIC = 0, EPI_LINEAR = 1, EPI_GELU = 2, EPI_DGELU = 3, EPI_RED = 4 };

__device__ __forceinline__ void unpack32(const uint4 (&q)[4], float2 (&f)[16]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[4 * i] = make_float2(bf16_lo(q[i].x), bf16_hi(q[i].x));
    f[4 * i + 1] = make_float2(bf16_lo(q[i].y), bf16_hi(q[i].y));
    f[4 * i + 2] = make_float2(bf16_lo(q[i].z), bf16_hi(q[i].z));
    f[4 * i + 3] = make_float2(bf16_lo(q[i].w), bf16_hi(q[i].w));
  }
}
// This thread's row (lane) of a [32 rows x 64 B] SWIZZLE_64B box: 16-byte chunk ch sits at ch ^ ((row >> 1) & 3).
__device__ __forceinline__ void box_store_row(uint32_t box, int lane, const float2 (&v)[16]) {
  const uint32_t row = box + lane * 64, sw = (lane >> 1) & 3;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row + ((i ^ sw) << 4)), "r"(pack_bf16(v[4 * i].x, v[4 * i].y)),
                 "r"(pack_bf16(v[4 * i + 1].x, v[4 * i + 1].y)), "r"(pack_bf16(v[4 * i + 2].x, v[4 * i + 2].y)),
                 "r"(pack_bf16(v[4 * i + 3].x, v[4 * i + 3].y))
                 : "memory");
}
__device__ __forceinline__ void box_load_row(uint32_t box, int lane, uint4 (&q)[4]) {
  const uint32_t row = box + lane * 64, sw = (lane >> 1) & 3;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(q[i].x), "=r"(q[i].y), "=r"(q[i].z), "=r"(q[i].w) : "r"(row + ((i ^ sw) << 4)));
}
// Warp-collective: the 32 x 32 bf16 block in `box` -> global through the tensor map (one bulk group of lane 0).
// PENDING = bulk groups of this warp that may still be reading OTHER boxes when `box` is overwritten.
template <int PENDING>
__device__ __forceinline__ void box_store_begin(int lane) {
  if (lane == 0) tma_store_wait_read<PENDING>();
  __syncwarp();
}
// bt < 0: 2-D tensor map; else 4-D (n, m, b2 = bt & 0xFFFF, b1 = bt >> 16)
__device__ __forceinline__ void box_store_issue(const CUtensorMap* tmap, uint32_t box, int lane, int n0, int m0, int bt) {
  fence_proxy_async_smem();
  __syncwarp();
  if (lane == 0) {
    if (bt < 0) tma_store_2d(tmap, box, n0, m0);
    else tma_store_4d(tmap, box, n0, m0, bt & 0xFFFF, bt >> 16);
    tma_store_commit();
  }
}
__device__ __forceinline__ void box_load_issue(const CUtensorMap* tmap, uint32_t box, uint32_t bar, int n0, int m0, int bt) {
  if (bt < 0) tma_load_2d(box, tmap, bar, n0, m0);
  else tma_load_4d(box, tmap, bar, n0, m0, bt & 0xFFFF, bt >> 16);
}
// dropout mask of row hash rh on 32 columns whose odd column hashes sit at shared address col_addr (no scaling)
__device__ __forceinline__ void drop32(float2 (&v)[16], uint32_t rh, uint32_t col_addr, uint32_t t32) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    uint4 h;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(h.x), "=r"(h.y), "=r"(h.z), "=r"(h.w) : "r"(col_addr + 16 * i));
    if (!drop_keep_rc(rh, h.x, t32)) v[2 * i].x = 0.f;
    if (!drop_keep_rc(rh, h.y, t32)) v[2 * i].y = 0.f;
    if (!drop_keep_rc(rh, h.z, t32)) v[2 * i + 1].x = 0.f;
    if (!drop_keep_rc(rh, h.w, t32)) v[2 * i + 1].y = 0.f;
  }
}

// One 32-column chunk of the warp's 32 accumulator rows (warp-collective; rows m0 .. m0+31, columns n0 .. n0+31).
// All global traffic goes through TMA boxes: box0 receives the residual (EPI_LINEAR) or the saved pre-activation
// (EPI_DGELU) -- already copied to `aux` registers by the caller -- and stages the pre-activation store of
// EPI_GELU; box1 stages the output.  Row-per-thread global accesses (32 different 128-byte lines per warp
// instruction) saturated the load/store unit and delayed the shared-memory broadcasts queued behind them.
template <int EPI>
__device__ __forceinline__ void epilogue_fast(const GemmParams& p, const CUtensorMap* tmap_out, const CUtensorMap* tmap_aux, int lane,
                                              int m0, int n0, const uint32_t (&acc)[32], const uint4 (&aux)[4], uint32_t box0,
                                              uint32_t box1, uint32_t bias_addr, uint32_t col_addr, uint32_t rh, int bt) {
  if constexpr (EPI == EPI_RED) {
    float* o = reinterpret_cast<float*>(p.out) + static_cast<long long>(m0 + lane) * p.ldc + n0;
    if (bt >= 0) o += (bt >> 16) * p.out_bs1 + (bt & 0xFFFF) * p.out_bs2;
    if (m0 + lane < p.M) {
#pragma unroll
      for (int i = 0; i < 32; i += 4)   // 128-bit vector reductions: 8 L2 transactions per chunk instead of 32
        if (n0 + i < p.N)               // N % 4 == 0: a vector is entirely inside or outside
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + i), "r"(acc[i]), "r"(acc[i + 1]), "r"(acc[i + 2]), "r"(acc[i + 3]) : "memory");
    }
  } else {
    float2 v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = make_float2(__uint_as_float(acc[2 * i]), __uint_as_float(acc[2 * i + 1]));
    if (EPI != EPI_DGELU && p.bias) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float4 b;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "r"(bias_addr + 16 * i));
        v[2 * i] = __fadd2_rn(v[2 * i], make_float2(b.x, b.y));
        v[2 * i + 1] = __fadd2_rn(v[2 * i + 1], make_float2(b.z, b.w));
      }
    }
    const float2 ds = make_float2(p.drop_scale, p.drop_scale);
    if (EPI == EPI_GELU) {
      if (p.preact_out) {
        box_store_begin<1>(lane);   // the output store of the previous chunk may still be reading box1
        box_store_row(box0, lane, v);
        box_store_issue(tmap_aux, box0, lane, n0, m0, bt);
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = gelu2(v[i]);
      if (p.drop_p > 0.f) {
        drop32(v, rh, col_addr, p.drop_thresh);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __fmul2_rn(v[i], ds);
      }
      box_store_begin<1>(lane);     // this chunk's pre-activation store may still be reading box0
    } else if (EPI == EPI_DGELU) {
      float2 u[16];
      unpack32(aux, u);
      if (p.act == 2) {   // ReLU backward: aux = forward output
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = make_float2(u[i].x > 0.f ? v[i].x : 0.f, u[i].y > 0.f ? v[i].y : 0.f);
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __fmul2_rn(v[i], gelu_grad2(u[i]));
      }
      if (p.drop_p > 0.f) {
        drop32(v, rh, col_addr, p.drop_thresh);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __fmul2_rn(v[i], ds);
      }
      box_store_begin<0>(lane);
    } else {  // EPI_LINEAR: [bias] [ReLU] [dropout] [+ residual]
      if (p.act == 2) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = make_float2(fmaxf(v[i].x, 0.f), fmaxf(v[i].y, 0.f));
      }
      if (p.drop_p > 0.f) drop32(v, rh, col_addr, p.drop_thresh);
      if (p.residual) {
        float2 rs[16];
        unpack32(aux, rs);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __ffma2_rn(v[i], ds, rs[i]);   // ds = 1 without dropout
      } else if (p.drop_p > 0.f) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __fmul2_rn(v[i], ds);
      }
      box_store_begin<0>(lane);
    }
    box_store_row(box1, lane, v);
    box_store_issue(tmap_out, box1, lane, n0, m0, bt);
  }
}

// CG = 1: one CTA per 128 x tile_n tile.  CG = 2: a CTA pair (cluster of 2) per 256 x tile_n tile with
// tcgen05.mma.cta_group::2 — each CTA stages its own 128 rows of A and tile_n/2 rows (columns of D) of
// B, which halves the L2 -> SMEM operand traffic per FLOP (the 1-CTA tile is L2-bandwidth bound).
template <int CG, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                         const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_aux,
                         const __grid_constant__ GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  // control block: barriers + TMEM base pointer; stage ring starts at the next 1024-byte boundary
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);  // [0..7] full, [8..15] empty, [16,17] tmem_full, [18,19] tmem_empty
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + 256);
  float* s_bias = reinterpret_cast<float*>(smem + 1024);          // [2][256] bias of the tile's columns (per accumulator stage)
  uint32_t* s_col = reinterpret_cast<uint32_t*>(smem + 3072);     // [2][256] odd dropout hashes of the tile's columns
  uint8_t* s_io = smem + GEMM_CTRL_BYTES;                          // [warp][2] boxes of 2 KB (1024-aligned)
  uint8_t* ring = smem + GEMM_CTRL_BYTES + GEMM_IO_BYTES;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;   // CTA rank inside the pair
  const bool leader = rank == 0;
  const uint32_t bar_base = smem_u32(bars);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (8 + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (16 + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (18 + s); };
  auto aux_bar = [&](int w) { return bar_base + 8u * (20 + w); };   // per epilogue warp: residual / pre-activation box landed

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), GEMM_EPI_WARPS * CG);
    }
    for (int w = 0; w < GEMM_EPI_WARPS; ++w) mbar_init(aux_bar(w), 1);
    fence_mbar_init();
  }
  if (warp == GEMM_EPI_WARPS + 1) {
    if (CG == 2) { tmem_alloc_cg2(smem_u32(tmem_ptr_smem), 512); tmem_relinquish_cg2(); }
    else { tmem_alloc(smem_u32(tmem_ptr_smem), 512); tmem_relinquish(); }
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr_smem, 0);   // warp-uniform for the compiler

  const int num_items = p.num_m_tiles * p.num_n_tiles * p.split_k * (p.batched ? p.batched : 1);   // m-tiles are 128*CG rows tall
  const int first_item = blockIdx.x / CG, item_stride = gridDim.x / CG;
  const int cta_b_rows = p.tile_n / CG;                              // B rows (D columns) staged by this CTA

  // Role -> warp mapping: the warp scheduler favours HIGHER warp ids, so the latency-critical single-thread
  // roles (TMA producer, MMA issuer) take the two highest warps and the epilogue warps the low ones.
  if (warp == GEMM_EPI_WARPS) {
    // ===================== TMA producer (every CTA) =====================
    // elect.sync rather than `lane == 0`: the compiler then knows one thread is active and emits bare
    // UTMALDG / UTCHMMA sequences instead of an ELECT + BRA.U.ANY loop around each of them
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = first_item; item < num_items; item += item_stride) {
        int mt, nt, kb0, nkb, b1, b2;
        decode_item(p, item, mt, nt, kb0, nkb, b1, b2);
        const int m0 = mt * GEMM_BM * CG + rank * GEMM_BM;
        const int n0 = nt * p.tile_n + rank * cta_b_rows;
        for (int kb = 0; kb < nkb; ++kb) {
          const int k0 = (kb0 + kb) * GEMM_BK;
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sa = smem_u32(ring + stage * p.stage_bytes);
          const uint32_t sb = sa + GEMM_A_BYTES;
          const uint32_t fb = full_bar(stage);
          if (leader) mbar_expect_tx(fb, CG * (GEMM_A_BYTES + p.b_bytes));
          if (p.batched) {   // 4-D tensor maps: (column, row, b2, b1)
            if (CG == 2) {
              if (!p.a_mn) {
                tma_load_4d_cg2(sa, &tmap_a, fb, k0, m0, b2, b1);
              } else {
                tma_load_4d_cg2(sa, &tmap_a, fb, m0, k0, b2, b1);
                tma_load_4d_cg2(sa + 8192, &tmap_a, fb, m0 + 64, k0, b2, b1);
              }
              if (!p.b_mn) {
                tma_load_4d_cg2(sb, &tmap_b, fb, k0, n0, b2, b1);
              } else {
                for (int c = 0; c * 64 < cta_b_rows; ++c) tma_load_4d_cg2(sb + c * 8192, &tmap_b, fb, n0 + c * 64, k0, b2, b1);
              }
            } else {
              if (!p.a_mn) {
                tma_load_4d(sa, &tmap_a, fb, k0, m0, b2, b1);
              } else {
                tma_load_4d(sa, &tmap_a, fb, m0, k0, b2, b1);
                tma_load_4d(sa + 8192, &tmap_a, fb, m0 + 64, k0, b2, b1);
              }
              if (!p.b_mn) {
                tma_load_4d(sb, &tmap_b, fb, k0, n0, b2, b1);
              } else {
                for (int c = 0; c * 64 < cta_b_rows; ++c) tma_load_4d(sb + c * 8192, &tmap_b, fb, n0 + c * 64, k0, b2, b1);
              }
            }
          } else if (CG == 2) {
            if (!p.a_mn) {
              tma_load_2d_cg2(sa, &tmap_a, fb, k0, m0);
            } else {
              tma_load_2d_cg2(sa, &tmap_a, fb, m0, k0);
              tma_load_2d_cg2(sa + 8192, &tmap_a, fb, m0 + 64, k0);
            }
            if (!p.b_mn) {
              tma_load_2d_cg2(sb, &tmap_b, fb, k0, n0);
            } else {
              for (int c = 0; c * 64 < cta_b_rows; ++c) tma_load_2d_cg2(sb + c * 8192, &tmap_b, fb, n0 + c * 64, k0);
            }
          } else {
            if (!p.a_mn) {
              tma_load_2d(sa, &tmap_a, fb, k0, m0);
            } else {
              tma_load_2d(sa, &tmap_a, fb, m0, k0);
              tma_load_2d(sa + 8192, &tmap_a, fb, m0 + 64, k0);
            }
            if (!p.b_mn) {
              tma_load_2d(sb, &tmap_b, fb, k0, n0);
            } else {
              for (int c = 0; c * 64 < cta_b_rows; ++c) tma_load_2d(sb + c * 8192, &tmap_b, fb, n0 + c * 64, k0);
            }
          }
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == GEMM_EPI_WARPS + 1) {
    // ===================== MMA issuer (even CTA of the pair only) =====================
    if (leader && elect_one()) {
      const uint32_t idesc = make_idesc_bf16(p.tile_n, p.a_mn, p.b_mn, 128 * CG);
      // descriptor templates: only the start-address word advances (per k-step and per stage)
      const uint64_t dta = make_smem_desc(0, p.a_mn ? 8192u : 16u, 1024);
      const uint64_t dtb = make_smem_desc(0, p.b_mn ? 8192u : 16u, 1024);
      const uint32_t a_hi = desc_hi(dta), b_hi = desc_hi(dtb);
      const uint32_t a_step = (p.a_mn ? 2048u : 32u) >> 4, b_step = (p.b_mn ? 2048u : 32u) >> 4;
      const uint32_t ring_lo = smem_u32(ring) >> 4, stage_lo = static_cast<uint32_t>(p.stage_bytes) >> 4;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int item = first_item; item < num_items; item += item_stride) {
        int mt, nt, kb0, nkb, b1, b2;
        decode_item(p, item, mt, nt, kb0, nkb, b1, b2);
        if (nkb == 0) continue;
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * p.tile_n;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t a_lo = desc_lo(dta) + ring_lo + stage * stage_lo;
          const uint32_t b_lo = desc_lo(dtb) + ring_lo + stage * stage_lo + (GEMM_A_BYTES >> 4);
          if (CG == 2) {
            umma_k4_cg2(d_tmem, a_hi, a_lo, a_step, b_hi, b_lo, b_step, idesc, kb != 0);
            umma_commit_cg2_mc(empty_bar(stage), 0x3);
          } else {
            umma_k4(d_tmem, a_hi, a_lo, a_step, b_hi, b_lo, b_step, idesc, kb != 0);
            umma_commit(empty_bar(stage));
          }
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        if (CG == 2) umma_commit_cg2_mc(tfull_bar(acc), 0x3); else umma_commit(tfull_bar(acc));
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue warps (every CTA: its own 128 accumulator rows) =====================
    const int quad = warp & 3;            // TMEM lane quadrant this warp may access
    const int half = warp >> 2;           // which warp of the quadrant: takes chunks half, half + GROUPS, ...
    int acc = 0;
    uint32_t acc_phase = 0, aux_phase = 0;
    for (int item = first_item; item < num_items; item += item_stride) {
      int mt, nt, kb0, nkb, b1, b2;
      decode_item(p, item, mt, nt, kb0, nkb, b1, b2);
      if (nkb == 0) continue;
      const int bt = p.batched ? ((b1 << 16) | b2) : -1;   // batch coordinates of the 4-D output / operand boxes
      constexpr bool FAST = EPI != EPI_GENERIC;
      constexpr bool STAGED = EPI == EPI_LINEAR || EPI == EPI_GELU || EPI == EPI_DGELU;
      constexpr bool AUX_IN = EPI == EPI_LINEAR || EPI == EPI_DGELU;   // residual / saved pre-activation arrive by TMA
      const int m0 = mt * GEMM_BM * CG + rank * GEMM_BM + quad * 32;   // first of this warp's 32 rows
      const int m = m0 + lane;
      const bool has_aux = AUX_IN && (EPI == EPI_DGELU || p.residual != nullptr) && m0 < p.M;
      const uint32_t box0 = smem_u32(s_io + warp * 4096), box1 = box0 + 2048;
      uint32_t rh = 0;
      if (STAGED) {
        // per-tile column constants, staged while the MMAs of this tile are still running.  Buffer `acc` was last
        // read two tiles ago; every warp passed the previous tile's barrier since then.
        const int et = threadIdx.x;   // epilogue threads are 0 .. 32 * GEMM_EPI_WARPS - 1
        if (et < p.tile_n) {
          const int n = nt * p.tile_n + et;
          s_bias[acc * 256 + et] = (p.bias && n < p.N) ? __ldg(p.bias + n) : 0.f;
          s_col[acc * 256 + et] = drop_colodd(static_cast<uint32_t>(n));
        }
        if (p.drop_p > 0.f) rh = drop_rowhash(p.drop_seed, static_cast<uint64_t>(m));
        if (has_aux && lane == 0 && half * 32 < p.tile_n && nt * p.tile_n + half * 32 < p.N) {   // first chunk's operand box, before the MMAs finish
          mbar_expect_tx(aux_bar(warp), 2048);
          box_load_issue(&tmap_aux, box0, aux_bar(warp), nt * p.tile_n + half * 32, m0, bt);
        }
        asm volatile("bar.sync 1, %0;" ::"n"(32 * GEMM_EPI_WARPS) : "memory");
      }
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * p.tile_n;
      bool released = false;
      for (int c = half * 32; c < p.tile_n; c += 32 * GEMM_EPI_GROUPS) {
        const int n0 = nt * p.tile_n + c;
        const bool live = m0 < p.M && n0 < p.N;   // warp-uniform
        uint4 aux[4] = {};
        if (has_aux && n0 < p.N) {
          mbar_wait(aux_bar(warp), aux_phase);
          aux_phase ^= 1;
          box_load_row(box0, lane, aux);
          const int cn = c + 32 * GEMM_EPI_GROUPS;   // this warp's next chunk of the tile: request it now
          __syncwarp();
          if (lane == 0 && cn < p.tile_n && nt * p.tile_n + cn < p.N) {
            fence_proxy_async_smem();
            mbar_expect_tx(aux_bar(warp), 2048);
            box_load_issue(&tmap_aux, box0, aux_bar(warp), nt * p.tile_n + cn, m0, bt);
          }
        }
        uint32_t r[32];
        tmem_ld32(t_row + c, r);
        tmem_ld_wait();
        if (c + 32 * GEMM_EPI_GROUPS >= p.tile_n) {
          // this warp's share of the accumulator is in registers: hand the TMEM stage back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (CG == 2) mbar_arrive_even_cta(tempty_bar(acc)); else mbar_arrive(tempty_bar(acc));
          }
          released = true;
        }
        if (FAST) {
          if (live) epilogue_fast<EPI>(p, &tmap_out, &tmap_aux, lane, m0, n0, r, aux, box0, box1, smem_u32(s_bias + acc * 256 + c),
                                       smem_u32(s_col + acc * 256 + c), rh, bt);
        } else if (n0 < p.N) {
          epilogue_chunk(p, m, n0, r);
        }
      }
      if (!released) {  // tile_n == 32: the second warp of the quadrant has no chunk
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (CG == 2) mbar_arrive_even_cta(tempty_bar(acc)); else mbar_arrive(tempty_bar(acc));
        }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (EPI != EPI_GENERIC && EPI != EPI_RED && lane == 0) tma_store_wait_all<0>();   // staged boxes fully written out
  }

  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == GEMM_EPI_WARPS + 1) {
    tc_fence_after();
    if (CG == 2) tmem_dealloc_cg2(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
  }
}

static int pick_tile_n(long long N) {
  // largest tile in {256,224,192,160,128} that divides N; else minimise padded columns
  const int cands[] = {256, 224, 192, 160, 128};
  for (int c : cands)
    if (N % c == 0) return c;
  if (N <= 256) return static_cast<int>(((N + 31) / 32) * 32);
  int best = 256;
  long long best_pad = -1;
  for (int c : cands) {
    long long pad = ((N + c - 1) / c) * c - N;
    if (best_pad < 0 || pad < best_pad) { best = c; best_pad = pad; }
  }
  return best;
}

template <int CG, int EPI>
static int launch_gemm(int ctas, int smem_bytes, cudaStream_t stream, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc,
                       const CUtensorMap& td, const GemmParams& p) {
  static DeviceOnce once;   // the attribute is per device: one opt-in per (kernel instantiation, device)
  if (int rc = once.run([] { XF_CUDA(cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<CG, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); return 0; }))
    return rc;
  if (CG == 1) {
    gemm_bf16_tcgen05_kernel<CG, EPI><<<ctas, GEMM_THREADS, smem_bytes, stream>>>(ta, tb, tc, td, p);
    return 0;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(ctas);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  XF_CUDA(cudaLaunchKernelEx(&cfg, gemm_bf16_tcgen05_kernel<CG, EPI>, ta, tb, tc, td, p));
  return 0;
}

}  // namespace xf

namespace xf { thread_local int g_gemm_cta_cap = 0; }

// Caps the persistent grid of every xf_gemm issued by THIS host thread whose max_ctas is 0 (0 = no cap).  Lets a caller
// that runs independent problems on several streams give each a fixed share of the SMs (spatial partition) instead of
// letting full-machine persistent grids queue behind one another.
extern "C" int xf_set_gemm_cta_cap(int ctas) {
  const int prev = xf::g_gemm_cta_cap;
  xf::g_gemm_cta_cap = ctas > 0 ? ctas : 0;
  return prev;
}

extern "C" int xf_gemm(const XfGemm* g, xf_stream_t stream_) {
  using namespace xf;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!g || !g->a || !g->b || !g->out) return fail(-1, "xf_gemm: null pointer");
  if (g->M <= 0 || g->N <= 0 || g->K <= 0) return fail(-2, "xf_gemm: bad shape M=%lld N=%lld K=%lld", (long long)g->M, (long long)g->N, (long long)g->K);
  int tile_n = g->tile_n > 0 ? g->tile_n : pick_tile_n(g->N);
  if (tile_n % 32 != 0 || tile_n < 32 || tile_n > 256) return fail(-3, "xf_gemm: tile_n %d not a multiple of 32 in [32,256]", tile_n);
  int split_k = g->split_k > 1 ? g->split_k : 1;
  if (split_k > 1 && !(g->out_dtype == 1 && g->accumulate == 1)) return fail(-4, "xf_gemm: split_k needs fp32 atomic accumulation");
  if (g->accumulate && g->out_dtype != 1) return fail(-5, "xf_gemm: accumulate needs fp32 output");
  if (g->drop_p < 0.f || g->drop_p >= 1.f) return fail(-6, "xf_gemm: drop_p out of range");
  if (g->act < 0 || g->act > 2) return fail(-6, "xf_gemm: act must be 0 (none), 1 (GELU) or 2 (ReLU)");

  const int nb1 = g->batch1 > 0 ? g->batch1 : 0, nb2 = g->batch2 > 0 ? g->batch2 : (nb1 > 0 ? 1 : 0);
  const bool batched = nb1 > 0;
  if (batched && (nb1 > 32767 || nb2 > 65535)) return fail(-9, "xf_gemm: batch counts out of range");

  // CTA pairs (cta_group::2) unless disabled or the problem is a single small tile
  int cg = g->cta_group == 1 ? 1 : 2;
  if (g->cta_group == 0 && g->M <= 128) cg = 1;
  if (cg == 2 && (tile_n % 32 != 0 || (tile_n / 2) % 8 != 0)) cg = 1;

  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = (int)g->M; p.N = (int)g->N; p.K = (int)g->K;
  p.tile_n = tile_n;
  p.a_mn = g->a_mn_major ? 1 : 0;
  p.b_mn = g->b_mn_major ? 1 : 0;
  p.num_m_tiles = (p.M + GEMM_BM * cg - 1) / (GEMM_BM * cg);
  p.num_n_tiles = (p.N + tile_n - 1) / tile_n;
  p.total_kb = (p.K + GEMM_BK - 1) / GEMM_BK;
  if (split_k > p.total_kb) split_k = p.total_kb;
  p.kb_per_split = (p.total_kb + split_k - 1) / split_k;
  split_k = (p.total_kb + p.kb_per_split - 1) / p.kb_per_split;  // no empty splits
  p.split_k = split_k;
  p.batched = batched ? nb1 * nb2 : 0;
  p.nb2 = batched ? nb2 : 1;
  p.items_per_batch = p.num_m_tiles * p.num_n_tiles * p.split_k;
  p.out_bs1 = g->out_bs1; p.out_bs2 = g->out_bs2;
  const int cta_b_rows = tile_n / cg;
  p.b_bytes = p.b_mn ? ((cta_b_rows + 63) / 64) * 8192 : cta_b_rows * 128;   // per CTA
  p.stage_bytes = GEMM_A_BYTES + ((p.b_bytes + 1023) / 1024) * 1024;
  p.stages = (227 * 1024 - 1024 - GEMM_CTRL_BYTES - GEMM_IO_BYTES) / p.stage_bytes;
  if (p.stages > 8) p.stages = 8;
  p.bias = g->bias;
  p.pos_table = g->pos_table;
  p.rows_in = (int)g->rows_in; p.rows_out = (int)g->rows_out; p.row_off = (int)g->row_off;
  p.act = g->act;
  p.preact_out = reinterpret_cast<__nv_bfloat16*>(g->preact_out);
  p.dact_in = reinterpret_cast<const __nv_bfloat16*>(g->dact_in);
  p.residual = reinterpret_cast<const __nv_bfloat16*>(g->residual);
  p.ldr = g->ldr;
  p.out = g->out;
  p.ldc = g->ldc;
  p.out_f32 = g->out_dtype == 1;
  p.accumulate = g->accumulate;
  const int esz = p.out_f32 ? 4 : 2;
  p.vec_ok = ((reinterpret_cast<uintptr_t>(g->out) & 15) == 0) && ((g->ldc * esz) % 16 == 0) && (p.N % 4 == 0) &&
             (!g->residual || (((reinterpret_cast<uintptr_t>(g->residual) & 15) == 0) && ((g->ldr * 2) % 16 == 0))) &&
             (!g->preact_out || (((reinterpret_cast<uintptr_t>(g->preact_out) & 15) == 0) && ((g->ldc * 2) % 16 == 0))) &&
             (!g->dact_in || (((reinterpret_cast<uintptr_t>(g->dact_in) & 15) == 0) && ((g->ldc * 2) % 16 == 0))) &&
             (!g->pos_table || ((reinterpret_cast<uintptr_t>(g->pos_table) & 15) == 0)) &&
             (!g->bias || ((reinterpret_cast<uintptr_t>(g->bias) & 15) == 0));
  p.drop_p = g->drop_p;
  p.drop_seed = drop_key(g->drop_seed, g->drop_stream);  // per-call key
  p.drop_stream = g->drop_stream;
  p.drop_first = g->drop_first;
  p.drop_thresh = drop_thresh32(g->drop_p);
  p.drop_scale = g->drop_p > 0.f ? 1.0f / (1.0f - g->drop_p) : 1.0f;
  if (split_k > 1 && (g->bias || g->pos_table || g->act || g->dact_in || g->residual || g->drop_p > 0.f))
    return fail(-8, "xf_gemm: split_k cannot be combined with a non-linear / additive epilogue");

  CUtensorMap ta, tb;
  int rc;
  if (batched) {
    if (!p.a_mn) rc = make_tmap_4d_bf16(&ta, g->a, nb1, nb2, p.M, p.K, g->a_ld, g->a_bs1, g->a_bs2, 64, 128);
    else         rc = make_tmap_4d_bf16(&ta, g->a, nb1, nb2, p.K, p.M, g->a_ld, g->a_bs1, g->a_bs2, 64, 64);
    if (rc) return rc;
    if (!p.b_mn) rc = make_tmap_4d_bf16(&tb, g->b, nb1, nb2, p.N, p.K, g->b_ld, g->b_bs1, g->b_bs2, 64, cta_b_rows);
    else         rc = make_tmap_4d_bf16(&tb, g->b, nb1, nb2, p.K, p.N, g->b_ld, g->b_bs1, g->b_bs2, 64, 64);
    if (rc) return rc;
  } else {
    if (!p.a_mn) rc = make_tmap_2d_bf16(&ta, g->a, p.M, p.K, g->a_ld, 64, 128);
    else         rc = make_tmap_2d_bf16(&ta, g->a, p.K, p.M, g->a_ld, 64, 64);
    if (rc) return rc;
    if (!p.b_mn) rc = make_tmap_2d_bf16(&tb, g->b, p.N, p.K, g->b_ld, 64, cta_b_rows);
    else         rc = make_tmap_2d_bf16(&tb, g->b, p.K, p.N, g->b_ld, 64, 64);
    if (rc) return rc;
  }

  // epilogue kind: the specialised epilogues need full, 16-byte aligned 32-column chunks and no row remap
  // (a last chunk that sticks out of N is clipped by the TMA boxes; bias / hash staging stops at N)
  const bool plain = !g->pos_table && g->rows_in == 0 && p.vec_ok && p.N % 8 == 0 && p.tile_n <= 256;
  const bool drop_pre_act = g->drop_p > 0.f && g->drop_first;   // only matters when an activation follows the mask
  int epi = EPI_GENERIC;
  if (plain && p.out_f32) {
    if (p.accumulate && !g->bias && !g->act && !g->dact_in && !g->residual && !(g->drop_p > 0.f)) epi = EPI_RED;
  } else if (plain) {
    if (g->act == 1 && !g->dact_in && !g->residual && !drop_pre_act) epi = EPI_GELU;
    else if ((g->act == 0 || g->act == 2) && g->dact_in && !g->residual && !g->bias && !g->preact_out) epi = EPI_DGELU;
    else if ((g->act == 0 || g->act == 2) && !g->dact_in && !g->preact_out) epi = EPI_LINEAR;
  }

  // output / operand boxes of the specialised epilogues: [32 columns x 32 rows] bf16, SWIZZLE_64B
  CUtensorMap tc, td;
  memset(&tc, 0, sizeof(tc));
  memset(&td, 0, sizeof(td));
  if (batched && epi == EPI_GENERIC)
    return fail(-9, "xf_gemm: batched GEMM needs a specialised epilogue (no pos_table / row remap, N %% 8 == 0, aligned operands)");
  if (epi == EPI_LINEAR || epi == EPI_GELU || epi == EPI_DGELU) {
    const void* auxp = epi == EPI_GELU ? g->preact_out : epi == EPI_DGELU ? g->dact_in : g->residual;
    const long long auxld = epi == EPI_LINEAR ? g->ldr : g->ldc;
    if (batched) {   // aux operands share the output's batch strides
      if ((rc = make_tmap_4d_bf16(&tc, g->out, nb1, nb2, p.M, p.N, g->ldc, g->out_bs1, g->out_bs2, 32, 32, 64))) return rc;
      if (auxp && (rc = make_tmap_4d_bf16(&td, auxp, nb1, nb2, p.M, p.N, auxld, g->out_bs1, g->out_bs2, 32, 32, 64))) return rc;
    } else {
      if ((rc = make_tmap_2d_bf16(&tc, g->out, p.M, p.N, g->ldc, 32, 32, 64))) return rc;
      if (auxp && (rc = make_tmap_2d_bf16(&td, auxp, p.M, p.N, auxld, 32, 32, 64))) return rc;
    }
  }
  const int smem_bytes = 1024 /*align slack*/ + GEMM_CTRL_BYTES + GEMM_IO_BYTES + p.stages * p.stage_bytes;
  const int items = p.num_m_tiles * p.num_n_tiles * p.split_k * (batched ? nb1 * nb2 : 1);
  int sms = g->max_ctas > 0 ? g->max_ctas : sm_count();
  if (g->max_ctas <= 0 && g_gemm_cta_cap > 0 && g_gemm_cta_cap < sms) sms = g_gemm_cta_cap;
  if (cg == 1) {
    int ctas = sms < items ? sms : items;
    switch (epi) {
      case EPI_LINEAR: rc = launch_gemm<1, EPI_LINEAR>(ctas, smem_bytes, stream, ta, tb, tc, td, p); break;
      case EPI_GELU:   rc = launch_gemm<1, EPI_GELU>(ctas, smem_bytes, stream, ta, tb, tc, td, p); break;
      case EPI_DGELU:  rc = launch_gemm<1, EPI_DGELU>(ctas, smem_bytes, stream, ta, tb, tc, td, p); break;
      case EPI_RED:    rc = launch_gemm<1, EPI_RED>(ctas, smem_bytes, stream, ta, tb, tc, td, p); break;
      default:         rc = launch_gemm<1, EPI_GENERIC>(ctas, smem_bytes, stream, ta, tb, tc, td, p); break;
    }
  } else {
    int clusters = sms / 2 < items ? sms / 2 : items;
    if (clusters < 1) clusters = 1;
    switch (epi) {
      case EPI_LINEAR: rc = launch_gemm<2, EPI_LINEAR>(2 * clusters, smem_bytes, stream, ta, tb, tc, td, p); break;
      case EPI_GELU:   rc = launch_gemm<2, EPI_GELU>(2 * clusters, smem_bytes, stream, ta, tb, tc, td, p); break;
      case EPI_DGELU:  rc = launch_gemm<2, EPI_DGELU>(2 * clusters, smem_bytes, stream, ta, tb, tc, td, p); break;
      case EPI_RED:    rc = launch_gemm<2, EPI_RED>(2 * clusters, smem_bytes, stream, ta, tb, tc, td, p); break;
      default:         rc = launch_gemm<2, EPI_GENERIC>(2 * clusters, smem_bytes, stream, ta, tb, tc, td, p); break;
    }
  }
  if (rc) return rc;
  g_launches.fetch_add(1);
  XF_CUDA(cudaGetLastError());
  return 0;
}
