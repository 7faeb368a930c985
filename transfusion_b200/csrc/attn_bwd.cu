// xf_attn_bwd: fused attention backward for sm_100a (recomputes the probabilities from Q, K and the saved
// LSE; no S x S tensor).  autograd backward of torch18_adapters.py:789-798 (+ head split/merge :544-555,607).
//
// Measured on B200 (tools/microbench/mma_bench.cu): one tcgen05.mma (M=128, K=16) costs max(N/2, ~46) cycles
// whatever the operand source, so score tiles must be >= 64 wide, and with a 224-wide head the three
// accumulators dQ, dK, dV (3 x 224 fp32 columns) cannot share the 512 TMEM columns with them.  The backward
// is therefore three launches of one templated kernel, each owning ONE accumulator:
//
//   MODE_DQ (query-stationary): C1 = Q K_j^T, C2 = dO V_j^T,  E = dS   = P o (C2 - delta) * scale,  dQ += E K_j
//   MODE_DK (key-stationary)  : C1 = K Q_i^T, C2 = V dO_i^T,  E = dS^T,                             dK += E Q_i
//   MODE_DV (key-stationary)  : C1 = K Q_i^T,                 E = P^T (dropout applied),            dV += E dO_i
//
// Per CTA (128 resident rows, one (batch, head)); 320 threads = warps 0-7 element-wise (TMEM lane quadrant w & 3,
// 32 of the 64 streamed columns each), warp 8 TMA producer, warp 9 MMA issuer:
//   * the first resident operand R1 (Q or K) is staged once through shared memory and copied into TENSOR
//     MEMORY with tcgen05.cp (dp/2 columns); C1 is a ".ts" MMA (A from TMEM), so R1 costs no shared memory
//     and no shared-memory bandwidth while streaming;
//   * the second resident operand R2 (dO or V) stays in shared memory (SWIZZLE_64B, 32-column chunks);
//   * streamed 64-row tiles T1/T2 arrive by TMA, one 4-D box instruction per tile (all 32-column chunks; issuing
//     seven 4 KB boxes per tile paced the producer): the tile that is also the B operand of the accumulate MMA
//     (read a second time MN-major) lives in a 3-stage ring, the tile that only feeds a score MMA in a
//     2-stage ring;
//   * the element-wise stage reads C1/C2 from TMEM (tcgen05.ld), and writes E as packed bf16 straight back to
//     TMEM (tcgen05.st): the accumulate MMA is ".ts" too, E never touches shared memory.
// TMEM columns (dp = 224): accumulator 224 | R1 112 | C1,C2 2 x 64 | E 32 = 496.
#include <stdlib.h>
#include <string.h>

#include "../../include/xfusion.h"
#include "host_common.cuh"
#include "ptx.cuh"

namespace xf {

constexpr int NB_BM = 128;   // resident rows per CTA
constexpr int NB_BN = 64;    // streamed rows per iteration
constexpr int NB_LONG = 3, NB_SHORT = 2;
constexpr int NB_EW_WARPS = 8;   // element-wise warps: two per TMEM lane quadrant, each owns 32 of the 64 streamed columns
constexpr int NB_THREADS = 32 * (NB_EW_WARPS + 2);
constexpr uint32_t NB_SW64 = 4;
// MODE_DVS (5-unit backward): key-stationary, BOTH score products (C1 = K Q_i^T, C2 = V dO_i^T), accumulates
//   dV += P_d^T dO_i  in TMEM and streams  E = scale * dS^T  (bf16, [b, h, key, query]) to a caller-provided workspace;
//   dQ = E^T K and dK = E Q then run as two batched GEMMs on the GEMM kernel: S^T and dP^T are computed once
//   (5 GEMM units for the whole backward instead of 8).  One [32 keys x 32 queries] TMA store box per warp and tile.
enum { MODE_DQ = 0, MODE_DK = 1, MODE_DV = 2, MODE_DVS = 3 };

struct AttnBwd2Params {
  int B, H, Sq, Sk, dp, nch, r_tiles, n_stream;
  float sl2, scale;
  const uint8_t* kpm;
  int kpm_start;
  const float* lse;
  const float* delta;
  int stat_stride;
  __nv_bfloat16* out;
  long long ldo;
  float drop_p, drop_scale;
  float pd_mul;   // MODE_DVS: P_d / (1 - p) = (P * scale) * pd_mul
  uint32_t drop_key, drop_thresh;
  long long* dbg;  // dev aid: clock64 stamps of CTA 0 (null in production)
};

#define NB_STAMP(role, i, ev) do { if (p.dbg && blockIdx.x == 0 && (i) < 64) p.dbg[((role) * 64 + (i)) * 8 + (ev)] = clock64(); } while (0)

template <int MODE, bool DROP>
__global__ void __launch_bounds__(NB_THREADS, 1)
attn_bwd2_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_r1, const __grid_constant__ CUtensorMap tmap_r2,
                         const __grid_constant__ CUtensorMap tmap_t1, const __grid_constant__ CUtensorMap tmap_t2,
                         const __grid_constant__ CUtensorMap tmap_s1, const __grid_constant__ CUtensorMap tmap_s2,
                         const __grid_constant__ AttnBwd2Params p) {
  constexpr bool ROWQ = MODE == MODE_DQ;   // resident rows are queries (else keys)
  constexpr bool HAS2 = MODE != MODE_DV;   // second score product C2 = R2 T2^T
  constexpr bool STORE = MODE == MODE_DVS; // E also goes to the workspace (tmap_s1), one 32 x 32 box per warp
  constexpr bool ACC_T2 = MODE == MODE_DV || MODE == MODE_DVS;   // the accumulate MMA's B operand is T2 (dO_i): T2 lives in the long ring
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + 256);
  float* s_stat = reinterpret_cast<float*>(smem + 512);   // [8 warps][96]: per-warp lse / delta / hashes of its 32 streamed columns
  const uint32_t r2_bytes = HAS2 ? p.nch * 8192u : 0u;    // [128 rows x 64 B] per 32-column chunk
  const uint32_t t_bytes = p.nch * 4096u;                 // [64 rows x 64 B] per chunk
  uint8_t* sR2 = smem + 3584;                              // control (512) + stats (3072), rounded up to 1024 below
  sR2 = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(sR2) + 1023) & ~uintptr_t(1023));
  uint8_t* sL = sR2 + r2_bytes;                            // long ring (tile also read MN-major by the accumulate MMA)
  uint8_t* sS = sL + NB_LONG * t_bytes;                    // short ring (tile only feeds a score MMA)
  uint8_t* sBox = sS + NB_SHORT * t_bytes;                 // MODE_DVS: [8 warps][32 rows x 64 B] SWIZZLE_64B store boxes
  uint8_t* sStage = sL;                                    // R1 staging (SWIZZLE_128B, 64-column chunks) aliases the rings
  const int nck = (p.dp + 63) / 64;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_u32(bars);
  const uint32_t R_FULL = bar0, R1_COPIED = bar0 + 8;
  auto L_FULL = [&](int s) { return bar0 + 8u * (2 + s); };
  auto L_EMPTY = [&](int s) { return bar0 + 8u * (5 + s); };
  auto S_FULL = [&](int s) { return bar0 + 8u * (8 + s); };
  auto S_EMPTY = [&](int s) { return bar0 + 8u * (10 + s); };
  auto C_FULL = [&](int s) { return bar0 + 8u * (12 + s); };
  auto C_EMPTY = [&](int s) { return bar0 + 8u * (14 + s); };
  const uint32_t E_FULL = bar0 + 8u * 16, E_EMPTY = bar0 + 8u * 17, ACC_DONE = bar0 + 8u * 18;

  int bid = blockIdx.x;
  const int rt = bid % p.r_tiles; bid /= p.r_tiles;
  const int hd = bid % p.H;
  const int b = bid / p.H;
  const int r0 = rt * NB_BM;
  const int n = p.n_stream;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_r1); tma_prefetch_desc(&tmap_r2);
    tma_prefetch_desc(&tmap_t1); tma_prefetch_desc(&tmap_t2);
    if (STORE) tma_prefetch_desc(&tmap_s1);
    mbar_init(R_FULL, 1);
    mbar_init(R1_COPIED, 1);
    for (int s = 0; s < NB_LONG; ++s) { mbar_init(L_FULL(s), 1); mbar_init(L_EMPTY(s), 1); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(S_FULL(s), 1); mbar_init(S_EMPTY(s), 1);
      mbar_init(C_FULL(s), 1); mbar_init(C_EMPTY(s), NB_EW_WARPS);
    }
    mbar_init(E_FULL, NB_EW_WARPS);
    mbar_init(E_EMPTY, 1);
    mbar_init(ACC_DONE, 1);
    fence_mbar_init();
  }
  if (warp == NB_EW_WARPS + 1) {
    tmem_alloc(smem_u32(tmem_ptr_smem), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr_smem, 0);   // warp-uniform for the compiler: no per-MMA R2UR waterfall
  const uint32_t tm_acc = tmem_base;
  const uint32_t tm_r1 = tmem_base + p.dp;                    // dp/2 columns
  const uint32_t tm_c = tmem_base + p.dp + p.dp / 2;          // DQ/DK: one buffer C1|C2 ; DV: two buffers of C1
  const uint32_t tm_e = tm_c + 128;                           // 32 columns: E as packed bf16 (A of the accumulate MMA)

  if (warp == NB_EW_WARPS) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      const int col0 = hd * p.dp;
      mbar_expect_tx(R_FULL, nck * 16384u + r2_bytes);
      for (int c = 0; c < nck; ++c) tma_load_3d(smem_u32(sStage + c * 16384), &tmap_r1, R_FULL, col0 + 64 * c, r0, b);
      if (HAS2)
        for (int c = 0; c < p.nch; ++c) tma_load_3d(smem_u32(sR2 + c * 8192), &tmap_r2, R_FULL, col0 + 32 * c, r0, b);
      mbar_wait(R1_COPIED, 0);   // the staging area is now free for the rings
      for (int i = 0; i < n; ++i) {
        const int ls = i % NB_LONG, ss = i % NB_SHORT;
        const uint32_t lpar = ((i / NB_LONG) & 1) ^ 1, spar = ((i / NB_SHORT) & 1) ^ 1;
        const uint32_t la = smem_u32(sL + ls * t_bytes), sa = smem_u32(sS + ss * t_bytes);
        if (!ACC_T2) {   // T1 -> long ring (score + accumulate), T2 -> short ring (score only)
          // the short-ring slot is released first (after the score MMAs of tile i-2, one accumulate MMA earlier than
          // the long-ring slot of tile i-3), so its load is requested first
          mbar_wait(S_EMPTY(ss), spar);
          mbar_expect_tx(S_FULL(ss), t_bytes);
          tma_load_4d(sa, &tmap_t2, S_FULL(ss), 0, i * NB_BN, hd * p.nch, b);   // all nch chunks in one instruction
          mbar_wait(L_EMPTY(ls), lpar);
          mbar_expect_tx(L_FULL(ls), t_bytes);
          tma_load_4d(la, &tmap_t1, L_FULL(ls), 0, i * NB_BN, hd * p.nch, b);
        } else {                 // T1 -> short ring (score only), T2 -> long ring (accumulate; MODE_DVS: score too)
          mbar_wait(S_EMPTY(ss), spar);
          mbar_expect_tx(S_FULL(ss), t_bytes);
          tma_load_4d(sa, &tmap_t1, S_FULL(ss), 0, i * NB_BN, hd * p.nch, b);
          mbar_wait(L_EMPTY(ls), lpar);
          mbar_expect_tx(L_FULL(ls), t_bytes);
          tma_load_4d(la, &tmap_t2, L_FULL(ls), 0, i * NB_BN, hd * p.nch, b);
        }
      }
    }
  } else if (warp == NB_EW_WARPS + 1) {
    // ===================== MMA issuer =====================
    // elect.sync (not `lane == 0`): the compiler then knows exactly one thread is active and emits bare UTCHMMA
    // sequences instead of an ELECT / BRA.U.ANY loop around every MMA (measured: 72 -> 48 cycles per N = 64 MMA)
    if (elect_one()) {
      const uint32_t idesc_c = make_idesc_bf16(NB_BN, 0, 0);
      const uint32_t idesc_acc = make_idesc_bf16(p.dp, 0, 1);
      const uint64_t dk = make_smem_desc(0, 16, 512, NB_SW64);      // K-major template
      const uint64_t dmn = make_smem_desc(0, 4096, 512, NB_SW64);   // MN-major template: 32-column chunks 4096 B apart
      const uint32_t hi_k = desc_hi(dk), hi_mn = desc_hi(dmn), lo_k = desc_lo(dk), lo_mn = desc_lo(dmn);
      const uint32_t r2lo = lo_k + (smem_u32(sR2) >> 4);
      const uint32_t l_base = smem_u32(sL) >> 4, s_base = smem_u32(sS) >> 4, t_lo = t_bytes >> 4;

      mbar_wait(R_FULL, 0);
      tc_fence_after();
      {  // R1: shared memory (SWIZZLE_128B staging) -> TMEM, one 128 x 32 B slice per k-step
        const uint32_t st = smem_u32(sStage);
        for (int kk = 0; kk < p.dp / 16; ++kk)
          tmem_cp_128x256b(tm_r1 + 8 * kk, make_smem_desc(st + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024));
        umma_commit(R1_COPIED);
      }
      auto do_acc = [&](int i) {
        const int ls = i % NB_LONG;
        mbar_wait(E_FULL, i & 1);
        if (MODE == MODE_DV) mbar_wait(L_FULL(ls), (i / NB_LONG) & 1);
        NB_STAMP(0, i + 1, 4);
        tc_fence_after();
        umma_ts_k4(tm_acc, tm_e, hi_mn, lo_mn + l_base + ls * t_lo, 64, idesc_acc, i != 0);   // K = 64 streamed rows
        umma_commit(L_EMPTY(ls));
        umma_commit(E_EMPTY);
      };
      for (int i = 0; i < n; ++i) {
        const int ls = i % NB_LONG, ss = i % NB_SHORT;
        const int cb = HAS2 ? 0 : (i & 1);
        const int cuse = HAS2 ? i : (i >> 1);
        NB_STAMP(0, i, 0);
        if (MODE != MODE_DV) mbar_wait(L_FULL(ls), (i / NB_LONG) & 1);   // a score product reads the long-ring tile
        NB_STAMP(0, i, 6);
        mbar_wait(S_FULL(ss), (i / NB_SHORT) & 1);
        NB_STAMP(0, i, 1);
        mbar_wait(C_EMPTY(cb), (cuse & 1) ^ 1);
        NB_STAMP(0, i, 2);
        tc_fence_after();
        const uint32_t c1 = tm_c + (HAS2 ? 0 : cb * 64), c2 = tm_c + 64;
        const uint32_t t1lo = lo_k + (!ACC_T2 ? l_base + ls * t_lo : s_base + ss * t_lo);
        const uint32_t t2lo = lo_k + (MODE == MODE_DVS ? l_base + ls * t_lo : s_base + ss * t_lo);
        for (int ch = 0; ch < p.nch; ++ch) {   // 32 head-dim columns (2 k-steps) per asm block
          umma_ts_k2(c1, tm_r1 + 16 * ch, hi_k, t1lo + ch * 256, 2, idesc_c, ch != 0);
          if (HAS2) umma_k2(c2, hi_k, r2lo + ch * 512, 2, hi_k, t2lo + ch * 256, 2, idesc_c, ch != 0);
        }
        umma_commit(C_FULL(cb));
        umma_commit(S_EMPTY(ss));
        NB_STAMP(0, i, 3);
        if (i >= 1) do_acc(i - 1);
        NB_STAMP(0, i, 5);
      }
      do_acc(n - 1);
      umma_commit(ACC_DONE);
    }
  } else {
    // ===================== element-wise stage + epilogue (warps 0-7) =====================
    // warp w: TMEM lane quadrant w & 3 (rows), column half w >> 2 (32 of the 64 streamed columns).  Two warps per
    // scheduler double the element-wise throughput, which otherwise paces the key-stationary passes.
    const int quad = warp & 3, half = warp >> 2;
    const int r = quad * 32 + lane;
    const int row_g = r0 + r;   // query (DQ) or key (DK/DV) index inside the batch element
    const uint32_t lane_sel = static_cast<uint32_t>(quad * 32) << 16;
    const long long stat_base = (static_cast<long long>(b) * p.H + hd) * p.stat_stride;
    const uint64_t bh = static_cast<uint64_t>(b * p.H + hd);
    // The element-wise stage paces all three passes, so it is written for instruction count: packed fp32 math
    // (fma/mul .f32x2), every scale folded into the exponent (P * scale = exp2(s * sl2 - (lse - log2 scale))),
    // per-tile column operands re-read as 128-bit shared-memory broadcasts, key-padding fixed up after the loop
    // on the (rare) tiles that need it, invalid key rows zeroed once in the epilogue.
    const float fold = ROWQ || HAS2 ? __log2f(p.scale) : __log2f(p.drop_scale);   // DQ/DK: dS carries `scale`; DV: P_d carries 1/(1-p)
    float lse_row = 0.f, delta_row = 0.f;
    bool row_valid = true;
    if (ROWQ) {
      if (row_g < p.Sq) { lse_row = p.lse[stat_base + row_g] - fold; delta_row = p.delta[stat_base + row_g]; }
    } else {
      row_valid = row_g < p.Sk && !(p.kpm && p.kpm[static_cast<long long>(b) * p.Sk + row_g] != 0);
    }
    // this thread's own hash: query hash (DQ) or (odd) key hash (DK/DV); the tile's column hashes come from shared memory
    const uint32_t h_row = !DROP ? 0u : ROWQ ? drop_rowhash(p.drop_key, bh * p.Sq + row_g) : drop_colhash(p.drop_key, static_cast<uint32_t>(row_g));
    const float2 sl2_2 = make_float2(p.sl2, p.sl2);
    const float2 ds_2 = make_float2(p.drop_scale, p.drop_scale);
    const float2 nlse_2 = make_float2(-lse_row, -lse_row), ndelta_2 = make_float2(-delta_row, -delta_row);
    const uint32_t t32 = p.drop_thresh;
    float* my_stat = s_stat + warp * 96;   // per warp: [32 lse | 32 delta | 32 hashes] of this warp's 32 streamed columns
    const uint32_t st_addr = smem_u32(my_stat);

    // key-stationary passes: the per-column statistics of tile i + 1 are requested one iteration ahead and only
    // touched (fold, shared-memory staging) an iteration later: warps issue in order, so an instruction that
    // consumes a load stalls everything behind it for the full memory latency
    float g_l = 0.f, g_d = 0.f;
    auto load_stats = [&](int i) {
      const int col = i * NB_BN + 32 * half + lane;
      const bool in = col < p.Sq;   // beyond the sequence: lse := +inf makes the probability exactly 0, delta := 0 keeps dS finite
      g_l = in ? __ldg(p.lse + stat_base + col) : INFINITY;
      if (HAS2) g_d = in ? __ldg(p.delta + stat_base + col) : 0.f;
    };
    if (!ROWQ) load_stats(0);

    for (int i = 0; i < n; ++i) {
      const int cb = HAS2 ? 0 : (i & 1);
      const int cuse = HAS2 ? i : (i >> 1);
      const int t0 = i * NB_BN + 32 * half;   // first streamed row of this warp's columns
      uint32_t bad = 0;
      __syncwarp();   // the previous iteration's shared-memory reads are done
      if (ROWQ) {
        const int key = t0 + lane;
        bool bk = key >= p.Sk;
        if (!bk && p.kpm && t0 + 32 > p.kpm_start) bk = p.kpm[static_cast<long long>(b) * p.Sk + key] != 0;
        bad = __ballot_sync(0xffffffffu, bk);
        if (DROP) my_stat[64 + lane] = __uint_as_float(drop_colhash(p.drop_key, static_cast<uint32_t>(key)));
      } else {
        my_stat[lane] = g_l - fold;
        if (HAS2) my_stat[32 + lane] = g_d;
        if (DROP) my_stat[64 + lane] = __uint_as_float(drop_rowhash(p.drop_key, bh * p.Sq + (t0 + lane)));
        if (i + 1 < n) load_stats(i + 1);
      }
      __syncwarp();
      if (warp == 0 && lane == 0) NB_STAMP(1, i, 0);
      mbar_wait(C_FULL(cb), cuse & 1);
      if (warp == 0 && lane == 0) NB_STAMP(1, i, 1);
      tc_fence_after();
      uint32_t c1[32], c2[32];
      tmem_ld32(tm_c + (HAS2 ? 0 : cb * 64) + lane_sel + 32 * half, c1);
      if (HAS2) tmem_ld32(tm_c + 64 + lane_sel + 32 * half, c2);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(C_EMPTY(cb));   // this warp's share of the score tile is in registers
      if (warp == 0 && lane == 0) NB_STAMP(1, i, 2);
      uint32_t pk[16];
      uint32_t ppk[STORE ? 16 : 1];   // MODE_DVS: P_d^T / (1 - p) of the same elements (the TMEM operand; pk = E goes to the workspace)
      const float2 pdm_2 = make_float2(p.pd_mul, p.pd_mul);
#pragma unroll
      for (int c = 0; c < 32; c += 4) {
        float4 st_l, st_d;
        uint4 hs;
        if (!ROWQ) asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(st_l.x), "=f"(st_l.y), "=f"(st_l.z), "=f"(st_l.w) : "r"(st_addr + 4 * c));
        if (!ROWQ && HAS2) asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(st_d.x), "=f"(st_d.y), "=f"(st_d.z), "=f"(st_d.w) : "r"(st_addr + 128 + 4 * c));
        if (DROP) asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(hs.x), "=r"(hs.y), "=r"(hs.z), "=r"(hs.w) : "r"(st_addr + 256 + 4 * c));
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {   // two column pairs per 128-bit broadcast
          const int cc = c + 2 * h2;
          const float2 nl = ROWQ ? nlse_2 : (h2 ? make_float2(-st_l.z, -st_l.w) : make_float2(-st_l.x, -st_l.y));
          const float2 x = __ffma2_rn(make_float2(__uint_as_float(c1[cc]), __uint_as_float(c1[cc + 1])), sl2_2, nl);
          float2 pr = make_float2(fast_exp2(x.x), fast_exp2(x.y));   // P * scale (DQ/DK) or P / (1 - p) (DV)
          bool k0 = true, k1 = true;
          if (DROP) {
            k0 = drop_keep_rc(h2 ? hs.z : hs.x, h_row, t32);   // (query hash) * (odd key hash), either orientation
            k1 = drop_keep_rc(h2 ? hs.w : hs.y, h_row, t32);
          }
          float2 e2;
          if (MODE == MODE_DV) {
            e2 = make_float2(k0 ? pr.x : 0.f, k1 ? pr.y : 0.f);
          } else {
            const float2 dpv = make_float2(k0 ? __uint_as_float(c2[cc]) : 0.f, k1 ? __uint_as_float(c2[cc + 1]) : 0.f);
            const float2 nd = ROWQ ? ndelta_2 : (h2 ? make_float2(-st_d.z, -st_d.w) : make_float2(-st_d.x, -st_d.y));
            e2 = __fmul2_rn(pr, __ffma2_rn(dpv, ds_2, nd));   // P * scale * (dP_dropped / (1 - p) - delta)
            if constexpr (STORE) {
              const float2 pd = __fmul2_rn(make_float2(k0 ? pr.x : 0.f, k1 ? pr.y : 0.f), pdm_2);
              ppk[cc >> 1] = pack_bf16(pd.x, pd.y);
            }
          }
          pk[cc >> 1] = pack_bf16(e2.x, e2.y);
        }
      }
      if (ROWQ && bad != 0) {   // masked / out-of-range keys (last tiles only): their dS is exactly 0
#pragma unroll
        for (int j = 0; j < 16; ++j)
          pk[j] &= (((bad >> (2 * j)) & 1u) ? 0u : 0x0000FFFFu) | (((bad >> (2 * j + 1)) & 1u) ? 0u : 0xFFFF0000u);
      }
      if constexpr (STORE) if (!row_valid) {   // padded / out-of-range key: its row of E is exactly 0 (dQ sums over it)
#pragma unroll
        for (int j = 0; j < 16; ++j) pk[j] = 0u;
      }
      if (warp == 0 && lane == 0) NB_STAMP(1, i, 3);
      // acc_{i-1} has consumed the previous E (for i = 0 the wait on the fresh barrier's opposite parity passes at
      // once; written without an `i > 0` test so the compiler does not peel a second copy of the loop body)
      mbar_wait(E_EMPTY, (i + 1) & 1);
      tc_fence_after();
      if (warp == 0 && lane == 0) NB_STAMP(1, i, 4);
      if constexpr (STORE) tmem_st16(tm_e + lane_sel + 16 * half, ppk);
      else tmem_st16(tm_e + lane_sel + 16 * half, pk);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(E_FULL);
      if (warp == 0 && lane == 0) NB_STAMP(1, i, 5);
      if constexpr (STORE) if (t0 < p.Sq) {
        // this warp's [32 keys x 32 queries] block of E -> workspace through a SWIZZLE_64B box (row-per-thread global
        // stores would cost 32 LSU wavefronts per instruction).  Off the MMA critical path: P_d^T is already in TMEM.
        const uint32_t box = smem_u32(sBox) + warp * 2048u;
        const uint32_t row = box + lane * 64, sw = (lane >> 1) & 3;
        if (lane == 0) tma_store_wait_read<0>();   // the previous tile's store has read the box (issued a tile ago)
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 4; ++j)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row + ((j ^ sw) << 4)), "r"(pk[4 * j]), "r"(pk[4 * j + 1]), "r"(pk[4 * j + 2]), "r"(pk[4 * j + 3]) : "memory");
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_4d(&tmap_s1, box, t0, r0 + quad * 32, 0, static_cast<int>(bh));
          tma_store_commit();
        }
      }
    }
    if (STORE && lane == 0) tma_store_wait_all<0>();   // workspace writes complete before the CTA retires

    // ---- epilogue: accumulator -> bf16 -> global (token-major, heads merged); the two warps of a quadrant
    //      alternate 32-column chunks
    mbar_wait(ACC_DONE, 0);
    tc_fence_after();
    const int limit = ROWQ ? p.Sq : p.Sk;
    const bool ok = row_g < limit;
    __nv_bfloat16* orow = p.out + (static_cast<long long>(b) * limit + row_g) * p.ldo + hd * p.dp;
    for (int c = 32 * half; c < p.dp; c += 64) {
      uint32_t o[32];
      tmem_ld32(tm_acc + lane_sel + c, o);
      tmem_ld_wait();
      if (!row_valid) {   // padded key: its dK / dV row is exactly 0 (the loop above does not mask rows)
#pragma unroll
        for (int k = 0; k < 32; ++k) o[k] = 0u;
      }
      if (ok) {
#pragma unroll
        for (int k = 0; k < 32; k += 8)
          *reinterpret_cast<uint4*>(orow + c + k) =
              make_uint4(pack_bf16(__uint_as_float(o[k]), __uint_as_float(o[k + 1])), pack_bf16(__uint_as_float(o[k + 2]), __uint_as_float(o[k + 3])),
                         pack_bf16(__uint_as_float(o[k + 4]), __uint_as_float(o[k + 5])), pack_bf16(__uint_as_float(o[k + 6]), __uint_as_float(o[k + 7])));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == NB_EW_WARPS + 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int MODE>
static int launch_bwd2(bool drop, int grid, int smem_bytes, cudaStream_t stream, const CUtensorMap& r1, const CUtensorMap& r2,
                       const CUtensorMap& t1, const CUtensorMap& t2, const CUtensorMap& s1, const CUtensorMap& s2,
                       const AttnBwd2Params& p) {
  static DeviceOnce once;
  if (int rc = once.run([] {
        XF_CUDA(cudaFuncSetAttribute(attn_bwd2_tcgen05_kernel<MODE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        XF_CUDA(cudaFuncSetAttribute(attn_bwd2_tcgen05_kernel<MODE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        return 0;
      }))
    return rc;
  if (drop) attn_bwd2_tcgen05_kernel<MODE, true><<<grid, NB_THREADS, smem_bytes, stream>>>(r1, r2, t1, t2, s1, s2, p);
  else attn_bwd2_tcgen05_kernel<MODE, false><<<grid, NB_THREADS, smem_bytes, stream>>>(r1, r2, t1, t2, s1, s2, p);
  g_launches.fetch_add(1);
  XF_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace xf

extern "C" int64_t xf_attn_bwd_workspace_bytes(int B, int H, int Sq, int Sk) {
  if (B <= 0 || H <= 0 || Sq <= 0 || Sk <= 0) return 0;
  const int64_t pitch = ((Sq + 63) / 64) * 64;
  return static_cast<int64_t>(B) * H * Sk * pitch * 2;   // E = scale * dS^T, bf16 [B, H, Sk, pitch]
}

extern "C" int xf_attn_bwd(const XfAttnBwd* a, xf_stream_t stream_) {
  using namespace xf;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!a || !a->q || !a->k || !a->v || !a->d_out || !a->lse || !a->delta || !a->dq || !a->dk || !a->dv)
    return fail(-1, "xf_attn_bwd: null pointer");
  if (a->dp % 32 || a->dp < 32 || a->dp > 224) return fail(-2, "xf_attn_bwd: padded head dim %d must be a multiple of 32 in [32,224]", a->dp);
  if (a->B <= 0 || a->H <= 0 || a->Sq <= 0 || a->Sk <= 0) return fail(-3, "xf_attn_bwd: bad shape");
  if (a->stat_stride % 64 || a->stat_stride < ((a->Sq + 63) / 64) * 64) return fail(-4, "xf_attn_bwd: stat_stride must be a multiple of 64 >= Sq rounded up to 64");
  if ((a->lddq % 8) || (a->lddk % 8) || (a->lddv % 8)) return fail(-5, "xf_attn_bwd: gradient leading dims must be multiples of 8");
  if (a->drop_p < 0.f || a->drop_p >= 1.f) return fail(-6, "xf_attn_bwd: drop_p out of range");

  AttnBwd2Params p;
  memset(&p, 0, sizeof(p));
  p.B = a->B; p.H = a->H; p.Sq = a->Sq; p.Sk = a->Sk; p.dp = a->dp; p.nch = a->dp / 32;
  p.sl2 = a->scale * 1.4426950408889634f;
  p.scale = a->scale;
  p.kpm = a->key_padding_mask;
  p.kpm_start = a->kpm_start;
  p.lse = a->lse; p.delta = a->delta; p.stat_stride = a->stat_stride;
  p.drop_p = a->drop_p;
  p.drop_scale = a->drop_p > 0.f ? 1.f / (1.f - a->drop_p) : 1.f;
  p.drop_key = drop_key(a->drop_seed, a->drop_stream);
  p.drop_thresh = drop_thresh32(a->drop_p);   // compared against the full hash word
  p.dbg = reinterpret_cast<long long*>(a->debug_timeline);
  const bool drop = a->drop_p > 0.f;

  const uint64_t cols = static_cast<uint64_t>(a->H) * a->dp;
  CUtensorMap q_stage, k_stage, do_res, v_res, k64, v64, q64, do64;
  int rc;
  if ((rc = make_tmap_3d_bf16(&q_stage, a->q, a->B, a->Sq, cols, a->ldq, 64, NB_BM, 128))) return rc;
  if ((rc = make_tmap_3d_bf16(&k_stage, a->k, a->B, a->Sk, cols, a->ldk, 64, NB_BM, 128))) return rc;
  if ((rc = make_tmap_3d_bf16(&do_res, a->d_out, a->B, a->Sq, cols, a->lddo, 32, NB_BM, 64))) return rc;
  if ((rc = make_tmap_3d_bf16(&v_res, a->v, a->B, a->Sk, cols, a->ldv, 32, NB_BM, 64))) return rc;
  if ((rc = make_tmap_chunks_bf16(&k64, a->k, a->B, a->Sk, cols, a->ldk, NB_BN, p.nch))) return rc;
  if ((rc = make_tmap_chunks_bf16(&v64, a->v, a->B, a->Sk, cols, a->ldv, NB_BN, p.nch))) return rc;
  if ((rc = make_tmap_chunks_bf16(&q64, a->q, a->B, a->Sq, cols, a->ldq, NB_BN, p.nch))) return rc;
  if ((rc = make_tmap_chunks_bf16(&do64, a->d_out, a->B, a->Sq, cols, a->lddo, NB_BN, p.nch))) return rc;

  const int ring = (NB_LONG + NB_SHORT) * p.nch * 4096;
  const int smem2 = 1024 + 4096 + p.nch * 8192 + ring;   // align slack + control/stats + R2 + rings
  const int smem1 = 1024 + 4096 + ring;
  const int q_tiles = (a->Sq + NB_BM - 1) / NB_BM, k_tiles = (a->Sk + NB_BM - 1) / NB_BM;
  CUtensorMap none;
  memset(&none, 0, sizeof(none));

  // ---- 5-unit path: one key-stationary pass (dV in TMEM, E = scale dS^T to the workspace), then dQ = E^T K and
  //      dK = E Q as batched GEMMs over (sample, head)
  const int64_t ws_need = xf_attn_bwd_workspace_bytes(a->B, a->H, a->Sq, a->Sk);
  if (a->workspace && a->workspace_bytes >= ws_need) {
    if (reinterpret_cast<uintptr_t>(a->workspace) & 15) return fail(-7, "xf_attn_bwd: workspace must be 16-byte aligned");
    const int64_t pitch = ((a->Sq + 63) / 64) * 64;
    __nv_bfloat16* ws_e = reinterpret_cast<__nv_bfloat16*>(a->workspace);
    CUtensorMap se;
    const uint64_t bh_n = static_cast<uint64_t>(a->B) * a->H;
    if ((rc = make_tmap_4d_bf16(&se, ws_e, bh_n, 1, a->Sk, a->Sq, pitch, a->Sk * pitch, 0, 32, 32, 64))) return rc;
    AttnBwd2Params pk = p;
    pk.r_tiles = k_tiles; pk.n_stream = (a->Sq + NB_BN - 1) / NB_BN;
    pk.out = reinterpret_cast<__nv_bfloat16*>(a->dv); pk.ldo = a->lddv;
    pk.pd_mul = p.drop_scale / a->scale;
    if ((rc = launch_bwd2<MODE_DVS>(drop, a->B * a->H * k_tiles, smem2 + NB_EW_WARPS * 2048, stream, k_stage, v_res, q64, do64, se, none, pk)))
      return rc;
    XfGemm g;
    memset(&g, 0, sizeof(g));
    g.batch1 = a->B; g.batch2 = a->H;
    g.a = ws_e; g.a_ld = pitch;
    g.a_bs1 = static_cast<int64_t>(a->H) * a->Sk * pitch; g.a_bs2 = a->Sk * pitch;
    g.b_mn_major = 1;
    g.b_bs2 = a->dp; g.out_bs2 = a->dp;
    g.N = a->dp;
    // dQ[b, q, h, :] = sum_k E[b, h, k, q] K[b, k, h, :]     (A stored [K][M])
    g.a_mn_major = 1;
    g.b = a->k; g.b_ld = a->ldk; g.b_bs1 = static_cast<int64_t>(a->Sk) * a->ldk;
    g.M = a->Sq; g.K = a->Sk;
    g.out = a->dq; g.ldc = a->lddq; g.out_bs1 = static_cast<int64_t>(a->Sq) * a->lddq;
    if ((rc = xf_gemm(&g, stream_))) return rc;
    // dK[b, k, h, :] = sum_q E[b, h, k, q] Q[b, q, h, :]     (A stored [M][K])
    g.a_mn_major = 0;
    g.b = a->q; g.b_ld = a->ldq; g.b_bs1 = static_cast<int64_t>(a->Sq) * a->ldq;
    g.M = a->Sk; g.K = a->Sq;
    g.out = a->dk; g.ldc = a->lddk; g.out_bs1 = static_cast<int64_t>(a->Sk) * a->lddk;
    return xf_gemm(&g, stream_);
  }
  {
    AttnBwd2Params pq = p;
    pq.r_tiles = q_tiles; pq.n_stream = (a->Sk + NB_BN - 1) / NB_BN;
    pq.out = reinterpret_cast<__nv_bfloat16*>(a->dq); pq.ldo = a->lddq;
    if ((rc = launch_bwd2<MODE_DQ>(drop, a->B * a->H * q_tiles, smem2, stream, q_stage, do_res, k64, v64, none, none, pq))) return rc;
  }
  {
    AttnBwd2Params pk = p;
    pk.r_tiles = k_tiles; pk.n_stream = (a->Sq + NB_BN - 1) / NB_BN;
    pk.out = reinterpret_cast<__nv_bfloat16*>(a->dk); pk.ldo = a->lddk;
    if (pk.dbg) pk.dbg += 2 * 64 * 8;
    if ((rc = launch_bwd2<MODE_DK>(drop, a->B * a->H * k_tiles, smem2, stream, k_stage, v_res, q64, do64, none, none, pk))) return rc;
    pk.out = reinterpret_cast<__nv_bfloat16*>(a->dv); pk.ldo = a->lddv;
    if (pk.dbg) pk.dbg += 2 * 64 * 8;
    if ((rc = launch_bwd2<MODE_DV>(drop, a->B * a->H * k_tiles, smem1, stream, k_stage, v_res, q64, do64, none, none, pk))) return rc;
  }
  return 0;
}
