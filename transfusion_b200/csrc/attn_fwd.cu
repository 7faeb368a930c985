// xf_attn_fwd: fused multi-head attention forward for sm_100a (flash-style, no S x S tensor).
//
// Replaces torch18_adapters.py:544-555 (head split), :578-597 (key-padding mask -> -inf), :789-798
// (_scaled_dot_product_attention: q/sqrt(d), bmm, softmax, dropout, bmm) and :607 (head merge).
// General Lq != Lk (the QKVEncoder-style cross-attention of cross_qkv_layers.py:70-77 is the same op).
//
// One CTA per (batch, head, 128-query tile); 320 threads:
//   warp 8    TMA producer: Q tile once, then K_j / V_j tiles (64 keys) into a 4-stage / 3-stage ring
//   warp 9    MMA issuer  : S_j = Q K_j^T  (tcgen05, M=128, N=64,  K=dp)  -> TMEM S[j&1]
//                           O  += P_j V_j  (tcgen05, M=128, N=dp,  K=64)  -> TMEM O
//   warps 0-7 softmax     : TMEM lane quadrant w & 3 (32 query rows), column half w >> 2 (32 of the 64 keys):
//                           tcgen05.ld the S half-row, key-padding / tail mask, row max exchanged with the partner
//                           warp through shared memory (named barrier of 64 threads), running max / partial sum
//                           in registers (log2 domain, lazy rescale of O in TMEM only when the max grows by
//                           > 2^8), dropout on P, P -> packed bf16 -> tcgen05.st back into the S buffer's TMEM
//                           columns; final O / l -> bf16 -> global (heads merged), LSE saved for backward.
// S is double-buffered in TMEM so QK^T of tile j+1 overlaps the softmax of tile j.
// Both A operands live in TENSOR MEMORY (".ts" MMAs): Q is staged once through shared memory (SWIZZLE_128B,
// aliasing the rings) and copied with tcgen05.cp (dp/2 columns); P never touches shared memory.
// K / V tiles: SWIZZLE_64B, 32-column chunks of [64 rows x 64 B] (no padding of a 224-wide head to 256), all chunks of
// a tile moved by ONE 4-D TMA instruction (seven 4 KB box instructions per tile paced the producer);
// K is the K-major B operand of S, V the MN-major B operand of PV.
// The softmax stage is written for instruction count (it paces the kernel together with the 16/clk exp2
// unit): packed fp32 math (fma/add .f32x2), product-form dropout hash with the tile's key hashes re-read as
// 128-bit shared-memory broadcasts, masking only on the tiles that need it.
#include <string.h>

#include "../../include/xfusion.h"
#include "host_common.cuh"
#include "ptx.cuh"

namespace xf {

constexpr int AF_BM = 128;
constexpr int AF_BN = 64;
constexpr int AF_SM_WARPS = 8;
constexpr int AF_THREADS = 32 * (AF_SM_WARPS + 2);
constexpr int AF_KST = 4, AF_VST = 3;        // ring depths
constexpr uint32_t AF_SW64 = 4;
constexpr float AF_RESCALE_THRESHOLD = 8.0f;  // log2 units

struct AttnFwdParams {
  int B, H, Sq, Sk, dp, nch, q_tiles;
  float sl2;  // log2(e) / sqrt(head_dim)
  const uint8_t* kpm;  // [B, Sk], 1 = ignore key; may be null
  int kpm_start;       // keys < kpm_start are never masked
  __nv_bfloat16* out;
  long long ldo;
  float* lse;  // [B, H, lse_stride], log2 domain: m + log2(l)
  int lse_stride;
  float drop_p, drop_scale;
  uint32_t drop_seed, drop_stream, drop_thresh;
  long long* dbg;  // dev aid: clock64 stamps of CTA 0 (null in production)
};

#define AF_STAMP(role, i, ev) do { if (p.dbg && blockIdx.x == 0 && (i) < 64) p.dbg[((role) * 64 + (i)) * 8 + (ev)] = clock64(); } while (0)

template <bool DROP>
__global__ void __launch_bounds__(AF_THREADS, 1)
attn_fwd_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                        const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ AttnFwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + 256);
  uint32_t* s_hash = reinterpret_cast<uint32_t*>(smem + 512);   // [8 warps][32]: key hashes of the warp's 32 columns
  float* s_xchg = reinterpret_cast<float*>(smem + 1536);        // [2][8 warps][32]: row max / row sum exchange
  const uint32_t t_bytes = p.nch * 4096u;                        // one K or V tile: nch x [64 rows x 64 B]
  uint8_t* sK = smem + 4096;
  uint8_t* sV = sK + AF_KST * t_bytes;
  uint8_t* sQ = sK;                                              // Q staging (SWIZZLE_128B, 64-column chunks) aliases the rings
  const int nck = (p.dp + 63) / 64;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_u32(bars);
  const uint32_t Q_FULL = bar0, Q_COPIED = bar0 + 8;
  auto K_FULL = [&](int s) { return bar0 + 8u * (2 + s); };
  auto K_EMPTY = [&](int s) { return bar0 + 8u * (6 + s); };
  auto V_FULL = [&](int s) { return bar0 + 8u * (10 + s); };
  auto V_EMPTY = [&](int s) { return bar0 + 8u * (13 + s); };
  auto S_FULL = [&](int s) { return bar0 + 8u * (16 + s); };
  const uint32_t P_FULL = bar0 + 8u * 18;
  const uint32_t O_READY = bar0 + 8u * 19;

  int bid = blockIdx.x;
  const int qt = bid % p.q_tiles; bid /= p.q_tiles;
  const int hd = bid % p.H;
  const int b = bid / p.H;
  const int q0 = qt * AF_BM;
  const int nkv = (p.Sk + AF_BN - 1) / AF_BN;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    mbar_init(Q_FULL, 1);
    mbar_init(Q_COPIED, 1);
    for (int s = 0; s < AF_KST; ++s) { mbar_init(K_FULL(s), 1); mbar_init(K_EMPTY(s), 1); }
    for (int s = 0; s < AF_VST; ++s) { mbar_init(V_FULL(s), 1); mbar_init(V_EMPTY(s), 1); }
    for (int s = 0; s < 2; ++s) mbar_init(S_FULL(s), 1);
    mbar_init(P_FULL, AF_SM_WARPS);
    mbar_init(O_READY, 1);
    fence_mbar_init();
  }
  if (warp == AF_SM_WARPS + 1) {
    tmem_alloc(smem_u32(tmem_ptr_smem), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr_smem, 0);   // warp-uniform for the compiler
  const uint32_t tmem_S = tmem_base;         // 2 x 64 columns
  const uint32_t tmem_O = tmem_base + 128;   // dp columns
  const uint32_t tmem_Q = tmem_base + 384;   // dp/2 columns: Q as packed bf16 (A operand of the score MMA)

  // The scheduler favours higher warp ids, so the single-thread roles sit above the ALU-heavy softmax warps.
  if (warp == AF_SM_WARPS) {
    if (elect_one()) {   // elect.sync, not `lane == 0`: bare UTMALDG / UTCHMMA sequences, no per-instruction ELECT loop
      const int col0 = hd * p.dp;
      mbar_expect_tx(Q_FULL, nck * 16384u);
      for (int c = 0; c < nck; ++c) tma_load_3d(smem_u32(sQ + c * 16384), &tmap_q, Q_FULL, col0 + 64 * c, q0, b);
      mbar_wait(Q_COPIED, 0);   // the staging area is now free for the rings
      for (int j = 0; j < nkv; ++j) {
        const int ks = j % AF_KST, vs = j % AF_VST;
        mbar_wait(K_EMPTY(ks), ((j / AF_KST) & 1) ^ 1);
        mbar_expect_tx(K_FULL(ks), t_bytes);
        tma_load_4d(smem_u32(sK + ks * t_bytes), &tmap_k, K_FULL(ks), 0, j * AF_BN, hd * p.nch, b);   // all chunks, one instruction
        mbar_wait(V_EMPTY(vs), ((j / AF_VST) & 1) ^ 1);
        mbar_expect_tx(V_FULL(vs), t_bytes);
        tma_load_4d(smem_u32(sV + vs * t_bytes), &tmap_v, V_FULL(vs), 0, j * AF_BN, hd * p.nch, b);
      }
    }
  } else if (warp == AF_SM_WARPS + 1) {
    if (elect_one()) {
      // descriptors are built once; only the start-address word changes per MMA, and a 32-column chunk's two
      // k-steps are issued from one asm block (issue rate matters for the N = 64 score MMAs)
      const uint32_t idesc_s = make_idesc_bf16(AF_BN, 0, 0);
      const uint32_t idesc_o = make_idesc_bf16(p.dp, 0, 1);
      const uint64_t dk = make_smem_desc(0, 16, 512, AF_SW64);       // K-major template
      const uint64_t dmn = make_smem_desc(0, 4096, 512, AF_SW64);    // MN-major template (V): 32-column chunks 4096 B apart
      const uint32_t hi_k = desc_hi(dk), hi_mn = desc_hi(dmn), lo_k = desc_lo(dk), lo_mn = desc_lo(dmn);
      const uint32_t k_base = smem_u32(sK) >> 4, v_base = smem_u32(sV) >> 4, t_lo = t_bytes >> 4;
      auto issue_s = [&](int j) {
        const int ks = j % AF_KST, sb = j & 1;
        AF_STAMP(0, j, 0);
        mbar_wait(K_FULL(ks), (j / AF_KST) & 1);
        AF_STAMP(0, j, 1);
        // S buffer sb last held P_{j-2}, consumed by PV_{j-2}: already issued by this thread, and the tensor pipe
        // runs in issue order, so no barrier is needed before overwriting it
        tc_fence_after();
        const uint32_t k_lo = lo_k + k_base + ks * t_lo;
        for (int ch = 0; ch < p.nch; ++ch)   // 32 head-dim columns (2 k-steps) per asm block
          umma_ts_k2(tmem_S + sb * AF_BN, tmem_Q + 16 * ch, hi_k, k_lo + ch * 256, 2, idesc_s, ch != 0);
        umma_commit(S_FULL(sb));
        umma_commit(K_EMPTY(ks));
        AF_STAMP(0, j, 3);
      };
      mbar_wait(Q_FULL, 0);
      tc_fence_after();
      {  // Q: shared memory -> TMEM, one 128 x 32 B slice per k-step
        const uint32_t qa = smem_u32(sQ);
        for (int kk = 0; kk < p.dp / 16; ++kk)
          tmem_cp_128x256b(tmem_Q + 8 * kk, make_smem_desc(qa + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024));
        umma_commit(Q_COPIED);
      }
      issue_s(0);
      for (int j = 0; j < nkv; ++j) {
        if (j + 1 < nkv) issue_s(j + 1);
        const int vs = j % AF_VST, sb = j & 1;
        AF_STAMP(0, j, 4);
        mbar_wait(P_FULL, j & 1);
        AF_STAMP(0, j, 5);
        mbar_wait(V_FULL(vs), (j / AF_VST) & 1);
        AF_STAMP(0, j, 6);
        tc_fence_after();
        umma_ts_k4(tmem_O, tmem_S + sb * AF_BN, hi_mn, lo_mn + v_base + vs * t_lo, 64, idesc_o, j != 0);   // A = P_j (TMEM, 32 columns)
        umma_commit(V_EMPTY(vs));
        umma_commit(O_READY);
        AF_STAMP(0, j, 7);
      }
    }
  } else {
    // ===================== softmax / correction / epilogue (warps 0-7) =====================
    const int quad = warp & 3, half = warp >> 2;
    const int r = quad * 32 + lane;     // row within the tile == TMEM lane
    const int q = q0 + r;
    const uint32_t lane_sel = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t pair_bar = 1 + quad;  // named barrier of the two warps that share these 32 rows
    float m_used = 0.f, l = 0.f;         // l: partial row sum over this warp's columns
    const uint32_t drop_rh = DROP ? drop_rowhash(p.drop_seed, static_cast<uint64_t>(b * p.H + hd) * p.Sq + q) : 0u;
    uint32_t* my_hash = s_hash + warp * 32;
    const uint32_t hash_addr = smem_u32(my_hash);
    const float2 sl2_2 = make_float2(p.sl2, p.sl2);
    const uint32_t t32 = p.drop_thresh;

    for (int j = 0; j < nkv; ++j) {
      const int sb = j & 1;
      const int k0 = j * AF_BN + 32 * half;   // first key of this warp's 32 columns
      // mask bits: bit c set -> key k0 + c is ignored (beyond Sk or key-padding); only tail / language tiles
      uint32_t bad = 0;
      __syncwarp();   // the previous tile's hash reads are done
      if (DROP) my_hash[lane] = drop_colhash(p.drop_seed, static_cast<uint32_t>(k0 + lane));
      if (k0 + 32 > p.Sk || (p.kpm && k0 + 32 > p.kpm_start)) {
        const int key = k0 + lane;
        bool bk = key >= p.Sk;
        if (!bk && p.kpm) bk = p.kpm[static_cast<long long>(b) * p.Sk + key] != 0;
        bad = __ballot_sync(0xffffffffu, bk);
      }
      __syncwarp();
      if (warp == 0 && lane == 0) AF_STAMP(1, j, 0);
      mbar_wait(S_FULL(sb), (j >> 1) & 1);
      if (warp == 0 && lane == 0) AF_STAMP(1, j, 1);
      tc_fence_after();
      uint32_t sr[32];
      tmem_ld32(tmem_S + lane_sel + sb * AF_BN + 32 * half, sr);
      tmem_ld_wait();
      if (warp == 0 && lane == 0) AF_STAMP(1, j, 2);
      float x[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) x[c] = __uint_as_float(sr[c]);
      if (bad != 0) {
#pragma unroll
        for (int c = 0; c < 32; ++c)
          if ((bad >> c) & 1u) x[c] = -INFINITY;
      }
      float mt = fmaxf(x[0], x[1]);
#pragma unroll
      for (int c = 2; c < 32; c += 2) mt = fmaxf(mt, fmaxf(x[c], x[c + 1]));
      mt *= p.sl2;   // log2 domain (sl2 > 0)
      // row max over all 64 keys: exchange with the partner warp (double-buffered by tile parity; the barrier also
      // orders the partner's S loads before this warp's P store below, which overlaps the partner's S columns)
      float* xb = s_xchg + (j & 1) * 256;
      xb[warp * 32 + lane] = mt;
      asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
      mt = fmaxf(mt, xb[(warp ^ 4) * 32 + lane]);
      bool need = false;
      float alpha = 1.f;
      int jv = j;
      asm volatile("" : "+r"(jv));   // opaque: no peeled copy of the loop body for the first tile
      if (jv == 0) {
        m_used = (mt == -INFINITY) ? 0.f : mt;
      } else {
        need = mt > m_used + AF_RESCALE_THRESHOLD;
      }
      const bool any_need = __any_sync(0xffffffffu, need);   // same rows, same values in both warps of the pair
      if (any_need) {
        const float m_new = fmaxf(m_used, mt);
        alpha = fast_exp2(m_used - m_new);
        l *= alpha;
        m_used = m_new;
      }
      // Groups of 8 keys.  ptxas would otherwise schedule all 32 exp2 back to back and all selects after them; both
      // warps of a scheduler then sit in the same phase and the 4-lane/clk exp2 unit, the FMA pipe and the ALU pipe
      // take turns instead of overlapping.  The formal dependency of group g+1's exponent offset on group g's last
      // result (fma(p, 0, -m) == -m: p is finite) keeps the groups in order, and the two warps fall out of phase.
      float2 negm_c = make_float2(-m_used, -m_used);
      const float2 zero_2 = make_float2(0.f, 0.f);
      float2 ps2 = make_float2(0.f, 0.f);
      uint32_t pk[16];
#pragma unroll
      for (int c = 0; c < 32; c += 8) {
        uint4 hs[2];
        if (DROP) {
          asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(hs[0].x), "=r"(hs[0].y), "=r"(hs[0].z), "=r"(hs[0].w) : "r"(hash_addr + 4 * c));
          asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(hs[1].x), "=r"(hs[1].y), "=r"(hs[1].z), "=r"(hs[1].w) : "r"(hash_addr + 4 * c + 16));
        }
        float2 last = zero_2;
#pragma unroll
        for (int h4 = 0; h4 < 4; ++h4) {
          const int cc = c + 2 * h4;
          const float2 t = __ffma2_rn(make_float2(x[cc], x[cc + 1]), sl2_2, negm_c);
          float2 pr = make_float2(fast_exp2(t.x), fast_exp2(t.y));
          ps2 = __fadd2_rn(ps2, pr);   // the row sum is taken before dropout
          if (DROP) {   // the 1/(1-p) scale is applied once to O in the epilogue
            const uint4 h = hs[h4 >> 1];
            if (!drop_keep_rc(drop_rh, (h4 & 1) ? h.z : h.x, t32)) pr.x = 0.f;
            if (!drop_keep_rc(drop_rh, (h4 & 1) ? h.w : h.y, t32)) pr.y = 0.f;
          }
          pk[cc >> 1] = pack_bf16(pr.x, pr.y);
          last = pr;
        }
        negm_c = __ffma2_rn(last, zero_2, negm_c);
      }
      l += ps2.x + ps2.y;
      if (warp == 0 && lane == 0) AF_STAMP(1, j, 3);
      if (jv > 0 && any_need) {
        // O may only be rescaled once PV_{j-1} has retired.  (Waiting only in this case is safe: the barrier can be
        // at most one phase ahead of j-1, because PV_j needs this warp's P_FULL arrival.)  The two warps of a pair
        // rescale alternate 32-column chunks of their rows.
        mbar_wait(O_READY, (j - 1) & 1);
        if (warp == 0 && lane == 0) AF_STAMP(1, j, 4);
        tc_fence_after();
        for (int c = 32 * half; c < p.dp; c += 64) {
          uint32_t o[32];
          tmem_ld32(tmem_O + lane_sel + c, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tmem_st32(tmem_O + lane_sel + c, o);
        }
        tmem_st_wait();
      }
      // P half-row -> TMEM (packed bf16, key 2c in the low half of column c): 16 of the first 32 columns of this S buffer
      tmem_st16(tmem_S + lane_sel + sb * AF_BN + 16 * half, pk);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(P_FULL);
      if (warp == 0 && lane == 0) AF_STAMP(1, j, 5);
    }

    // ---- epilogue: O / l -> bf16, heads merged; LSE (log2 domain).  Row sums of the two column halves are added.
    {
      float* xb = s_xchg + (nkv & 1) * 256;   // the buffer the last tile did not use
      xb[warp * 32 + lane] = l;
      asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
      l += xb[(warp ^ 4) * 32 + lane];
    }
    mbar_wait(O_READY, (nkv - 1) & 1);
    tc_fence_after();
    const float inv = l > 0.f ? (DROP ? p.drop_scale : 1.f) / l : 0.f;
    const bool row_ok = q < p.Sq;
    __nv_bfloat16* orow = p.out + (static_cast<long long>(b) * p.Sq + q) * p.ldo + hd * p.dp;
    for (int c = 32 * half; c < p.dp; c += 64) {
      uint32_t o[32];
      tmem_ld32(tmem_O + lane_sel + c, o);
      tmem_ld_wait();
      if (row_ok) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          const uint4 v = make_uint4(pack_bf16(__uint_as_float(o[i]) * inv, __uint_as_float(o[i + 1]) * inv),
                                     pack_bf16(__uint_as_float(o[i + 2]) * inv, __uint_as_float(o[i + 3]) * inv),
                                     pack_bf16(__uint_as_float(o[i + 4]) * inv, __uint_as_float(o[i + 5]) * inv),
                                     pack_bf16(__uint_as_float(o[i + 6]) * inv, __uint_as_float(o[i + 7]) * inv));
          *reinterpret_cast<uint4*>(orow + c + i) = v;
        }
      }
    }
    if (half == 0 && row_ok && p.lse)
      p.lse[(static_cast<long long>(b) * p.H + hd) * p.lse_stride + q] = l > 0.f ? m_used + log2f(l) : -INFINITY;
  }

  tc_fence_before();
  __syncthreads();
  if (warp == AF_SM_WARPS + 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace xf

extern "C" int xf_attn_fwd(const XfAttnFwd* a, xf_stream_t stream_) {
  using namespace xf;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!a || !a->q || !a->k || !a->v || !a->out) return fail(-1, "xf_attn_fwd: null pointer");
  if (a->dp % 32 || a->dp < 32 || a->dp > 256) return fail(-2, "xf_attn_fwd: padded head dim %d must be a multiple of 32 in [32,256]", a->dp);
  if (a->B <= 0 || a->H <= 0 || a->Sq <= 0 || a->Sk <= 0) return fail(-3, "xf_attn_fwd: bad shape");
  if ((a->ldo % 8) || (reinterpret_cast<uintptr_t>(a->out) & 15)) return fail(-4, "xf_attn_fwd: output must be 16-byte aligned with ld %% 8 == 0");
  if (a->drop_p < 0.f || a->drop_p >= 1.f) return fail(-6, "xf_attn_fwd: drop_p out of range");
  AttnFwdParams p;
  memset(&p, 0, sizeof(p));
  p.B = a->B; p.H = a->H; p.Sq = a->Sq; p.Sk = a->Sk; p.dp = a->dp;
  p.nch = a->dp / 32;
  p.q_tiles = (a->Sq + AF_BM - 1) / AF_BM;
  p.sl2 = a->scale * 1.4426950408889634f;
  p.kpm = a->key_padding_mask;
  p.kpm_start = a->kpm_start;
  p.out = reinterpret_cast<__nv_bfloat16*>(a->out);
  p.ldo = a->ldo;
  p.lse = a->lse;
  p.dbg = reinterpret_cast<long long*>(a->debug_timeline);
  p.lse_stride = a->lse_stride > 0 ? a->lse_stride : a->Sq;
  p.drop_p = a->drop_p;
  p.drop_scale = a->drop_p > 0.f ? 1.f / (1.f - a->drop_p) : 1.f;
  p.drop_seed = drop_key(a->drop_seed, a->drop_stream); p.drop_stream = a->drop_stream;
  p.drop_thresh = drop_thresh32(a->drop_p);   // compared against the full hash word

  CUtensorMap tq, tk, tv;
  int rc;
  const uint64_t cols = static_cast<uint64_t>(a->H) * a->dp;
  if ((rc = make_tmap_3d_bf16(&tq, a->q, a->B, a->Sq, cols, a->ldq, 64, AF_BM, 128))) return rc;
  if ((rc = make_tmap_chunks_bf16(&tk, a->k, a->B, a->Sk, cols, a->ldk, AF_BN, p.nch))) return rc;
  if ((rc = make_tmap_chunks_bf16(&tv, a->v, a->B, a->Sk, cols, a->ldv, AF_BN, p.nch))) return rc;

  const int ring = (AF_KST + AF_VST) * p.nch * 4096;
  const int stage = ((a->dp + 63) / 64) * 16384;   // Q staging aliases the rings
  const int smem_bytes = 1024 + 4096 + (ring > stage ? ring : stage);
  static DeviceOnce once;
  if (int rc1 = once.run([] {
        XF_CUDA(cudaFuncSetAttribute(attn_fwd_tcgen05_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        XF_CUDA(cudaFuncSetAttribute(attn_fwd_tcgen05_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        return 0;
      }))
    return rc1;
  const int grid = a->B * a->H * p.q_tiles;
  if (a->drop_p > 0.f) attn_fwd_tcgen05_kernel<true><<<grid, AF_THREADS, smem_bytes, stream>>>(tq, tk, tv, p);
  else attn_fwd_tcgen05_kernel<false><<<grid, AF_THREADS, smem_bytes, stream>>>(tq, tk, tv, p);
  g_launches.fetch_add(1);
  XF_CUDA(cudaGetLastError());
  return 0;
}
