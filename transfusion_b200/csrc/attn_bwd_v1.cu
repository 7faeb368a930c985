// xf_attn_bwd: fused attention backward for sm_100a (recomputes S from Q, K and the saved LSE; no
// S x S tensor).  autograd backward of torch18_adapters.py:789-798 (+ head split/merge :544-555,607).
//
// Two launches of one templated kernel (5 GEMMs of the textbook backward become 3 + 4 because S and
// dP are recomputed in both; the accumulators of dQ, dK and dV (3 x dp fp32 columns) do not fit the
// 512 TMEM columns together with the score tiles when dp = 224):
//
//   DKV = false ("dQ pass", query-stationary):  resident R1 = Q, R2 = dO tiles [128 x dp];
//        stream T1 = K_j, T2 = V_j (32 keys):  C1 = Q K_j^T, C2 = dO V_j^T,  dS = P o (C2 - delta) * scale,
//        dQ += dS K_j.
//   DKV = true  ("dK/dV pass", key-stationary): resident R1 = K, R2 = V tiles [128 x dp];
//        stream T1 = Q_i, T2 = dO_i (32 queries): C1 = K Q_i^T (= S^T), C2 = V dO_i^T (= dP^T),
//        dV += P^T dO_i,  dK += dS^T Q_i.
//
// In both: C1/C2 are tcgen05 MMAs (M=128, N=32, K=dp) into TMEM, the element-wise stage runs one
// thread per resident row (TMEM lane) out of registers, writes the bf16 tiles E1 (= P^T, DKV only) and
// E2 (= dS or dS^T) into SWIZZLE_64B shared memory as K-major A operands, and the accumulating MMAs
// (M=128, N=dp, K=32) read the streamed tile a second time as an MN-major B operand.  TMA feeds a
// 3-stage ring of streamed tiles.  All tiles use 32-column (64-byte) chunks with SWIZZLE_64B so a
// 224-wide head needs exactly 7 chunks (no padding to 256), which is what lets the 3-stage ring fit.
//
// CTA = 192 threads: warps 0-3 element-wise + epilogue, warp 4 TMA producer, warp 5 MMA issuer.
#include <string.h>

#include "../../include/xfusion.h"
#include "host_common.cuh"
#include "ptx.cuh"

namespace xf {

constexpr int AB_BM = 128;      // resident rows per CTA
constexpr int AB_BN = 32;       // streamed rows per iteration
constexpr int AB_STAGES = 3;
constexpr int AB_THREADS = 192;
constexpr uint32_t SW64 = 4;    // UMMA layout code for SWIZZLE_64B

struct AttnBwdParams {
  int B, H, Sq, Sk, dp, nch, r_tiles, n_stream, ncbuf;
  float sl2, scale;
  const uint8_t* kpm;      // [B, Sk] 1 = ignore, or null
  int kpm_start;           // keys < kpm_start are never masked
  const float* lse;        // [B, H, stat_stride] log2 domain
  const float* delta;      // [B, H, stat_stride]
  int stat_stride;
  __nv_bfloat16* out2; long long ld2;  // DQ: dq ; DKV: dk
  __nv_bfloat16* out1; long long ld1;  // DKV: dv
  float drop_p, drop_scale;
  uint32_t drop_seed, drop_stream, drop_thresh;
  long long* dbg;  // dev aid: per-iteration clock64() stamps of CTA 0 (null in production)
};

#define AB_STAMP(role, i, ev) do { if (p.dbg && blockIdx.x == 0 && (i) < 64) p.dbg[((role) * 64 + (i)) * 8 + (ev)] = clock64(); } while (0)

template <bool DKV, bool DROP>
__global__ void __launch_bounds__(AB_THREADS, 1)
attn_bwd_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_r1, const __grid_constant__ CUtensorMap tmap_r2,
                        const __grid_constant__ CUtensorMap tmap_t1, const __grid_constant__ CUtensorMap tmap_t2,
                        const __grid_constant__ AttnBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + 256);
  const uint32_t r_bytes = p.nch * 8192u;   // [128 rows x 64 B] per chunk
  const uint32_t t_bytes = p.nch * 2048u;   // [32 rows x 64 B] per chunk
  uint8_t* sR1 = smem + 1024;
  uint8_t* sR2 = sR1 + r_bytes;
  uint8_t* sT = sR2 + r_bytes;              // ring: stage s -> T1 at sT + s*2*t_bytes, T2 right after
  uint8_t* sE1 = sT + AB_STAGES * 2 * t_bytes;
  uint8_t* sE2 = sE1 + 8192;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_u32(bars);
  const uint32_t R_FULL = bar0;
  auto T_FULL = [&](int s) { return bar0 + 8u * (1 + s); };
  auto T_EMPTY = [&](int s) { return bar0 + 8u * (4 + s); };
  auto C_FULL = [&](int s) { return bar0 + 8u * (7 + s); };
  auto C_EMPTY = [&](int s) { return bar0 + 8u * (9 + s); };
  const uint32_t E_FULL = bar0 + 8u * 11;
  const uint32_t E_EMPTY = bar0 + 8u * 12;
  const uint32_t ACC_DONE = bar0 + 8u * 13;

  int bid = blockIdx.x;
  const int rt = bid % p.r_tiles; bid /= p.r_tiles;
  const int hd = bid % p.H;
  const int b = bid / p.H;
  const int r0 = rt * AB_BM;
  const int n = p.n_stream;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_r1); tma_prefetch_desc(&tmap_r2);
    tma_prefetch_desc(&tmap_t1); tma_prefetch_desc(&tmap_t2);
    mbar_init(R_FULL, 1);
    for (int s = 0; s < AB_STAGES; ++s) { mbar_init(T_FULL(s), 1); mbar_init(T_EMPTY(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(C_FULL(s), 1); mbar_init(C_EMPTY(s), 4); }
    mbar_init(E_FULL, 4);
    mbar_init(E_EMPTY, 1);
    mbar_init(ACC_DONE, 1);
    fence_mbar_init();
  }
  if (warp == 5) {
    tmem_alloc(smem_u32(tmem_ptr_smem), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t tmem_acc2 = tmem_base;
  const uint32_t tmem_acc1 = tmem_base + p.dp;                       // DKV only
  const uint32_t tmem_C = tmem_base + (DKV ? 2 : 1) * p.dp;          // ncbuf x (C1: 32 | C2: 32)

  // warps 0-3: element-wise stage (TMEM lane quadrant = warp id); warp 4: TMA producer; warp 5: MMA issuer
  // (highest ids: the scheduler favours them over the ALU-heavy element-wise warps)
  if (warp == 4) {
    if (lane == 0) {
      const int col0 = hd * p.dp;
      mbar_expect_tx(R_FULL, 2 * r_bytes);
      for (int c = 0; c < p.nch; ++c) {
        tma_load_3d(smem_u32(sR1 + c * 8192), &tmap_r1, R_FULL, col0 + 32 * c, r0, b);
        tma_load_3d(smem_u32(sR2 + c * 8192), &tmap_r2, R_FULL, col0 + 32 * c, r0, b);
      }
      for (int i = 0; i < n; ++i) {
        const int st = i % AB_STAGES;
        mbar_wait(T_EMPTY(st), ((i / AB_STAGES) & 1) ^ 1);
        mbar_expect_tx(T_FULL(st), 2 * t_bytes);
        const uint32_t t1 = smem_u32(sT + st * 2 * t_bytes), t2 = t1 + t_bytes;
        for (int c = 0; c < p.nch; ++c) {
          tma_load_3d(t1 + c * 2048, &tmap_t1, T_FULL(st), col0 + 32 * c, i * AB_BN, b);
          tma_load_3d(t2 + c * 2048, &tmap_t2, T_FULL(st), col0 + 32 * c, i * AB_BN, b);
        }
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      // The score MMAs are N = 32: the tensor pipe retires one every ~46 cycles (measured), so this thread
      // must issue them with a handful of instructions each: descriptors are built once (only the
      // start-address word varies) and each 32-column chunk (2 k-steps) is one asm block.
      const uint32_t idesc_c = make_idesc_bf16(AB_BN, 0, 0);
      const uint32_t idesc_acc = make_idesc_bf16(p.dp, 0, 1);
      const int nch = p.nch;
      const uint64_t dk = make_smem_desc(0, 16, 512, SW64);      // K-major template (start = 0)
      const uint64_t dmn = make_smem_desc(0, 2048, 512, SW64);   // MN-major template
      const uint32_t hi_k = desc_hi(dk), hi_mn = desc_hi(dmn), lo_k = desc_lo(dk), lo_mn = desc_lo(dmn);
      const uint32_t r1lo = lo_k + (smem_u32(sR1) >> 4), r2lo = lo_k + (smem_u32(sR2) >> 4);
      const uint32_t e1lo = lo_k + (smem_u32(sE1) >> 4), e2lo = lo_k + (smem_u32(sE2) >> 4);
      const uint32_t t_base = smem_u32(sT) >> 4, t_lo = t_bytes >> 4;
      auto do_acc = [&](int i) {
        const int st = i % AB_STAGES;
        const uint32_t t1s = t_base + st * 2 * t_lo, t2s = t1s + t_lo;
        mbar_wait(E_FULL, i & 1);
        AB_STAMP(0, i + 1, 4);
        tc_fence_after();
        umma_k2(tmem_acc2, hi_k, e2lo, 2, hi_mn, lo_mn + t1s, 64, idesc_acc, i != 0);
        if (DKV) umma_k2(tmem_acc1, hi_k, e1lo, 2, hi_mn, lo_mn + t2s, 64, idesc_acc, i != 0);
        umma_commit(T_EMPTY(st));
        umma_commit(E_EMPTY);
      };
      mbar_wait(R_FULL, 0);
      for (int i = 0; i < n; ++i) {
        const int st = i % AB_STAGES, cb = i % p.ncbuf;
        AB_STAMP(0, i, 0);
        mbar_wait(T_FULL(st), (i / AB_STAGES) & 1);
        AB_STAMP(0, i, 1);
        mbar_wait(C_EMPTY(cb), ((i / p.ncbuf) & 1) ^ 1);
        AB_STAMP(0, i, 2);
        tc_fence_after();
        const uint32_t t1s = lo_k + t_base + st * 2 * t_lo, t2s = t1s + t_lo;
        const uint32_t c1 = tmem_C + cb * 64, c2 = c1 + 32;
        for (int ch = 0; ch < nch; ++ch) {   // resident chunks are 8192 B apart, streamed chunks 2048 B
          umma_k2(c1, hi_k, r1lo + ch * 512, 2, hi_k, t1s + ch * 128, 2, idesc_c, ch != 0);
          umma_k2(c2, hi_k, r2lo + ch * 512, 2, hi_k, t2s + ch * 128, 2, idesc_c, ch != 0);
        }
        umma_commit(C_FULL(cb));
        AB_STAMP(0, i, 3);
        if (i >= 1) do_acc(i - 1);
        AB_STAMP(0, i, 5);
      }
      do_acc(n - 1);
      umma_commit(ACC_DONE);
    }
  } else {
    // ===================== element-wise stage + epilogue =====================
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    const int row_g = r0 + r;  // query index (DQ) or key index (DKV) within the batch element
    const uint32_t lane_sel = static_cast<uint32_t>(quad * 32) << 16;
    const long long stat_base = (static_cast<long long>(b) * p.H + hd) * p.stat_stride;
    float lse_row = 0.f, delta_row = 0.f;
    bool row_valid = true;
    if (!DKV) {
      if (row_g < p.Sq) { lse_row = p.lse[stat_base + row_g]; delta_row = p.delta[stat_base + row_g]; }
    } else {
      row_valid = row_g < p.Sk && !(p.kpm && p.kpm[static_cast<long long>(b) * p.Sk + row_g] != 0);
    }
    const uint64_t bh = static_cast<uint64_t>(b * p.H + hd);
    const uint32_t swz = (static_cast<uint32_t>(r) >> 1) & 3u;
    uint8_t* e1row = sE1 + r * 64;
    uint8_t* e2row = sE2 + r * 64;
    // dropout: decision(row = (b,h,q), col = key).  dQ pass: this thread's row hash is constant, one hash per key
    // pair.  dK/dV pass: this thread's key is constant (half-word select + pair term), the 32 row hashes of the
    // streamed queries are computed by the 32 lanes once per iteration and broadcast with shuffles.
    const uint32_t rh_row = DROP && !DKV ? drop_rowhash(p.drop_seed, bh * p.Sq + row_g) : 0u;
    const uint32_t colterm = (static_cast<uint32_t>(row_g) >> 1) * 0x9E3779B9U;
    const uint32_t colshift = (row_g & 1) * 16;
    const float sc = p.scale;

    for (int i = 0; i < n; ++i) {
      const int cb = i % p.ncbuf;
      const int t0 = i * AB_BN;
      // global-memory operands of this iteration are requested BEFORE waiting on the MMA so their
      // latency hides behind it: key-padding bits (dQ pass) / per-query LSE and delta (dK/dV pass)
      uint32_t badbits = 0, rh_lane = 0;
      float ls[32], ds[32];
      if (!DKV) {
        const int key = t0 + lane;
        bool bad = key >= p.Sk;
        if (!bad && p.kpm && t0 + AB_BN > p.kpm_start) bad = p.kpm[static_cast<long long>(b) * p.Sk + key] != 0;
        badbits = __ballot_sync(0xffffffffu, bad);
      } else {
        const float4* lp = reinterpret_cast<const float4*>(p.lse + stat_base + t0);
        const float4* dl = reinterpret_cast<const float4*>(p.delta + stat_base + t0);
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4) {
          const float4 L = __ldg(lp + c4);
          const float4 Dl = __ldg(dl + c4);
          ls[4 * c4] = L.x; ls[4 * c4 + 1] = L.y; ls[4 * c4 + 2] = L.z; ls[4 * c4 + 3] = L.w;
          ds[4 * c4] = Dl.x; ds[4 * c4 + 1] = Dl.y; ds[4 * c4 + 2] = Dl.z; ds[4 * c4 + 3] = Dl.w;
        }
        if (DROP) rh_lane = drop_rowhash(p.drop_seed, bh * p.Sq + (t0 + lane));
        const int qvalid = p.Sq - t0;   // columns >= qvalid are beyond the sequence
        badbits = (!row_valid) ? 0xffffffffu : (qvalid >= 32 ? 0u : (0xffffffffu << (qvalid < 0 ? 0 : qvalid)));
      }
      if (warp == 0 && lane == 0) AB_STAMP(1, i, 0);
      mbar_wait(C_FULL(cb), (i / p.ncbuf) & 1);
      if (warp == 0 && lane == 0) AB_STAMP(1, i, 1);
      tc_fence_after();
      uint32_t c1[32], c2[32];
      tmem_ld32(tmem_C + lane_sel + cb * 64, c1);
      tmem_ld32(tmem_C + lane_sel + cb * 64 + 32, c2);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(C_EMPTY(cb));
      if (warp == 0 && lane == 0) AB_STAMP(1, i, 2);

      float e1[32], e2[32];
      if (!DKV) {
        // columns = keys t0 + c; row statistics are scalars
#pragma unroll
        for (int c = 0; c < 32; c += 2) {
          const float pr0 = ((badbits >> c) & 1u) ? 0.f : fast_exp2(fmaf(__uint_as_float(c1[c]), p.sl2, -lse_row));
          const float pr1 = ((badbits >> (c + 1)) & 1u) ? 0.f : fast_exp2(fmaf(__uint_as_float(c1[c + 1]), p.sl2, -lse_row));
          float dp0 = __uint_as_float(c2[c]), dp1 = __uint_as_float(c2[c + 1]);
          if (DROP) {
            const uint32_t hsh = drop_pairhash(rh_row, static_cast<uint32_t>(t0 + c) >> 1);
            dp0 = drop_keep_lo(hsh, p.drop_thresh) ? dp0 * p.drop_scale : 0.f;
            dp1 = drop_keep_hi(hsh, p.drop_thresh) ? dp1 * p.drop_scale : 0.f;
          }
          e2[c] = pr0 * (dp0 - delta_row) * sc;
          e2[c + 1] = pr1 * (dp1 - delta_row) * sc;
        }
      } else {
        // columns = queries t0 + c; per-column statistics, this thread's key is fixed
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const float pr = ((badbits >> c) & 1u) ? 0.f : fast_exp2(fmaf(__uint_as_float(c1[c]), p.sl2, -ls[c]));
          float dpv = __uint_as_float(c2[c]);
          float pd = pr;
          if (DROP) {
            const uint32_t hsh = mix32(__shfl_sync(0xffffffffu, rh_lane, c) + colterm);
            const bool keep = ((hsh >> colshift) & 0xFFFFu) >= p.drop_thresh;
            dpv = keep ? dpv * p.drop_scale : 0.f;
            pd = keep ? pr * p.drop_scale : 0.f;
          }
          e1[c] = pd;
          e2[c] = pr * (dpv - ds[c]) * sc;
        }
      }
      if (warp == 0 && lane == 0) AB_STAMP(1, i, 3);
      if (i > 0) mbar_wait(E_EMPTY, (i - 1) & 1);
      if (warp == 0 && lane == 0) AB_STAMP(1, i, 4);
#pragma unroll
      for (int sgm = 0; sgm < 4; ++sgm) {
        const uint32_t off = (static_cast<uint32_t>(sgm) ^ swz) << 4;
        *reinterpret_cast<uint4*>(e2row + off) =
            make_uint4(pack_bf16(e2[8 * sgm], e2[8 * sgm + 1]), pack_bf16(e2[8 * sgm + 2], e2[8 * sgm + 3]),
                       pack_bf16(e2[8 * sgm + 4], e2[8 * sgm + 5]), pack_bf16(e2[8 * sgm + 6], e2[8 * sgm + 7]));
        if (DKV)
          *reinterpret_cast<uint4*>(e1row + off) =
              make_uint4(pack_bf16(e1[8 * sgm], e1[8 * sgm + 1]), pack_bf16(e1[8 * sgm + 2], e1[8 * sgm + 3]),
                         pack_bf16(e1[8 * sgm + 4], e1[8 * sgm + 5]), pack_bf16(e1[8 * sgm + 6], e1[8 * sgm + 7]));
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(E_FULL);
      if (warp == 0 && lane == 0) AB_STAMP(1, i, 5);
    }

    // ---- epilogue: accumulators -> bf16 -> global (token-major, heads merged)
    mbar_wait(ACC_DONE, 0);
    tc_fence_after();
    const int limit = DKV ? p.Sk : p.Sq;
    const bool ok = row_g < limit;
    const long long tok = static_cast<long long>(b) * limit + row_g;
#pragma unroll 1
    for (int which = 0; which < (DKV ? 2 : 1); ++which) {
      const uint32_t tacc = which == 0 ? tmem_acc2 : tmem_acc1;
      __nv_bfloat16* orow = which == 0 ? p.out2 + tok * p.ld2 + hd * p.dp : p.out1 + tok * p.ld1 + hd * p.dp;
      for (int c = 0; c < p.dp; c += 32) {
        uint32_t o[32];
        tmem_ld32(tacc + lane_sel + c, o);
        tmem_ld_wait();
        if (ok) {
#pragma unroll
          for (int k = 0; k < 32; k += 8)
            *reinterpret_cast<uint4*>(orow + c + k) =
                make_uint4(pack_bf16(__uint_as_float(o[k]), __uint_as_float(o[k + 1])), pack_bf16(__uint_as_float(o[k + 2]), __uint_as_float(o[k + 3])),
                           pack_bf16(__uint_as_float(o[k + 4]), __uint_as_float(o[k + 5])), pack_bf16(__uint_as_float(o[k + 6]), __uint_as_float(o[k + 7])));
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace xf

extern "C" int xf_attn_bwd_v1(const XfAttnBwd* a, xf_stream_t stream_) {
  using namespace xf;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!a || !a->q || !a->k || !a->v || !a->d_out || !a->lse || !a->delta || !a->dq || !a->dk || !a->dv)
    return fail(-1, "xf_attn_bwd: null pointer");
  if (a->dp % 32 || a->dp < 32 || a->dp > 224) return fail(-2, "xf_attn_bwd: padded head dim %d must be a multiple of 32 in [32,224]", a->dp);
  if (a->B <= 0 || a->H <= 0 || a->Sq <= 0 || a->Sk <= 0) return fail(-3, "xf_attn_bwd: bad shape");
  if (a->stat_stride % 32 || a->stat_stride < ((a->Sq + 31) / 32) * 32) return fail(-4, "xf_attn_bwd: stat_stride must be a multiple of 32 >= Sq rounded up to 32");
  if ((a->lddq % 8) || (a->lddk % 8) || (a->lddv % 8)) return fail(-5, "xf_attn_bwd: gradient leading dims must be multiples of 8");
  if (a->drop_p < 0.f || a->drop_p >= 1.f) return fail(-6, "xf_attn_bwd: drop_p out of range");

  AttnBwdParams p;
  memset(&p, 0, sizeof(p));
  p.B = a->B; p.H = a->H; p.Sq = a->Sq; p.Sk = a->Sk; p.dp = a->dp; p.nch = a->dp / 32;
  p.sl2 = a->scale * 1.4426950408889634f;
  p.scale = a->scale;
  p.kpm = a->key_padding_mask;
  p.kpm_start = a->kpm_start;
  p.dbg = reinterpret_cast<long long*>(a->debug_timeline);
  p.lse = a->lse; p.delta = a->delta; p.stat_stride = a->stat_stride;
  p.drop_p = a->drop_p;
  p.drop_scale = a->drop_p > 0.f ? 1.f / (1.f - a->drop_p) : 1.f;
  p.drop_seed = drop_key(a->drop_seed, a->drop_stream); p.drop_stream = a->drop_stream;
  p.drop_thresh = drop_thresh16(a->drop_p);

  const uint64_t cols = static_cast<uint64_t>(a->H) * a->dp;
  CUtensorMap q128, do128, k32, v32, k128, v128, q32, do32;
  int rc;
  if ((rc = make_tmap_3d_bf16(&q128, a->q, a->B, a->Sq, cols, a->ldq, 32, AB_BM, 64))) return rc;
  if ((rc = make_tmap_3d_bf16(&do128, a->d_out, a->B, a->Sq, cols, a->lddo, 32, AB_BM, 64))) return rc;
  if ((rc = make_tmap_3d_bf16(&k32, a->k, a->B, a->Sk, cols, a->ldk, 32, AB_BN, 64))) return rc;
  if ((rc = make_tmap_3d_bf16(&v32, a->v, a->B, a->Sk, cols, a->ldv, 32, AB_BN, 64))) return rc;
  if ((rc = make_tmap_3d_bf16(&k128, a->k, a->B, a->Sk, cols, a->ldk, 32, AB_BM, 64))) return rc;
  if ((rc = make_tmap_3d_bf16(&v128, a->v, a->B, a->Sk, cols, a->ldv, 32, AB_BM, 64))) return rc;
  if ((rc = make_tmap_3d_bf16(&q32, a->q, a->B, a->Sq, cols, a->ldq, 32, AB_BN, 64))) return rc;
  if ((rc = make_tmap_3d_bf16(&do32, a->d_out, a->B, a->Sq, cols, a->lddo, 32, AB_BN, 64))) return rc;

  const int smem_bytes = 1024 + 1024 + 2 * p.nch * 8192 + AB_STAGES * 2 * p.nch * 2048 + 2 * 8192;
  static bool attr_set = false;
  if (!attr_set) {
    XF_CUDA(cudaFuncSetAttribute(attn_bwd_tcgen05_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    XF_CUDA(cudaFuncSetAttribute(attn_bwd_tcgen05_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    XF_CUDA(cudaFuncSetAttribute(attn_bwd_tcgen05_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    XF_CUDA(cudaFuncSetAttribute(attn_bwd_tcgen05_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  // dQ pass
  {
    AttnBwdParams pq = p;
    pq.r_tiles = (a->Sq + AB_BM - 1) / AB_BM;
    pq.n_stream = (a->Sk + AB_BN - 1) / AB_BN;
    pq.ncbuf = (512 - a->dp) / 64 >= 2 ? 2 : 1;
    pq.out2 = reinterpret_cast<__nv_bfloat16*>(a->dq); pq.ld2 = a->lddq;
    if (a->drop_p > 0.f) attn_bwd_tcgen05_kernel<false, true><<<a->B * a->H * pq.r_tiles, AB_THREADS, smem_bytes, stream>>>(q128, do128, k32, v32, pq);
    else attn_bwd_tcgen05_kernel<false, false><<<a->B * a->H * pq.r_tiles, AB_THREADS, smem_bytes, stream>>>(q128, do128, k32, v32, pq);
    g_launches.fetch_add(1);
    XF_CUDA(cudaGetLastError());
  }
  // dK / dV pass
  {
    AttnBwdParams pk = p;
    pk.r_tiles = (a->Sk + AB_BM - 1) / AB_BM;
    pk.n_stream = (a->Sq + AB_BN - 1) / AB_BN;
    pk.ncbuf = (512 - 2 * a->dp) / 64 >= 2 ? 2 : 1;
    pk.out2 = reinterpret_cast<__nv_bfloat16*>(a->dk); pk.ld2 = a->lddk;
    pk.out1 = reinterpret_cast<__nv_bfloat16*>(a->dv); pk.ld1 = a->lddv;
    if (pk.dbg) pk.dbg += 2 * 64 * 8;
    if (a->drop_p > 0.f) attn_bwd_tcgen05_kernel<true, true><<<a->B * a->H * pk.r_tiles, AB_THREADS, smem_bytes, stream>>>(k128, v128, q32, do32, pk);
    else attn_bwd_tcgen05_kernel<true, false><<<a->B * a->H * pk.r_tiles, AB_THREADS, smem_bytes, stream>>>(k128, v128, q32, do32, pk);
    g_launches.fetch_add(1);
    XF_CUDA(cudaGetLastError());
  }
  return 0;
}
