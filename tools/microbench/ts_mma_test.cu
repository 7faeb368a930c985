// Experiment: A operand of tcgen05.mma taken from TMEM (".ts" form), filled either by tcgen05.cp from
// shared memory or by tcgen05.st from registers (packed bf16), checked against the SS form and a CPU result.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cstring>
#include <cmath>
#include <cuda_bf16.h>
#include "ptx.cuh"
using namespace xf;

__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_cp_128x256b(uint32_t taddr, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}

__global__ void __launch_bounds__(128, 1) k(const __nv_bfloat16* A, const __nv_bfloat16* B, float* out, int pack_mode) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint32_t tmem_ptr;
  __shared__ uint64_t bar;
  uint8_t* sA = smem;           // 128 x 128 B
  uint8_t* sB = smem + 16384;   // 64 x 128 B
  const int r = threadIdx.x;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
  if (threadIdx.x < 32) { tmem_alloc(smem_u32(&tmem_ptr), 512); tmem_relinquish(); }
  // K-major SWIZZLE_128B: 16-byte segment s of row r at ((s ^ (r & 7)) << 4)
  for (int s = 0; s < 8; ++s) {
    *reinterpret_cast<uint4*>(sA + r * 128 + ((s ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(A + r * 64 + s * 8);
    if (r < 64) *reinterpret_cast<uint4*>(sB + r * 128 + ((s ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(B + r * 64 + s * 8);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_ptr;
  const uint32_t D_SS = tm, D_CP = tm + 64, D_ST = tm + 128, A_CP = tm + 256, A_ST = tm + 320;
  // (c) registers -> TMEM: row r, 64 bf16 packed two per 32-bit column
  {
    uint32_t v[32];
    const uint32_t* arow = reinterpret_cast<const uint32_t*>(A + r * 64);
    for (int c = 0; c < 32; ++c) {
      uint32_t w = arow[c];                       // low half = element 2c, high half = element 2c+1
      if (pack_mode == 1) w = (w >> 16) | (w << 16);
      v[c] = w;
    }
    tmem_st32(A_ST + (static_cast<uint32_t>((r >> 5) * 32) << 16), v);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_bf16(64, 0, 0);
    for (int kk = 0; kk < 4; ++kk) {
      const uint64_t da = make_smem_desc(smem_u32(sA) + kk * 32, 16, 1024);
      const uint64_t db = make_smem_desc(smem_u32(sB) + kk * 32, 16, 1024);
      umma_bf16(D_SS, da, db, idesc, kk != 0);
    }
    for (int kk = 0; kk < 4; ++kk) tmem_cp_128x256b(A_CP + 8 * kk, make_smem_desc(smem_u32(sA) + kk * 32, 16, 1024));
    for (int kk = 0; kk < 4; ++kk) {
      const uint64_t db = make_smem_desc(smem_u32(sB) + kk * 32, 16, 1024);
      umma_ts(D_CP, A_CP + 8 * kk, db, idesc, kk != 0);
      umma_ts(D_ST, A_ST + 8 * kk, db, idesc, kk != 0);
    }
    umma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  const uint32_t lane_sel = static_cast<uint32_t>((r >> 5) * 32) << 16;
  for (int t = 0; t < 3; ++t) {
    for (int c = 0; c < 64; c += 32) {
      uint32_t v[32];
      tmem_ld32(tm + t * 64 + lane_sel + c, v);
      tmem_ld_wait();
      for (int i = 0; i < 32; ++i) out[(t * 128 + r) * 64 + c + i] = __uint_as_float(v[i]);
    }
  }
  // dump what tcgen05.cp put into TMEM for the first k-slice (8 columns) to learn the layout
  {
    uint32_t v[32];
    tmem_ld32(A_CP + lane_sel, v);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) out[3 * 128 * 64 + r * 32 + i] = __uint_as_float(v[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem_ptr, 512); }
}

int main() {
  std::vector<__nv_bfloat16> hA(128 * 64), hB(64 * 64);
  std::vector<float> fA(128 * 64), fB(64 * 64);
  srand(1);
  for (int i = 0; i < 128 * 64; ++i) { fA[i] = (rand() % 7) - 3; hA[i] = __float2bfloat16(fA[i]); }
  for (int i = 0; i < 64 * 64; ++i) { fB[i] = (rand() % 5) - 2; hB[i] = __float2bfloat16(fB[i]); }
  std::vector<float> ref(128 * 64);
  for (int m = 0; m < 128; ++m) for (int n = 0; n < 64; ++n) { float s = 0; for (int kk = 0; kk < 64; ++kk) s += fA[m * 64 + kk] * fB[n * 64 + kk]; ref[m * 64 + n] = s; }
  __nv_bfloat16 *dA, *dB; float* dO;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dO, (3 * 128 * 64 + 128 * 32) * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int mode = 0; mode < 2; ++mode) {
    cudaMemset(dO, 0, (3 * 128 * 64 + 128 * 32) * 4);
    k<<<1, 128, 40 * 1024>>>(dA, dB, dO, mode);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<float> h(3 * 128 * 64 + 128 * 32);
    cudaMemcpy(h.data(), dO, h.size() * 4, cudaMemcpyDeviceToHost);
    const char* names[3] = {"SS", "TS via tcgen05.cp", "TS via tcgen05.st"};
    for (int t = 0; t < 3; ++t) {
      int bad = 0; double maxd = 0;
      for (int i = 0; i < 128 * 64; ++i) { double d = fabs(h[t * 128 * 64 + i] - ref[i]); if (d > 1e-3) ++bad; if (d > maxd) maxd = d; }
      printf("pack_mode %d  %-20s mismatches %5d / 8192  max|diff| %.1f\n", mode, names[t], bad, maxd);
    }
    if (mode == 0) {
      printf("tcgen05.cp TMEM dump, row 1, first 8 columns (as bf16 pairs lo,hi):");
      for (int c = 0; c < 8; ++c) { uint32_t w; memcpy(&w, &h[3 * 128 * 64 + 1 * 32 + c], 4); uint32_t lo = w << 16, hi = w & 0xffff0000u; float flo, fhi; memcpy(&flo, &lo, 4); memcpy(&fhi, &hi, 4); printf(" (%g,%g)", flo, fhi); }
      printf("\nexpected A[1][0..15]:");
      for (int c = 0; c < 16; ++c) printf(" %g", fA[64 + c]);
      printf("\n");
    }
  }
  return 0;
}
