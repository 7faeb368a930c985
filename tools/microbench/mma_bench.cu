// Micro-benchmark: cycles per tcgen05.mma (M=128, K=16, bf16) for different N, swizzle modes, operand
// major-ness and accumulator dependency patterns.  One CTA per SM (148), smem contents are arbitrary.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../transfusion_b200/csrc mma_bench.cu -o mma_bench
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "ptx.cuh"
using namespace xf;

struct Cfg { int n; int layout; int a_mn; int b_mn; int nacc; int reps; int ts; };

__global__ void __launch_bounds__(128, 1) k(const Cfg c, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint32_t tmem_ptr;
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
  if (threadIdx.x < 32) { tmem_alloc(smem_u32(&tmem_ptr), 512); tmem_relinquish(); }
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) {
    const uint32_t tm = tmem_ptr;
    const uint32_t idesc = make_idesc_bf16(c.n, c.a_mn, c.b_mn);
    const uint32_t sa = smem_u32(smem), sb = sa + 64 * 1024;
    const uint32_t lay = c.layout;  // 2 = SW128, 4 = SW64
    const uint32_t sbo = lay == 2 ? 1024 : 512;
    // 4 k-steps per "stage" like the real kernels
    uint64_t da[4], db[4];
    for (int kk = 0; kk < 4; ++kk) {
      da[kk] = make_smem_desc(sa + (c.a_mn ? kk * 2 * sbo : kk * 32 % (lay == 2 ? 128 : 64) + (kk * 32 / (lay == 2 ? 128 : 64)) * 16384), c.a_mn ? 8192 : 16, sbo, lay);
      db[kk] = make_smem_desc(sb + (c.b_mn ? kk * 2 * sbo : kk * 32 % (lay == 2 ? 128 : 64) + (kk * 32 / (lay == 2 ? 128 : 64)) * 16384), c.b_mn ? 8192 : 16, sbo, lay);
    }
    // 4 MMAs per asm block (one predicate setup), accumulators alternate between nacc buffers
    const uint32_t acc0 = tm, acc1 = tm + (c.nacc > 1 ? c.n : 0);
    long long t0 = clock64();
    if (!c.ts) {
#pragma unroll 4
    for (int r = 0; r < c.reps; ++r) {
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "setp.ne.b32 p, 1, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::f16 [%0], %2, %6, %10, p;\n\t"
          "tcgen05.mma.cta_group::1.kind::f16 [%1], %3, %7, %10, p;\n\t"
          "tcgen05.mma.cta_group::1.kind::f16 [%0], %4, %8, %10, p;\n\t"
          "tcgen05.mma.cta_group::1.kind::f16 [%1], %5, %9, %10, p;\n\t}\n"
          ::"r"(acc0), "r"(acc1), "l"(da[0]), "l"(da[1]), "l"(da[2]), "l"(da[3]), "l"(db[0]), "l"(db[1]), "l"(db[2]), "l"(db[3]),
            "r"(idesc)
          : "memory");
    }
    } else {
      const uint32_t ta = tm + 480;  // A operand: 32 columns of TMEM (4 k-slices)
#pragma unroll 4
      for (int r = 0; r < c.reps; ++r) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t.reg .b32 t1, t2, t3;\n\t"
            "setp.ne.b32 p, 1, 0;\n\t"
            "add.u32 t1, %2, 8;\n\tadd.u32 t2, %2, 16;\n\tadd.u32 t3, %2, 24;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], [%2], %3, %7, p;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%1], [t1], %4, %7, p;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], [t2], %5, %7, p;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%1], [t3], %6, %7, p;\n\t}\n"
            ::"r"(acc0), "r"(acc1), "r"(ta), "l"(db[0]), "l"(db[1]), "l"(db[2]), "l"(db[3]), "r"(idesc)
            : "memory");
      }
    }
    long long t1 = clock64();
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem_ptr, 512); }
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024);
  std::vector<Cfg> cfgs;
  for (int lay : {2, 4})
    for (int n : {32, 64, 128, 224, 256})
      for (int nacc : {1}) {
        if (nacc * n > 512) continue;
        cfgs.push_back({n, lay, 0, 0, nacc, 512, 0});
      }
  for (int n : {32, 128, 224}) cfgs.push_back({n, 2, 0, 1, 1, 512, 0});   // B MN-major
  for (int n : {128, 224}) cfgs.push_back({n, 2, 1, 1, 1, 512, 0});       // both MN-major
  for (int n : {32, 224}) cfgs.push_back({n, 4, 0, 1, 1, 512, 0});        // SW64, B MN-major
  for (int n : {32, 64, 128, 224}) cfgs.push_back({n, 2, 0, 0, 1, 512, 1});      // A from TMEM (.ts), B K-major
  for (int n : {64, 224}) cfgs.push_back({n, 2, 0, 1, 1, 512, 1});               // A from TMEM, B MN-major
  printf("%-6s %-6s %-5s %-5s %-5s %12s %12s %10s\n", "N", "swz", "a_mn", "b_mn", "nacc", "issue cyc/mma", "total cyc/mma", "math floor");
  for (auto& c : cfgs) {
    for (int it = 0; it < 2; ++it) k<<<148, 128, 205 * 1024>>>(c, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    const double nm = c.reps * 4.0;
    printf("%-6d %-6s %-5d %-5d %-5d %12.1f %12.1f %10.1f%s\n", c.n, c.layout == 2 ? "128B" : "64B", c.a_mn, c.b_mn, c.nacc, h[0] / nm, h[1] / nm, c.n / 2.0, c.ts ? "  (A from TMEM)" : "");
  }
  return 0;
}
