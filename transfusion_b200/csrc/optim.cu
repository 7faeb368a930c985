// Fused multi-tensor optimizer step for the fusion path: global-norm gradient clipping
// (Lightning gradient_clip_val / clip_grad_norm_, runner/run_experiment.py:445-446) + RAdam
// (runner/metrics_losses/radam_optim.py:30-104) + the bf16 weight copy the next forward needs, in ONE HBM pass
// over (param, grad, exp_avg, exp_avg_sq): 16 B read + 12 B (+ 2 B bf16) written per parameter, no temporaries
// (the reference runs ~12 ATen kernels per tensor: float() copies, mul_/addcmul_/add_/sqrt/add_/addcdiv_/copy_).
// HBM-bound: 128-bit loads / stores, grid = a multiple of the SM count, up to 32 tensors per launch.
#include <string.h>

#include "../../include/xfusion.h"
#include "host_common.cuh"
#include "ptx.cuh"

namespace xf {

struct OptJobs {
  int n;
  long long unit_start[XF_OPT_MAX_JOBS + 1];   // prefix sums of 4-element units per job
  XfRAdamJob job[XF_OPT_MAX_JOBS];
};

__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) red[warp] = v;
  __syncthreads();
  v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
  if (warp == 0) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  }
  return v;   // valid in thread 0
}

// out += sum over all jobs of g^2   (the fusion path's share of the global gradient norm)
__global__ void __launch_bounds__(256) grad_sqnorm_multi_kernel(const __grid_constant__ OptJobs js, float* __restrict__ out) {
  __shared__ float red[8];
  const long long total = js.unit_start[js.n];
  float acc = 0.f;
  int j = 0;
  for (long long t = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; t < total;
       t += static_cast<long long>(gridDim.x) * blockDim.x) {
    while (t >= js.unit_start[j + 1]) ++j;
    const XfRAdamJob& J = js.job[j];
    const long long e = (t - js.unit_start[j]) * 4;
    if (e + 4 <= J.n) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(J.grad + e));
      acc += g.x * g.x + g.y * g.y + g.z * g.z + g.w * g.w;
    } else {
      for (long long k = e; k < J.n; ++k) { const float g = J.grad[k]; acc += g * g; }
    }
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(out, acc);
}

struct RAdamScalars {
  float beta1, beta2, eps;
  float omb1, omb2;  // 1 - beta, evaluated in double on the host like the reference's Python floats (1.f - 0.999f is 4.7e-5 off)
  float decay;       // 1 - weight_decay * lr  (p <- p * decay before the step; 1 when weight_decay = 0)
  float step_lr;     // step_size * lr
  int mode;          // 0: moments only (N_sma < 5 and not degenerated_to_sgd), 1: adaptive step, 2: SGD-like step
  float max_norm;    // <= 0: no clipping
  const float* sqnorm;        // device scalar: sum of squares of ALL gradients that take part in the clip (or NULL)
};

__device__ __forceinline__ void radam_elem(float& p, float g, float& m, float& v, const RAdamScalars& h, float coef) {
  g *= coef;
  v = h.beta2 * v + h.omb2 * g * g;              // radam_optim.py:62
  m = h.beta1 * m + h.omb1 * g;                  // :63
  if (h.mode == 1) {                             // :92-97
    p *= h.decay;
    p -= h.step_lr * m / (sqrtf(v) + h.eps);
  } else if (h.mode == 2) {                      // :98-102
    p *= h.decay;
    p -= h.step_lr * m;
  }
}

__global__ void __launch_bounds__(256) radam_multi_kernel(const __grid_constant__ OptJobs js, const __grid_constant__ RAdamScalars h) {
  // clip coefficient: torch.nn.utils.clip_grad_norm_ -> min(1, max_norm / (total_norm + 1e-6))
  float coef = 1.f;
  if (h.max_norm > 0.f && h.sqnorm) coef = fminf(1.f, h.max_norm / (sqrtf(__ldg(h.sqnorm)) + 1e-6f));
  const long long total = js.unit_start[js.n];
  int j = 0;
  for (long long t = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; t < total;
       t += static_cast<long long>(gridDim.x) * blockDim.x) {
    while (t >= js.unit_start[j + 1]) ++j;
    const XfRAdamJob& J = js.job[j];
    const long long e = (t - js.unit_start[j]) * 4;
    if (e + 4 <= J.n) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(J.grad + e));
      float4 p = *reinterpret_cast<const float4*>(J.param + e);
      float4 m = *reinterpret_cast<const float4*>(J.exp_avg + e);
      float4 v = *reinterpret_cast<const float4*>(J.exp_avg_sq + e);
      radam_elem(p.x, g.x, m.x, v.x, h, coef); radam_elem(p.y, g.y, m.y, v.y, h, coef);
      radam_elem(p.z, g.z, m.z, v.z, h, coef); radam_elem(p.w, g.w, m.w, v.w, h, coef);
      *reinterpret_cast<float4*>(J.exp_avg + e) = m;
      *reinterpret_cast<float4*>(J.exp_avg_sq + e) = v;
      if (h.mode != 0) {
        *reinterpret_cast<float4*>(J.param + e) = p;
        if (J.param_bf16) *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(J.param_bf16) + e) = make_uint2(pack_bf16(p.x, p.y), pack_bf16(p.z, p.w));
      }
    } else {
      for (long long k = e; k < J.n; ++k) {
        float p = J.param[k], m = J.exp_avg[k], v = J.exp_avg_sq[k];
        radam_elem(p, J.grad[k], m, v, h, coef);
        J.exp_avg[k] = m; J.exp_avg_sq[k] = v;
        if (h.mode != 0) {
          J.param[k] = p;
          if (J.param_bf16) reinterpret_cast<__nv_bfloat16*>(J.param_bf16)[k] = __float2bfloat16(p);
        }
      }
    }
  }
}

static int fill_jobs(OptJobs& js, const XfRAdamJob* jobs, int n_jobs, bool need_state, const char* who) {
  if (!jobs || n_jobs <= 0 || n_jobs > XF_OPT_MAX_JOBS) return fail(-1, "%s: need 1..%d jobs", who, XF_OPT_MAX_JOBS);
  memset(&js, 0, sizeof(js));
  js.n = n_jobs;
  for (int i = 0; i < n_jobs; ++i) {
    const XfRAdamJob& J = jobs[i];
    if (!J.grad || J.n < 0) return fail(-2, "%s: job %d: null gradient or negative size", who, i);
    if (need_state && (!J.param || !J.exp_avg || !J.exp_avg_sq)) return fail(-2, "%s: job %d: null parameter / state pointer", who, i);
    uintptr_t al = reinterpret_cast<uintptr_t>(J.grad);
    if (need_state) al |= reinterpret_cast<uintptr_t>(J.param) | reinterpret_cast<uintptr_t>(J.exp_avg) | reinterpret_cast<uintptr_t>(J.exp_avg_sq);
    if (al & 15) return fail(-3, "%s: job %d: pointers must be 16-byte aligned", who, i);
    if (need_state && J.param_bf16 && (reinterpret_cast<uintptr_t>(J.param_bf16) & 7)) return fail(-3, "%s: job %d: bf16 copy must be 8-byte aligned", who, i);
    js.job[i] = J;
    js.unit_start[i + 1] = js.unit_start[i] + (J.n + 3) / 4;
  }
  return 0;
}

}  // namespace xf

using namespace xf;

extern "C" int xf_grad_sqnorm(const XfRAdamJob* jobs, int n_jobs, float* out, xf_stream_t s) {
  if (!out) return fail(-1, "xf_grad_sqnorm: null output");
  OptJobs js;
  if (int rc = fill_jobs(js, jobs, n_jobs, false, "xf_grad_sqnorm")) return rc;
  const long long total = js.unit_start[js.n];
  if (total == 0) return 0;
  long long ctas = (total + 255) / 256;
  const long long cap = 8ll * sm_count();
  if (ctas > cap) ctas = cap;
  grad_sqnorm_multi_kernel<<<static_cast<int>(ctas), 256, 0, reinterpret_cast<cudaStream_t>(s)>>>(js, out);
  g_launches.fetch_add(1);
  XF_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int xf_radam_step(const XfRAdamJob* jobs, int n_jobs, const XfRAdam* a, xf_stream_t s) {
  if (!a) return fail(-1, "xf_radam_step: null hyper-parameters");
  if (!(a->beta1_d >= 0. && a->beta1_d < 1. && a->beta2_d >= 0. && a->beta2_d < 1.) || a->lr_d < 0. || a->eps < 0.f)
    return fail(-4, "xf_radam_step: invalid hyper-parameters");
  if (a->step < 1) return fail(-4, "xf_radam_step: step counts from 1");
  OptJobs js;
  if (int rc = fill_jobs(js, jobs, n_jobs, true, "xf_radam_step")) return rc;
  const long long total = js.unit_start[js.n];
  if (total == 0) return 0;
  // radam_optim.py:66-87 in double precision on the host (the reference evaluates it with Python floats)
  const double b1 = a->beta1_d, b2 = a->beta2_d, t = static_cast<double>(a->step);
  const double beta2_t = pow(b2, t);
  const double n_max = 2.0 / (1.0 - b2) - 1.0;
  const double n_sma = n_max - 2.0 * t * beta2_t / (1.0 - beta2_t);
  RAdamScalars h;
  memset(&h, 0, sizeof(h));
  h.beta1 = static_cast<float>(a->beta1_d); h.beta2 = static_cast<float>(a->beta2_d); h.eps = a->eps;
  h.omb1 = static_cast<float>(1.0 - a->beta1_d); h.omb2 = static_cast<float>(1.0 - a->beta2_d);
  h.decay = a->weight_decay_d != 0. ? static_cast<float>(1.0 - a->weight_decay_d * a->lr_d) : 1.f;
  double step_size = -1.0;
  if (n_sma >= 5.0) {
    step_size = sqrt((1.0 - beta2_t) * (n_sma - 4.0) / (n_max - 4.0) * (n_sma - 2.0) / n_sma * n_max / (n_max - 2.0)) / (1.0 - pow(b1, t));
    h.mode = 1;
  } else if (a->degenerated_to_sgd) {
    step_size = 1.0 / (1.0 - pow(b1, t));
    h.mode = 2;
  } else {
    h.mode = 0;
  }
  h.step_lr = static_cast<float>(step_size * a->lr_d);
  h.max_norm = a->max_grad_norm;
  h.sqnorm = a->grad_sqnorm;
  long long ctas = (total + 255) / 256;
  const long long cap = 8ll * sm_count();
  if (ctas > cap) ctas = cap;
  radam_multi_kernel<<<static_cast<int>(ctas), 256, 0, reinterpret_cast<cudaStream_t>(s)>>>(js, h);
  g_launches.fetch_add(1);
  XF_CUDA(cudaGetLastError());
  return 0;
}
