// optim.cu -- fused AdamW on flat buffers + gradient sum-of-squares (global-norm clipping).
//
// Replaces the reference's DeepSpeed ZeRO-2 + DeepSpeedCPUAdam optimizer-offload step
// (llavamod/train/align_trainer.py:404-417, llavamod/config/dpconfig/zero2_offload.json:16-25): fp32 master
// weights and moments stay resident in HBM; one streaming pass reads grad + 3 fp32 states and writes 3 fp32
// states + the bf16 model copy (26 B/param), no PCIe traffic.  Arithmetic = torch.optim.AdamW (decoupled
// weight decay, bias correction) with HF Trainer's clip_grad_norm_(max_grad_norm) folded in as a scalar.
#include "common.cuh"

namespace {

// per-block partial sums + a ticket: the LAST block to finish adds the partials in block order, so the global gradient norm (and with it
// the clip coefficient and every weight after the step) is bit-reproducible from run to run -- a plain atomicAdd per block is not.
// One optimizer step runs at a time per process (calls on one stream are ordered), which is what the static scratch assumes.
constexpr int SUMSQ_MAX_BLOCKS = 4096;
__device__ float g_sumsq_partial[SUMSQ_MAX_BLOCKS];
__device__ unsigned int g_sumsq_ticket = 0;

__global__ void __launch_bounds__(256) sumsq_kernel(const void* __restrict__ g, int is_f32, int64_t n, float* __restrict__ out) {
  __shared__ float red[32];
  __shared__ bool s_last;
  float a = 0.f;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  if (is_f32) {
    const float* p = (const float*)g;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) a = fmaf(p[i], p[i], a);
  } else {
    const __nv_bfloat16* p = (const __nv_bfloat16*)g;
    const int64_t nv = n >> 3;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
      uint4 u = ldg_nc_v4(reinterpret_cast<const uint4*>(p) + i);
      float f[8] = {bf16lo(u.x), bf16hi(u.x), bf16lo(u.y), bf16hi(u.y), bf16lo(u.z), bf16hi(u.z), bf16lo(u.w), bf16hi(u.w)};
#pragma unroll
      for (int j = 0; j < 8; ++j) a = fmaf(f[j], f[j], a);
    }
    for (int64_t i = (nv << 3) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
      float f = __bfloat162float(p[i]);
      a = fmaf(f, f, a);
    }
  }
  a = block_sum(a, red);
  if (threadIdx.x == 0) {
    g_sumsq_partial[blockIdx.x] = a;
    __threadfence();
    s_last = (atomicAdd(&g_sumsq_ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    float t = 0.f;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) t += __ldcg(&g_sumsq_partial[i]);     // fixed assignment of partials to threads
    t = block_sum(t, red);
    if (threadIdx.x == 0) { out[0] += t; g_sumsq_ticket = 0u; }
  }
}

struct AdamArgs {
  float lr, beta1, beta2, eps, wd, bc1, bc2_sqrt, max_norm, grad_scale;
};

__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ master, float* __restrict__ m, float* __restrict__ v,
                                                   const void* __restrict__ grad, int grad_is_f32, __nv_bfloat16* __restrict__ model,
                                                   int64_t n, const float* __restrict__ gnorm_sq, AdamArgs a) {
  float coef = a.grad_scale;
  if (gnorm_sq && a.max_norm > 0.f) {
    const float total = sqrtf(gnorm_sq[0]) * a.grad_scale;           // clip_grad_norm_: coef = max_norm / (total + 1e-6), clamped to 1
    coef *= fminf(1.f, a.max_norm / (total + 1e-6f));
  }
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float g = grad_is_f32 ? ((const float*)grad)[i] : __bfloat162float(((const __nv_bfloat16*)grad)[i]);
    g *= coef;
    float p = master[i];
    p *= (1.f - a.lr * a.wd);
    float mi = m[i] * a.beta1 + (1.f - a.beta1) * g;
    float vi = v[i] * a.beta2 + (1.f - a.beta2) * g * g;
    const float denom = sqrtf(vi) / a.bc2_sqrt + a.eps;
    p -= (a.lr / a.bc1) * (mi / denom);
    master[i] = p; m[i] = mi; v[i] = vi;
    if (model) model[i] = __float2bfloat16_rn(p);
  }
}

}  // namespace

extern "C" int lmod_sumsq(const void* g, int is_f32, int64_t count, float* out_accum, void* stream) {
  LMOD_CHECK_ARG(g && out_accum && count > 0, "lmod_sumsq: bad arguments");
  LMOD_CHECK_ARG(is_f32 || ((uintptr_t)g % 16 == 0), "lmod_sumsq: bf16 buffer must be 16B aligned");
  int64_t blocks = (count / 8 + 255) / 256;
  int64_t cap = (int64_t)lmod_num_sms() * 8;
  if (cap > SUMSQ_MAX_BLOCKS) cap = SUMSQ_MAX_BLOCKS;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  sumsq_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(g, is_f32, count, out_accum);
  LMOD_LAUNCH_OK();
  return LMOD_OK;
}

extern "C" int lmod_adamw(float* master, float* m, float* v, const void* grad, int grad_is_f32, void* model_bf16, int64_t count,
                          float lr, float beta1, float beta2, float eps, float wd, int64_t step, const float* gnorm_sq,
                          float max_norm, float grad_scale, void* stream) {
  LMOD_CHECK_ARG(master && m && v && grad && count > 0 && step >= 1, "lmod_adamw: bad arguments");
  AdamArgs a;
  a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.wd = wd;
  a.bc1 = (float)(1.0 - pow((double)beta1, (double)step));
  a.bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
  a.max_norm = max_norm; a.grad_scale = grad_scale;
  int64_t blocks = (count + 255) / 256;
  int64_t cap = (int64_t)lmod_num_sms() * 16;
  if (blocks > cap) blocks = cap;
  adamw_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(master, m, v, grad, grad_is_f32, (__nv_bfloat16*)model_bf16, count, gnorm_sq, a);
  LMOD_LAUNCH_OK();
  return LMOD_OK;
}
