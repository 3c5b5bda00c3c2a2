// Fused Adam / AdamW step over a flat fp32 parameter buffer (one streaming pass, HBM-bound).
//
// Reference: the optimisers the training scripts construct -- torch.optim.Adam(lr 1e-3) / Adam(1e-4, wd 1e-6)
// (train_IEDB_wFT.py:74,97) and torch.optim.AdamW(wd 1e-6) (train_Cancer_wFT.py:98,122) -- stepped at
// procedures/train.py:28,122.  Same arithmetic as torch's `_single_tensor_adam` (amsgrad = False, maximize = False):
//     Adam : g' = g + wd p                AdamW : p <- p (1 - lr wd), g' = g
//     m <- m + (1 - b1)(g' - m)           v <- b2 v + (1 - b2) g' g'
//     p <- p - (lr / (1 - b1^t)) m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// The bias corrections are host doubles (as in torch) passed as floats.  torch runs this as ~9 foreach kernels per
// parameter list (7 x 25.3 MB of traffic each way); here p, g, m, v are read once and p, m, v written once:
// 28 B / parameter = 177 MB per step for the 6.33 M-parameter model.
#include "common.cuh"

namespace is {

__global__ void adam_tick_kernel(float* __restrict__ step) { *step += 1.0f; }

// step_dev != NULL (CUDA-graph capture): the step count lives on the device and the bias corrections are formed here
__global__ void __launch_bounds__(256)
fused_adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                  int64_t n, float lr, float beta1, float beta2, float eps, float weight_decay, int decoupled,
                  float step_size, float inv_bc2_sqrt, float grad_scale, const float* __restrict__ step_dev) {
    if (step_dev != nullptr) {
        const float t = *step_dev;
        step_size = lr / (1.0f - powf(beta1, t));
        inv_bc2_sqrt = rsqrtf(1.0f - powf(beta2, t));
    }
    const int64_t n4 = n >> 2;
    const float omb1 = 1.0f - beta1, omb2 = 1.0f - beta2;
    const float decay = decoupled ? 1.0f - lr * weight_decay : 1.0f;
    const float l2 = decoupled ? 0.0f : weight_decay;
    auto upd = [&](float& pp, float gg, float& mm, float& vv) {
        gg *= grad_scale;
        pp *= decay;
        gg = fmaf(l2, pp, gg);
        mm = fmaf(omb1, gg - mm, mm);
        vv = fmaf(omb2 * gg, gg, beta2 * vv);
        const float denom = sqrtf(vv) * inv_bc2_sqrt + eps;
        pp -= step_size * (mm / denom);
    };
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 pp = reinterpret_cast<float4*>(p)[i], mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
        const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + i);
        upd(pp.x, gg.x, mm.x, vv.x); upd(pp.y, gg.y, mm.y, vv.y); upd(pp.z, gg.z, mm.z, vv.z); upd(pp.w, gg.w, mm.w, vv.w);
        reinterpret_cast<float4*>(p)[i] = pp; reinterpret_cast<float4*>(m)[i] = mm; reinterpret_cast<float4*>(v)[i] = vv;
    }
    for (int64_t i = 4 * n4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float pp = p[i], mm = m[i], vv = v[i];
        upd(pp, g[i], mm, vv);
        p[i] = pp; m[i] = mm; v[i] = vv;
    }
}

}  // namespace is

using namespace is;

extern "C" {

// One Adam (decoupled = 0) / AdamW (decoupled = 1) step over n contiguous fp32 values.  p, g, m, v: 16-byte aligned.
// step_size = lr / (1 - beta1^t); inv_bc2_sqrt = 1 / sqrt(1 - beta2^t); grad_scale multiplies g first (1 / world_size
// after a summing all-reduce, else 1).
int is_fused_adam(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                  float weight_decay, int decoupled, float step_size, float inv_bc2_sqrt, float grad_scale, void* stream) {
    if (n < 0) return IS_ERR_ARG;
    if (n == 0) return IS_OK;
    if (((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
          reinterpret_cast<uintptr_t>(v)) & 15) != 0) return IS_ERR_ARG;
    const int sms = current_num_sms();
    int64_t blocks = ((n >> 2) + 255) / 256;
    if (blocks > (int64_t)sms * 8) blocks = (int64_t)sms * 8;
    if (blocks < 1) blocks = 1;
    fused_adam_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay,
                                                                     decoupled, step_size, inv_bc2_sqrt, grad_scale, nullptr);
    IS_LAUNCH_CHECK();
    return IS_OK;
}

// Capturable form (the whole training step replayed from a CUDA graph): `step` is a device float holding the number of
// steps taken so far; `tick` != 0 increments it first (once per optimiser step, before the first parameter run).
int is_fused_adam_capturable(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                             float eps, float weight_decay, int decoupled, float* step, int tick, float grad_scale,
                             void* stream) {
    if (n < 0 || step == nullptr) return IS_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (tick) adam_tick_kernel<<<1, 1, 0, st>>>(step);
    if (n == 0) return IS_OK;
    if (((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
          reinterpret_cast<uintptr_t>(v)) & 15) != 0) return IS_ERR_ARG;
    const int sms = current_num_sms();
    int64_t blocks = ((n >> 2) + 255) / 256;
    if (blocks > (int64_t)sms * 8) blocks = (int64_t)sms * 8;
    if (blocks < 1) blocks = 1;
    fused_adam_kernel<<<(int)blocks, 256, 0, st>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, decoupled, 0.0f, 0.0f,
                                                   grad_scale, step);
    IS_LAUNCH_CHECK();
    return IS_OK;
}

}  // extern "C"
