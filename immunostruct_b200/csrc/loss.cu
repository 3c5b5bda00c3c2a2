// Fused training losses: prediction term + VAE reconstruction MSE + KL divergence in one pass.
//
// Reference: ``Losses.regression_loss`` / ``Losses.BCE_loss`` (immunostruct/utils/loss.py:13-31):
//   BCE_loss        = 5.0 * BCEWithLogits(out, y, pos_weight) + 0.1 * MSE(recon, seq) + 0.1 * KLD
//   regression_loss = 2.0 * MSE(out, y)                       + 0.5 * MSE(recon, seq) + 0.5 * KLD
//   KLD = -0.5 * mean(1 + logvar - mu^2 - exp(logvar));   without --sequence-loss only the first
//   term (unweighted).  The weights are passed in by the host so the same kernel serves all four.
// The reconstruction term streams recon and seq once (memory bound: 8 bytes per element); all sums
// are two-stage with a fixed order (no atomics).
#include "common.cuh"

namespace is {

__device__ __forceinline__ float block_sum(float v, float* red) {
    v = warp_sum(v);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float s = 0.0f;
    if (threadIdx.x == 0)
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
    return s;   // valid in thread 0
}

__global__ void __launch_bounds__(IS_THREADS)
mse_partial_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n, float* __restrict__ partial) {
    __shared__ float red[8];
    float s = 0.0f;
    const int64_t n4 = n >> 2;
    const float4* a4 = reinterpret_cast<const float4*>(a);
    const float4* b4 = reinterpret_cast<const float4*>(b);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 x = __ldg(a4 + i), y = __ldg(b4 + i);
        const float d0 = x.x - y.x, d1 = x.y - y.y, d2 = x.z - y.z, d3 = x.w - y.w;
        s += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
    }
    if (blockIdx.x == 0)
        for (int64_t i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) { const float d = a[i] - b[i]; s += d * d; }
    s = block_sum(s, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

// out[0] = total, out[1] = prediction term (unweighted mean), out[2] = recon MSE, out[3] = KLD
__global__ void __launch_bounds__(IS_THREADS)
loss_final_kernel(const float* __restrict__ partial, int nparts, int64_t n_recon,
                  const float* __restrict__ mu, const float* __restrict__ logvar, int64_t n_lat,
                  const float* __restrict__ logits, const float* __restrict__ y, int64_t B, int mode, float pos_weight,
                  float w_pred, float w_mse, float w_kld, float* __restrict__ out) {
    __shared__ float red[8];
    float s_pred = 0.0f, s_kld = 0.0f, s_mse = 0.0f;
    for (int64_t i = threadIdx.x; i < B; i += blockDim.x) {
        const float x = logits[i], t = y[i];
        if (mode == 0) {   // binary_cross_entropy_with_logits with pos_weight
            const float lw = 1.0f + (pos_weight - 1.0f) * t;
            s_pred += (1.0f - t) * x + lw * (log1pf(expf(-fabsf(x))) + fmaxf(-x, 0.0f));
        } else {
            const float d = x - t;
            s_pred += d * d;
        }
    }
    if (w_kld != 0.0f)
        for (int64_t i = threadIdx.x; i < n_lat; i += blockDim.x) {
            const float m = mu[i], lv = logvar[i];
            s_kld += 1.0f + lv - m * m - expf(lv);
        }
    for (int i = threadIdx.x; i < nparts; i += blockDim.x) s_mse += partial[i];
    const float pred = block_sum(s_pred, red);
    const float kld = block_sum(s_kld, red);
    const float mse = block_sum(s_mse, red);
    if (threadIdx.x == 0) {
        const float p = pred / (float)B;
        const float k = (w_kld != 0.0f) ? -0.5f * kld / (float)n_lat : 0.0f;
        const float m = (w_mse != 0.0f) ? mse / (float)n_recon : 0.0f;
        out[0] = w_pred * p + w_mse * m + w_kld * k;
        out[1] = p; out[2] = m; out[3] = k;
    }
}

__global__ void __launch_bounds__(IS_THREADS)
loss_bwd_recon_kernel(const float* __restrict__ recon, const float* __restrict__ seq, int64_t n,
                      const float* __restrict__ gout, float coef, float* __restrict__ g_recon) {
    const float c = coef * gout[0];
    const int64_t n4 = n >> 2;
    const float4* a4 = reinterpret_cast<const float4*>(recon);
    const float4* b4 = reinterpret_cast<const float4*>(seq);
    float4* g4 = reinterpret_cast<float4*>(g_recon);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 x = __ldg(a4 + i), y = __ldg(b4 + i);
        g4[i] = make_float4(c * (x.x - y.x), c * (x.y - y.y), c * (x.z - y.z), c * (x.w - y.w));
    }
    if (blockIdx.x == 0)
        for (int64_t i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) g_recon[i] = c * (recon[i] - seq[i]);
}

__global__ void __launch_bounds__(IS_THREADS)
loss_bwd_small_kernel(const float* __restrict__ mu, const float* __restrict__ logvar, int64_t n_lat,
                      const float* __restrict__ logits, const float* __restrict__ y, int64_t B, int mode, float pos_weight,
                      float w_pred, float w_kld, const float* __restrict__ gout,
                      float* __restrict__ g_mu, float* __restrict__ g_logvar, float* __restrict__ g_logits) {
    const float go = gout[0];
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B) {
        const float x = logits[i], t = y[i];
        float g;
        if (mode == 0) {
            const float lw = 1.0f + (pos_weight - 1.0f) * t;
            const float sig = 1.0f / (1.0f + expf(-x));
            g = (1.0f - t) - lw * (1.0f - sig);
        } else {
            g = 2.0f * (x - t);
        }
        g_logits[i] = go * w_pred * g / (float)B;
    }
    if (g_mu && i < n_lat) {
        const float c = go * w_kld / (float)n_lat;
        g_mu[i] = c * mu[i];
        g_logvar[i] = c * (-0.5f) * (1.0f - expf(logvar[i]));
    }
}

}  // namespace is

using namespace is;

extern "C" {

int is_loss_num_partials(void) { return 592; }    // 148 SMs x 4

// recon/seq may be NULL when w_mse == 0 (no --sequence-loss); partial: float [is_loss_num_partials()]
int is_loss_fwd(const float* recon, const float* seq, int64_t n_recon, const float* mu, const float* logvar,
                int64_t n_lat, const float* logits, const float* y, int64_t B, int mode, float pos_weight,
                float w_pred, float w_mse, float w_kld, float* partial, float* out, void* stream) {
    if (B <= 0 || (mode != 0 && mode != 1)) return IS_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    int nparts = 0;
    if (w_mse != 0.0f) {
        if (!recon || !seq || n_recon <= 0) return IS_ERR_ARG;
        if ((((uintptr_t)recon) | ((uintptr_t)seq)) & 15) return IS_ERR_ARG;
        nparts = is_loss_num_partials();
        mse_partial_kernel<<<nparts, IS_THREADS, 0, st>>>(recon, seq, n_recon, partial);
        IS_LAUNCH_CHECK();
    }
    loss_final_kernel<<<1, IS_THREADS, 0, st>>>(partial, nparts, n_recon, mu, logvar, n_lat, logits, y, B, mode,
                                                pos_weight, w_pred, w_mse, w_kld, out);
    IS_LAUNCH_CHECK();
    return IS_OK;
}

// gout: device scalar (upstream gradient).  g_recon / g_mu / g_logvar may be NULL when unweighted.
int is_loss_bwd(const float* recon, const float* seq, int64_t n_recon, const float* mu, const float* logvar,
                int64_t n_lat, const float* logits, const float* y, int64_t B, int mode, float pos_weight,
                float w_pred, float w_mse, float w_kld, const float* gout,
                float* g_recon, float* g_mu, float* g_logvar, float* g_logits, void* stream) {
    if (B <= 0 || (mode != 0 && mode != 1)) return IS_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (w_mse != 0.0f && g_recon) {
        if ((((uintptr_t)recon) | ((uintptr_t)seq) | ((uintptr_t)g_recon)) & 15) return IS_ERR_ARG;
        loss_bwd_recon_kernel<<<592, IS_THREADS, 0, st>>>(recon, seq, n_recon, gout, 2.0f * w_mse / (float)n_recon, g_recon);
        IS_LAUNCH_CHECK();
    }
    const int64_t m = (w_kld != 0.0f && g_mu) ? (B > n_lat ? B : n_lat) : B;
    loss_bwd_small_kernel<<<(unsigned)((m + IS_THREADS - 1) / IS_THREADS), IS_THREADS, 0, st>>>(
        mu, logvar, n_lat, logits, y, B, mode, pos_weight, w_pred, w_kld, gout,
        (w_kld != 0.0f) ? g_mu : nullptr, g_logvar, g_logits);
    IS_LAUNCH_CHECK();
    return IS_OK;
}

}  // extern "C"
