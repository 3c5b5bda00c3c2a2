// Fusion attention over the fused scalars, collapsed to its closed form.
//
// Reference: ``combined_attention = MultiHeadAttention(D, n_head=8, input_dim=1)`` applied to the
// 104 (hybrid, D=16) or 208 (comparative, D=32) fused scalars treated as a sequence of 1-d tokens,
// followed by ``mean(dim=2)`` (immunostruct/models/hybrid_models.py:275,344-347;
// comparative_models.py:392,484-486; attention block layers.py:29-48,67-78).
//
// Because every token is a single scalar c_i, q_i = wq c_i + bq, k_j = wk c_j + bk, v_j = wv c_j + bv,
// and the per-head score is  s_ij = (A_h c_i + C_h) c_j + (terms constant in j)  with
// A_h = sum_{k in h} wq_k wk_k / sqrt(dh),  C_h = sum_{k in h} bq_k wk_k / sqrt(dh); softmax over j
// is invariant to the j-constant terms.  The output projection followed by the channel mean is a
// single linear functional of the concatenated heads, so
//     out_i = btilde + sum_h ( alpha_h * E^h_i[c] + beta_h ),   E^h_i[c] = sum_j softmax_j(gamma c_j) c_j,
//     gamma = A_h c_i + C_h,  alpha_h = sum_{k in h} wtilde_k wv_k,  beta_h = sum_{k in h} wtilde_k bv_k,
//     wtilde = column means of w_concat.weight,  btilde = mean(w_concat.bias).
// The host computes the 4H+1 coefficients with ordinary differentiable torch ops
// (immunostruct_b200/layers.py: FusionAttention); these kernels do the O(L^2 H) part, one CTA per
// sample, and its backward.  coef layout: [A(H) | C(H) | alpha(H) | beta(H) | btilde].
#include "common.cuh"

#define IS_FUS_LMAX 256
#define IS_FUS_HMAX 8

namespace is {

__device__ __forceinline__ void fusion_pair(const float* __restrict__ c, int L, float gamma, float cmax, float cmin,
                                            float& mx, float& Z, float& S1, float& S2) {
    mx = gamma > 0.0f ? gamma * cmax : gamma * cmin;
    Z = 0.0f; S1 = 0.0f; S2 = 0.0f;
    for (int j = 0; j < L; ++j) {
        const float cj = c[j];
        const float e = exp_comp(gamma * cj - mx);       // one MUFU, product rounding compensated (common.cuh)
        Z += e; S1 = fmaf(e, cj, S1); S2 = fmaf(e * cj, cj, S2);
    }
}

__global__ void __launch_bounds__(IS_THREADS)
fusion_fwd_kernel(const float* __restrict__ cin, int L, int H, const float* __restrict__ coef, float* __restrict__ out) {
    __shared__ float c[IS_FUS_LMAX];
    __shared__ float Eh[IS_FUS_LMAX * IS_FUS_HMAX];
    __shared__ float s_mm[2];
    __shared__ float red[2 * 8];
    const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float lmax = -INFINITY, lmin = INFINITY;
    for (int i = tid; i < L; i += IS_THREADS) {
        const float v = cin[(int64_t)b * L + i];
        c[i] = v; lmax = fmaxf(lmax, v); lmin = fminf(lmin, v);
    }
    lmax = warp_max(lmax); lmin = -warp_max(-lmin);
    if (lane == 0) { red[warp] = lmax; red[8 + warp] = lmin; }
    __syncthreads();
    if (tid == 0) {
        float a = red[0], m = red[8];
        for (int w = 1; w < 8; ++w) { a = fmaxf(a, red[w]); m = fminf(m, red[8 + w]); }
        s_mm[0] = a; s_mm[1] = m;
    }
    __syncthreads();
    const float cmax = s_mm[0], cmin = s_mm[1];
    for (int idx = tid; idx < L * H; idx += IS_THREADS) {
        const int i = idx / H, h = idx - i * H;
        const float gamma = coef[h] * c[i] + coef[H + h];
        float mx, Z, S1, S2;
        fusion_pair(c, L, gamma, cmax, cmin, mx, Z, S1, S2);
        Eh[idx] = S1 / Z;
    }
    __syncthreads();
    for (int i = tid; i < L; i += IS_THREADS) {
        float o = coef[4 * H];
        for (int h = 0; h < H; ++h) o += coef[2 * H + h] * Eh[i * H + h] + coef[3 * H + h];
        out[(int64_t)b * L + i] = o;
    }
}

// gcoef_part [B, 4H]: per-sample partial sums for (A, C, alpha, beta); the host sums over B.
__global__ void __launch_bounds__(IS_THREADS)
fusion_bwd_kernel(const float* __restrict__ cin, int L, int H, const float* __restrict__ coef,
                  const float* __restrict__ gout, float* __restrict__ gc, float* __restrict__ gcoef_part) {
    __shared__ float c[IS_FUS_LMAX];
    __shared__ float go[IS_FUS_LMAX];
    __shared__ float gdir[IS_FUS_LMAX];                       // direct path: sum_h ggamma_ih A_h
    __shared__ float s_gam[IS_FUS_LMAX * IS_FUS_HMAX];
    __shared__ float s_mx[IS_FUS_LMAX * IS_FUS_HMAX];
    __shared__ float s_rz[IS_FUS_LMAX * IS_FUS_HMAX];
    __shared__ float s_E[IS_FUS_LMAX * IS_FUS_HMAX];
    __shared__ float s_gg[IS_FUS_LMAX * IS_FUS_HMAX];         // ggamma_ih
    __shared__ float s_mm[2];
    __shared__ float red[2 * 8];
    const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float lmax = -INFINITY, lmin = INFINITY;
    for (int i = tid; i < L; i += IS_THREADS) {
        const float v = cin[(int64_t)b * L + i];
        c[i] = v; go[i] = gout[(int64_t)b * L + i];
        lmax = fmaxf(lmax, v); lmin = fminf(lmin, v);
    }
    lmax = warp_max(lmax); lmin = -warp_max(-lmin);
    if (lane == 0) { red[warp] = lmax; red[8 + warp] = lmin; }
    __syncthreads();
    if (tid == 0) {
        float a = red[0], m = red[8];
        for (int w = 1; w < 8; ++w) { a = fmaxf(a, red[w]); m = fminf(m, red[8 + w]); }
        s_mm[0] = a; s_mm[1] = m;
    }
    __syncthreads();
    const float cmax = s_mm[0], cmin = s_mm[1];
    for (int idx = tid; idx < L * H; idx += IS_THREADS) {
        const int i = idx / H, h = idx - i * H;
        const float gamma = coef[h] * c[i] + coef[H + h];
        float mx, Z, S1, S2;
        fusion_pair(c, L, gamma, cmax, cmin, mx, Z, S1, S2);
        const float rz = 1.0f / Z, E = S1 * rz;
        // variance by a second pass around the mean: E[c^2] - E[c]^2 cancels catastrophically in fp32
        float var = 0.0f;
        for (int j = 0; j < L; ++j) {
            const float dcj = c[j] - E;
            var = fmaf(exp_comp(gamma * c[j] - mx), dcj * dcj, var);
        }
        var *= rz;
        s_gam[idx] = gamma; s_mx[idx] = mx; s_rz[idx] = rz; s_E[idx] = E;
        s_gg[idx] = go[i] * coef[2 * H + h] * var;             // dL/dgamma_ih
    }
    __syncthreads();
    // per-sample coefficient partials (fixed order over i)
    if (tid < 4 * H) {
        const int kind = tid / H, h = tid - kind * H;
        float s = 0.0f;
        for (int i = 0; i < L; ++i) {
            const float gg = s_gg[i * H + h];
            s += kind == 0 ? gg * c[i] : kind == 1 ? gg : kind == 2 ? go[i] * s_E[i * H + h] : go[i];
        }
        gcoef_part[(int64_t)b * 4 * H + tid] = s;
    }
    for (int i = tid; i < L; i += IS_THREADS) {
        float s = 0.0f;
        for (int h = 0; h < H; ++h) s += s_gg[i * H + h] * coef[h];
        gdir[i] = s;
    }
    __syncthreads();
    // dL/dc_j = gdir_j + sum_{i,h} go_i alpha_h p_ij (1 + gamma_ih (c_j - E_ih)); two threads per j (the halves of the i
    // range), combined in a fixed order
    float* part = s_mx;                                       // re-used only after the loops that read it
    const int half = tid & 1;
    float sacc[2 * IS_FUS_LMAX / IS_THREADS];                 // L <= 256: two passes of 128 values of j
#pragma unroll
    for (int ps = 0; ps < 2 * IS_FUS_LMAX / IS_THREADS; ++ps) {
        const int j = (tid >> 1) + ps * (IS_THREADS / 2);
        float s = 0.0f;
        if (j < L) {
            const float cj = c[j];
            const int ib = half ? (L + 1) / 2 : 0, ie = half ? L : (L + 1) / 2;
            for (int i = ib; i < ie; ++i) {
                const float goi = go[i];
                for (int h = 0; h < H; ++h) {
                    const int idx = i * H + h;
                    const float gam = s_gam[idx];
                    const float p = exp_comp(gam * cj - s_mx[idx]) * s_rz[idx];
                    s = fmaf(goi * coef[2 * H + h] * p, 1.0f + gam * (cj - s_E[idx]), s);
                }
            }
        }
        sacc[ps] = s;
    }
    __syncthreads();
#pragma unroll
    for (int ps = 0; ps < 2 * IS_FUS_LMAX / IS_THREADS; ++ps) {
        const int j = (tid >> 1) + ps * (IS_THREADS / 2);
        if (j < L && half) part[j] = sacc[ps];
    }
    __syncthreads();
#pragma unroll
    for (int ps = 0; ps < 2 * IS_FUS_LMAX / IS_THREADS; ++ps) {
        const int j = (tid >> 1) + ps * (IS_THREADS / 2);
        if (j < L && !half) gc[(int64_t)b * L + j] = gdir[j] + (sacc[ps] + part[j]);
    }
}

}  // namespace is

using namespace is;

extern "C" {

int is_fusion_attn_fwd(const float* c, int n_samples, int L, int n_head, const float* coef, float* out, void* stream) {
    if (n_samples <= 0 || L <= 0 || L > IS_FUS_LMAX || n_head <= 0 || n_head > IS_FUS_HMAX) return IS_ERR_ARG;
    fusion_fwd_kernel<<<n_samples, IS_THREADS, 0, (cudaStream_t)stream>>>(c, L, n_head, coef, out);
    IS_LAUNCH_CHECK();
    return IS_OK;
}

int is_fusion_attn_bwd(const float* c, int n_samples, int L, int n_head, const float* coef, const float* gout,
                       float* gc, float* gcoef_part, void* stream) {
    if (n_samples <= 0 || L <= 0 || L > IS_FUS_LMAX || n_head <= 0 || n_head > IS_FUS_HMAX) return IS_ERR_ARG;
    fusion_bwd_kernel<<<n_samples, IS_THREADS, 0, (cudaStream_t)stream>>>(c, L, n_head, coef, gout, gc, gcoef_part);
    IS_LAUNCH_CHECK();
    return IS_OK;
}

}  // extern "C"
