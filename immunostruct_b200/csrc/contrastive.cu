// PairedContrastiveLoss forward + backward (utils/contrastive.py:37-83; built at procedures/train.py:74-78).
//
//   gate   : exactly two distinct target values in the batch, else the loss is 0 (:38-43); is_imm = t > mean(t)
//   z_s    = projector(E_s) = relu(BN_train(E_s W1^T)) W2^T          s in {cancer, wt}   (:27-32,47-48)
//   zc_s   = z_s - mean_b z_s                                          (:54-55)
//   std    = sqrt(var_unbiased(zc_s) + 1e-4);  L_std = sum_s mean_j relu(1 - std_sj) / 2        (:58-61)
//   S      = zc_c zc_w^T / Z   [B,B];   L_pair = sum_ij w_ij (S_ij - [i == j] is_imm_i)^2       (:64,70-74)
//   C      = zc_c^T zc_w / B   [Z,Z];   L_corr = sum_kl w_kl (C_kl - [k == l])^2                (:67,77-80)
//            (w = 1 on the diagonal, lambda_off_diag elsewhere)
//   loss   = gate (L_pair + L_corr + L_std);  BatchNorm running statistics are updated (momentum 0.1, unbiased
//            variance) for the cancer call and then for the wild-type call -- only when the gate is open, as in the
//            reference, which returns before touching the projector.
//
// The whole thing is ~56 MFLOP per 256-pair batch: latency-, not throughput-bound.  One strided SIMT fp32 tile GEMM
// (32 x 32 output tile, 256 threads, two optional terms, operand transforms and loss epilogues) covers the eight
// GEMM-shaped products; four small column-statistics kernels cover BatchNorm, centring and their backward.  No
// floating-point atomics: per-CTA loss partials are summed in CTA order.
#include "common.cuh"

namespace is {
namespace ctr {

constexpr int TS = 32;

struct Term {
    const float* A; int64_t sam, sak;      // A(m, k) = A[m * sam + k * sak]
    const float* B; int64_t sbk, sbn;      // B(k, n) = B[k * sbk + n * sbn]
    int K;
    float alpha;
};

struct BnView {                            // relu(gamma (y - mu) rstd + beta), per side
    const float* mu;                       // [2][Z]
    const float* rstd;                     // [2][Z]
    const float* gamma;                    // [Z]
    const float* beta;                     // [Z]
    int Z;
    int rows_per_side;                     // B (row index / B = side)
};

__device__ __forceinline__ float bn_pre(const BnView& v, float y, int side, int col) {
    return fmaf(__ldg(v.gamma + col) * (y - __ldg(v.mu + side * v.Z + col)), __ldg(v.rstd + side * v.Z + col), __ldg(v.beta + col));
}

enum { EPI_STORE = 0, EPI_PAIR = 1, EPI_CORR = 2, EPI_RELU_MASK = 3 };
enum { XF_NONE = 0, XF_A_BNRELU = 1, XF_B_BNRELU = 2 };

struct Epi {
    float* C; int64_t ldc;                 // output (EPI_PAIR / EPI_CORR: the gradient matrix dS / dC)
    int accumulate;                        // C += result
    const float* imm;                      // EPI_PAIR: is_imm [B]
    float lambda, grad_scale;              // off-diagonal weight; 2 / Z or 2 / B
    float* partial;                        // per-CTA loss partial [gridDim.y * gridDim.x]
    const float* y;                        // EPI_RELU_MASK: pre-activation source y [rows, Z]
    int side;                              // EPI_RELU_MASK / XF_A: side of this launch (blockIdx.z added)
};

// grid = (ceil(N/32), ceil(M/32), batch); per-z pointer offsets in elements
template <int XF, int EPI>
__global__ void __launch_bounds__(256)
gemm_kernel(Term t1, Term t2, Epi ep, BnView bn, int M, int N, int64_t zA, int64_t zB, int64_t zC) {
    __shared__ float As[TS][TS + 1], Bs[TS][TS + 1];
    __shared__ float red[256];
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;      // thread: column tx, rows ty + 8 i
    const int m0 = blockIdx.y * TS, n0 = blockIdx.x * TS, z = blockIdx.z;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
    for (int term = 0; term < 2; ++term) {
        const Term& t = term == 0 ? t1 : t2;
        if (t.K <= 0) continue;
        const float* A = t.A + z * zA;
        const float* B = t.B + z * zB;
        float part[4] = {0.f, 0.f, 0.f, 0.f};
        for (int k0 = 0; k0 < t.K; k0 += TS) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int idx = tid + 256 * i;
                // A tile: fastest thread index follows the unit stride
                int am, ak;
                if (t.sak == 1) { am = idx >> 5; ak = idx & 31; } else { ak = idx >> 5; am = idx & 31; }
                float av = 0.0f;
                if (m0 + am < M && k0 + ak < t.K) {
                    av = __ldg(A + (int64_t)(m0 + am) * t.sam + (int64_t)(k0 + ak) * t.sak);
                    if (XF == XF_A_BNRELU) av = fmaxf(bn_pre(bn, av, ep.side + z, k0 + ak), 0.0f);
                }
                As[am][ak] = av;
                int bk, bn_;
                if (t.sbn == 1) { bk = idx >> 5; bn_ = idx & 31; } else { bn_ = idx >> 5; bk = idx & 31; }
                float bv = 0.0f;
                if (n0 + bn_ < N && k0 + bk < t.K) {
                    bv = __ldg(B + (int64_t)(k0 + bk) * t.sbk + (int64_t)(n0 + bn_) * t.sbn);
                    if (XF == XF_B_BNRELU) bv = fmaxf(bn_pre(bn, bv, (k0 + bk) / bn.rows_per_side, n0 + bn_), 0.0f);
                }
                Bs[bk][bn_] = bv;
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < TS; ++k) {
                const float b = Bs[k][tx];
#pragma unroll
                for (int i = 0; i < 4; ++i) part[i] = fmaf(As[ty + 8 * i][k], b, part[i]);
            }
            __syncthreads();
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i] = fmaf(t.alpha, part[i], acc[i]);
    }
    float* C = ep.C + z * zC;
    float lsum = 0.0f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty + 8 * i, n = n0 + tx;
        if (m >= M || n >= N) continue;
        float v = acc[i];
        if (EPI == EPI_PAIR || EPI == EPI_CORR) {
            const float ideal = (m == n) ? (EPI == EPI_PAIR ? __ldg(ep.imm + m) : 1.0f) : 0.0f;
            const float w = (m == n) ? 1.0f : ep.lambda;
            const float d = v - ideal;
            lsum = fmaf(w * d, d, lsum);
            C[(int64_t)m * ep.ldc + n] = ep.grad_scale * w * d;
        } else if (EPI == EPI_RELU_MASK) {
            const float pre = bn_pre(bn, __ldg(ep.y + z * zA + (int64_t)m * bn.Z + n), ep.side + z, n);
            C[(int64_t)m * ep.ldc + n] = pre > 0.0f ? v : 0.0f;
        } else {
            if (ep.accumulate) v += C[(int64_t)m * ep.ldc + n];
            C[(int64_t)m * ep.ldc + n] = v;
        }
    }
    if (EPI == EPI_PAIR || EPI == EPI_CORR) {
        red[tid] = lsum;
        __syncthreads();
        for (int s = 128; s > 0; s >>= 1) {
            if (tid < s) red[tid] += red[tid + s];
            __syncthreads();
        }
        if (tid == 0) ep.partial[blockIdx.y * gridDim.x + blockIdx.x] = red[0];
    }
}

// ---- gate: scal[0] = gate (1 / 0), imm[i] = t_i > mean(t).  One CTA. ------------------------------------------
__global__ void __launch_bounds__(256)
gate_kernel(const float* __restrict__ t, int n, float* __restrict__ imm, float* __restrict__ scal) {
    __shared__ float s_lo[256], s_hi[256], s_sum[256];
    __shared__ int s_bad[256];
    const int tid = threadIdx.x;
    float lo = INFINITY, hi = -INFINITY, sum = 0.0f;
    for (int i = tid; i < n; i += 256) { const float v = t[i]; lo = fminf(lo, v); hi = fmaxf(hi, v); sum += v; }
    s_lo[tid] = lo; s_hi[tid] = hi; s_sum[tid] = sum;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (tid < s) { s_lo[tid] = fminf(s_lo[tid], s_lo[tid + s]); s_hi[tid] = fmaxf(s_hi[tid], s_hi[tid + s]); s_sum[tid] += s_sum[tid + s]; }
        __syncthreads();
    }
    lo = s_lo[0]; hi = s_hi[0];
    const float mean = s_sum[0] / (float)n;
    int bad = 0;
    for (int i = tid; i < n; i += 256) {
        const float v = t[i];
        bad |= (v != lo && v != hi) ? 1 : 0;
        imm[i] = v > mean ? 1.0f : 0.0f;
    }
    s_bad[tid] = bad;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (tid < s) s_bad[tid] |= s_bad[tid + s];
        __syncthreads();
    }
    if (tid == 0) scal[0] = (lo != hi && !s_bad[0]) ? 1.0f : 0.0f;
}

// column statistics helper: 32 columns x 8 row lanes per CTA; returns the sum over rows of f(row) for column
// c0 + tx in every thread with ty == 0 (fixed order)
template <typename F>
__device__ __forceinline__ float col_reduce(F f, int rows, float (*red)[33]) {
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    float s = 0.0f;
    for (int r = ty; r < rows; r += 8) s += f(r);
    red[ty][tx] = s;
    __syncthreads();
    float tot = 0.0f;
    if (ty == 0)
        for (int l = 0; l < 8; ++l) tot += red[l][tx];
    __syncthreads();
    return tot;
}

// BatchNorm batch statistics of y [2][B][Z] + running-statistics update (cancer call, then wild-type call)
__global__ void __launch_bounds__(256)
bn_stats_kernel(const float* __restrict__ y, int B, int Z, float bn_eps, float momentum, const float* __restrict__ scal,
                float* __restrict__ mu, float* __restrict__ rstd, float* __restrict__ run_mean, float* __restrict__ run_var,
                int64_t* __restrict__ n_tracked) {
    __shared__ float red[8][33];
    __shared__ float s_mu[32];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5, c = blockIdx.x * 32 + tx;
    const bool ok = c < Z;
    const float gate = scal[0];
    for (int side = 0; side < 2; ++side) {
        const float* ys = y + (int64_t)side * B * Z;
        const float sum = col_reduce([&](int r) { return ok ? __ldg(ys + (int64_t)r * Z + c) : 0.0f; }, B, red);
        if (ty == 0) s_mu[tx] = sum / (float)B;
        __syncthreads();
        const float m = s_mu[tx];
        const float ss = col_reduce([&](int r) { const float d = ok ? __ldg(ys + (int64_t)r * Z + c) - m : 0.0f; return d * d; }, B, red);
        if (ty == 0 && ok) {
            const float var = ss / (float)B;
            mu[side * Z + c] = m;
            rstd[side * Z + c] = rsqrtf(var + bn_eps);
            if (run_mean != nullptr && gate != 0.0f) {
                const float unb = ss / (float)max(B - 1, 1);
                run_mean[c] = (1.0f - momentum) * run_mean[c] + momentum * m;
                run_var[c] = (1.0f - momentum) * run_var[c] + momentum * unb;
            }
        }
        __syncthreads();
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && n_tracked != nullptr && gate != 0.0f) *n_tracked += 2;
}

// centre z [2][B][Z] in place, std per (side, column), std-loss partial per CTA
__global__ void __launch_bounds__(256)
z_stats_kernel(float* __restrict__ z, int B, int Z, float* __restrict__ stdv, float* __restrict__ partial) {
    __shared__ float red[8][33];
    __shared__ float s_mu[32];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5, c = blockIdx.x * 32 + tx;
    const bool ok = c < Z;
    float loss = 0.0f;
    for (int side = 0; side < 2; ++side) {
        float* zs = z + (int64_t)side * B * Z;
        const float sum = col_reduce([&](int r) { return ok ? zs[(int64_t)r * Z + c] : 0.0f; }, B, red);
        if (ty == 0) s_mu[tx] = sum / (float)B;
        __syncthreads();
        const float m = s_mu[tx];
        const float ss = col_reduce([&](int r) {
            if (!ok) return 0.0f;
            const float d = zs[(int64_t)r * Z + c] - m;
            zs[(int64_t)r * Z + c] = d;
            return d * d; }, B, red);
        if (ty == 0 && ok) {
            const float sd = sqrtf(ss / (float)max(B - 1, 1) + 1e-4f);
            stdv[side * Z + c] = sd;
            loss += fmaxf(1.0f - sd, 0.0f) / (2.0f * (float)Z);
        }
        __syncthreads();
    }
    // ty == 0 threads hold the per-column losses: sum over the CTA's 32 columns in lane order
    if (ty == 0) {
        red[0][tx] = loss;
        __syncwarp();
        if (tx == 0) {
            float s = 0.0f;
            for (int i = 0; i < 32; ++i) s += red[0][i];
            partial[blockIdx.x] = s;
        }
    }
}

// out[0] = gate * (sum pair partials + sum corr partials + sum std partials); out[1..3] = the three terms
__global__ void finalize_kernel(const float* __restrict__ p_pair, int n_pair, const float* __restrict__ p_corr, int n_corr,
                                const float* __restrict__ p_std, int n_std, const float* __restrict__ scal, float* __restrict__ out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    float a = 0.0f, b = 0.0f, c = 0.0f;
    for (int i = 0; i < n_pair; ++i) a += p_pair[i];
    for (int i = 0; i < n_corr; ++i) b += p_corr[i];
    for (int i = 0; i < n_std; ++i) c += p_std[i];
    out[0] = scal[0] * (a + b + c);
    out[1] = a; out[2] = b; out[3] = c;
}

// backward of the centring + std hinge:  gz = scale * (G + g_std - colmean(G + g_std)),  in place on G [2][B][Z]
__global__ void __launch_bounds__(256)
center_bwd_kernel(float* __restrict__ G, const float* __restrict__ zc, const float* __restrict__ stdv, int B, int Z,
                  const float* __restrict__ scal, const float* __restrict__ gout) {
    __shared__ float red[8][33];
    __shared__ float s_mu[32];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5, c = blockIdx.x * 32 + tx;
    const bool ok = c < Z;
    const float scale = scal[0] * gout[0];
    for (int side = 0; side < 2; ++side) {
        float* gs = G + (int64_t)side * B * Z;
        const float* zs = zc + (int64_t)side * B * Z;
        const float sd = ok ? stdv[side * Z + c] : 1.0f;
        const float k = (ok && sd < 1.0f) ? -1.0f / (2.0f * (float)Z * sd * (float)max(B - 1, 1)) : 0.0f;
        const float sum = col_reduce([&](int r) {
            if (!ok) return 0.0f;
            const float v = fmaf(k, zs[(int64_t)r * Z + c], gs[(int64_t)r * Z + c]);
            gs[(int64_t)r * Z + c] = v;
            return v; }, B, red);
        if (ty == 0) s_mu[tx] = sum / (float)B;
        __syncthreads();
        const float m = s_mu[tx];
        if (ok)
            for (int r = ty; r < B; r += 8) gs[(int64_t)r * Z + c] = scale * (gs[(int64_t)r * Z + c] - m);
        __syncthreads();
    }
}

// BatchNorm backward on gpre [2][B][Z] in place -> gy; g_gamma / g_beta [Z] summed over both sides
__global__ void __launch_bounds__(256)
bn_bwd_kernel(float* __restrict__ gpre, const float* __restrict__ y, const float* __restrict__ mu, const float* __restrict__ rstd,
              const float* __restrict__ gamma, int B, int Z, float* __restrict__ g_gamma, float* __restrict__ g_beta) {
    __shared__ float red[8][33];
    __shared__ float s_a[32], s_b[32];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5, c = blockIdx.x * 32 + tx;
    const bool ok = c < Z;
    float gg = 0.0f, gb = 0.0f;
    for (int side = 0; side < 2; ++side) {
        float* gs = gpre + (int64_t)side * B * Z;
        const float* ys = y + (int64_t)side * B * Z;
        const float m = ok ? mu[side * Z + c] : 0.0f, rs = ok ? rstd[side * Z + c] : 0.0f;
        const float s1 = col_reduce([&](int r) { return ok ? gs[(int64_t)r * Z + c] : 0.0f; }, B, red);
        const float s2 = col_reduce([&](int r) { return ok ? gs[(int64_t)r * Z + c] * ((ys[(int64_t)r * Z + c] - m) * rs) : 0.0f; }, B, red);
        if (ty == 0) { s_a[tx] = s1; s_b[tx] = s2; gg += s2; gb += s1; }
        __syncthreads();
        const float a = s_a[tx] / (float)B, b = s_b[tx] / (float)B;
        if (ok) {
            const float gr = gamma[c] * rs;
            for (int r = ty; r < B; r += 8) {
                const float yh = (ys[(int64_t)r * Z + c] - m) * rs;
                gs[(int64_t)r * Z + c] = gr * (gs[(int64_t)r * Z + c] - a - yh * b);
            }
        }
        __syncthreads();
    }
    if (ty == 0 && ok) { g_gamma[c] = gg; g_beta[c] = gb; }
}

template <int XF, int EPI>
static void launch_gemm(const Term& t1, const Term& t2, const Epi& ep, const BnView& bn, int M, int N, int batch,
                        int64_t zA, int64_t zB, int64_t zC, cudaStream_t st) {
    dim3 grid((N + TS - 1) / TS, (M + TS - 1) / TS, batch);
    gemm_kernel<XF, EPI><<<grid, 256, 0, st>>>(t1, t2, ep, bn, M, N, zA, zB, zC);
}

}  // namespace ctr
}  // namespace is

using namespace is;
using namespace is::ctr;

extern "C" {

// floats of scratch the forward needs (kept for the backward): y, zc [2][B][Z] each; dS [B][B]; dC [Z][Z];
// mu, rstd, std [2][Z] each; imm [B]; scal [4]; loss partials
int64_t is_contrastive_scratch_floats(int B, int Z) {
    const int64_t tb = (B + 31) / 32, tz = (Z + 31) / 32;
    return 4 * (int64_t)B * Z + (int64_t)B * B + (int64_t)Z * Z + 6 * (int64_t)Z + B + 4 + tb * tb + tz * tz + tz + 64;
}

struct CtrLayout {
    float *y, *zc, *dS, *dC, *mu, *rstd, *stdv, *imm, *scal, *p_pair, *p_corr, *p_std;
    int n_pair, n_corr, n_std;
};
static CtrLayout ctr_layout(float* s, int B, int Z) {
    CtrLayout L;
    const int64_t tb = (B + 31) / 32, tz = (Z + 31) / 32;
    L.y = s; s += 2 * (int64_t)B * Z;
    L.zc = s; s += 2 * (int64_t)B * Z;
    L.dS = s; s += (int64_t)B * B;
    L.dC = s; s += (int64_t)Z * Z;
    L.mu = s; s += 2 * Z;
    L.rstd = s; s += 2 * Z;
    L.stdv = s; s += 2 * Z;
    L.imm = s; s += B;
    L.scal = s; s += 4;
    L.p_pair = s; s += tb * tb; L.n_pair = (int)(tb * tb);
    L.p_corr = s; s += tz * tz; L.n_corr = (int)(tz * tz);
    L.p_std = s; L.n_std = (int)tz;
    return L;
}

// Forward.  Ec, Ew [B, D]; target [B]; W1 [Z, D]; gamma, beta [Z]; W2 [Z, Z]; run_mean / run_var [Z] and n_tracked (int64)
// may be NULL.  out [4] = {loss, pair term, corr term, std term}.  scratch: is_contrastive_scratch_floats(B, Z) floats,
// kept unchanged until the backward.  11 launches.
int is_contrastive_fwd(const float* Ec, const float* Ew, const float* target, int B, int D, int Z, const float* W1,
                       const float* gamma, const float* beta, const float* W2, float bn_eps, float momentum,
                       float* run_mean, float* run_var, int64_t* n_tracked, float lambda_off, float* scratch, float* out,
                       void* stream) {
    if (B < 2 || D <= 0 || Z <= 0) return IS_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    CtrLayout L = ctr_layout(scratch, B, Z);
    const int64_t BZ = (int64_t)B * Z;
    const Term none = {nullptr, 0, 0, nullptr, 0, 0, 0, 0.0f};
    BnView bn = {L.mu, L.rstd, gamma, beta, Z, B};
    gate_kernel<<<1, 256, 0, st>>>(target, B, L.imm, L.scal);
    // y_s = E_s W1^T
    for (int s = 0; s < 2; ++s) {
        Term t = {s == 0 ? Ec : Ew, D, 1, W1, 1, D, D, 1.0f};
        Epi ep = {L.y + s * BZ, Z, 0, nullptr, 0.f, 0.f, nullptr, nullptr, 0};
        launch_gemm<XF_NONE, EPI_STORE>(t, none, ep, bn, B, Z, 1, 0, 0, 0, st);
    }
    const int tz = (Z + 31) / 32;
    bn_stats_kernel<<<tz, 256, 0, st>>>(L.y, B, Z, bn_eps, momentum, L.scal, L.mu, L.rstd, run_mean, run_var, n_tracked);
    {   // z_s = relu(bn(y_s)) W2^T   (both sides in one launch: blockIdx.z = side)
        Term t = {L.y, Z, 1, W2, 1, Z, Z, 1.0f};
        Epi ep = {L.zc, Z, 0, nullptr, 0.f, 0.f, nullptr, nullptr, 0};
        launch_gemm<XF_A_BNRELU, EPI_STORE>(t, none, ep, bn, B, Z, 2, BZ, 0, BZ, st);
    }
    z_stats_kernel<<<tz, 256, 0, st>>>(L.zc, B, Z, L.stdv, L.p_std);
    {   // S = zc_c zc_w^T / Z  -> dS, pair partials
        Term t = {L.zc, Z, 1, L.zc + BZ, 1, Z, Z, 1.0f / (float)Z};
        Epi ep = {L.dS, B, 0, L.imm, lambda_off, 2.0f / (float)Z, L.p_pair, nullptr, 0};
        launch_gemm<XF_NONE, EPI_PAIR>(t, none, ep, bn, B, B, 1, 0, 0, 0, st);
    }
    {   // C = zc_c^T zc_w / B  -> dC, corr partials
        Term t = {L.zc, 1, Z, L.zc + BZ, Z, 1, B, 1.0f / (float)B};
        Epi ep = {L.dC, Z, 0, nullptr, lambda_off, 2.0f / (float)B, L.p_corr, nullptr, 0};
        launch_gemm<XF_NONE, EPI_CORR>(t, none, ep, bn, Z, Z, 1, 0, 0, 0, st);
    }
    finalize_kernel<<<1, 32, 0, st>>>(L.p_pair, L.n_pair, L.p_corr, L.n_corr, L.p_std, L.n_std, L.scal, out);
    IS_LAUNCH_CHECK();
    return IS_OK;
}

// Backward.  gout [1] (device).  work: 2 * B * Z floats.  Outputs: gEc, gEw [B, D]; gW1 [Z, D]; g_gamma, g_beta [Z];
// gW2 [Z, Z] (any of the parameter outputs may be NULL: skipped).  9 launches.
int is_contrastive_bwd(const float* Ec, const float* Ew, int B, int D, int Z, const float* W1, const float* gamma,
                       const float* beta, const float* W2, const float* scratch, const float* gout, float* work,
                       float* gEc, float* gEw, float* gW1, float* g_gamma, float* g_beta, float* gW2, void* stream) {
    if (B < 2 || D <= 0 || Z <= 0) return IS_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    CtrLayout L = ctr_layout(const_cast<float*>(scratch), B, Z);
    const int64_t BZ = (int64_t)B * Z;
    const Term none = {nullptr, 0, 0, nullptr, 0, 0, 0, 0.0f};
    BnView bn = {L.mu, L.rstd, gamma, beta, Z, B};
    float* G = work;                                   // [2][B][Z]: G -> gz -> (B3 writes gpre into a second buffer)
    const float* zc_c = L.zc;
    const float* zc_w = L.zc + BZ;
    {   // G_c = dS zc_w + zc_w dC^T
        Term a = {L.dS, B, 1, zc_w, Z, 1, B, 1.0f};
        Term b = {zc_w, Z, 1, L.dC, 1, Z, Z, 1.0f};
        Epi ep = {G, Z, 0, nullptr, 0.f, 0.f, nullptr, nullptr, 0};
        launch_gemm<XF_NONE, EPI_STORE>(a, b, ep, bn, B, Z, 1, 0, 0, 0, st);
    }
    {   // G_w = dS^T zc_c + zc_c dC
        Term a = {L.dS, 1, B, zc_c, Z, 1, B, 1.0f};
        Term b = {zc_c, Z, 1, L.dC, Z, 1, Z, 1.0f};
        Epi ep = {G + BZ, Z, 0, nullptr, 0.f, 0.f, nullptr, nullptr, 0};
        launch_gemm<XF_NONE, EPI_STORE>(a, b, ep, bn, B, Z, 1, 0, 0, 0, st);
    }
    const int tz = (Z + 31) / 32;
    center_bwd_kernel<<<tz, 256, 0, st>>>(G, L.zc, L.stdv, B, Z, L.scal, gout);      // G = gz
    if (gW2 != nullptr) {   // gW2[n][k] = sum_(s,b) gz[(s,b)][n] a[(s,b)][k],  a = relu(bn(y)) recomputed on load
        Term t = {G, 1, Z, L.y, Z, 1, 2 * B, 1.0f};
        Epi ep = {gW2, Z, 0, nullptr, 0.f, 0.f, nullptr, nullptr, 0};
        launch_gemm<XF_B_BNRELU, EPI_STORE>(t, none, ep, bn, Z, Z, 1, 0, 0, 0, st);
    }
    // gpre = (gz W2) [pre > 0]  -> the zc half of the scratch is dead now, but it is const here: use dS? no -- B x B may be
    // smaller than 2 B Z.  gpre overwrites G in place is impossible (G is the A operand), so the caller's work buffer
    // holds a second [2][B][Z] block.
    float* gpre = work + 2 * BZ;
    {
        Term t = {G, Z, 1, W2, Z, 1, Z, 1.0f};
        Epi ep = {gpre, Z, 0, nullptr, 0.f, 0.f, nullptr, L.y, 0};
        launch_gemm<XF_NONE, EPI_RELU_MASK>(t, none, ep, bn, B, Z, 2, BZ, 0, BZ, st);
    }
    // g_gamma / g_beta are needed by nobody when NULL, but the kernel writes them: point at dead work space then
    float* gg = g_gamma ? g_gamma : G, * gb = g_beta ? g_beta : G + Z;
    bn_bwd_kernel<<<tz, 256, 0, st>>>(gpre, L.y, L.mu, L.rstd, gamma, B, Z, gg, gb);  // gpre = gy
    for (int s = 0; s < 2; ++s) {   // gE_s = gy_s W1
        Term t = {gpre + s * BZ, Z, 1, W1, D, 1, Z, 1.0f};
        Epi ep = {s == 0 ? gEc : gEw, D, 0, nullptr, 0.f, 0.f, nullptr, nullptr, 0};
        launch_gemm<XF_NONE, EPI_STORE>(t, none, ep, bn, B, D, 1, 0, 0, 0, st);
    }
    if (gW1 != nullptr) {   // gW1[n][d] = sum_b gy_c[b][n] Ec[b][d] + sum_b gy_w[b][n] Ew[b][d]
        Term a = {gpre, 1, Z, Ec, D, 1, B, 1.0f};
        Term b = {gpre + BZ, 1, Z, Ew, D, 1, B, 1.0f};
        Epi ep = {gW1, D, 0, nullptr, 0.f, 0.f, nullptr, nullptr, 0};
        launch_gemm<XF_NONE, EPI_STORE>(a, b, ep, bn, Z, D, 1, 0, 0, 0, st);
    }
    IS_LAUNCH_CHECK();
    return IS_OK;
}

}  // extern "C"
