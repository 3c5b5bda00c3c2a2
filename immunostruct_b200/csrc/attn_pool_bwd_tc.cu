// Backward of the per-graph single-head attention + global mean pool on the tcgen05 tensor cores (pooled rows only:
// every model's training path; forward = attn_pool_tc.cu, nothing but QKV is saved).
//
// Reference: MultiHeadAttention / SelfAttention per graph followed by global_mean_pool (models/layers.py:13-22, 29-48,
// 67-78; models/hybrid_models.py:92-97 / 326-331), differentiated by torch autograd there.  With
//     S = Q K^T / 8,  P = softmax_rows(S),  pooled = (1/n) sum_i (P V)_i,  g0 = g_pooled / n
// every row of gO equals g0, so dP_ij = V_j . g0 =: c_j does not depend on i and
//     D_i = sum_j P_ij c_j,   G_ij = P_ij (c_j - D_i),
//     gQ_i = (1/8) sum_j G_ij K_j,   gK_j = (1/8) sum_i G_ij Q_i,   gV_j = (sum_i P_ij) g0.
//
// One CTA walks graphs.  Q (pre-multiplied by log2(e)/8) and K are staged once per graph as bf16 x 3 canonical operand
// tiles.  Then four identical rounds -- {gQ, gK} x {128-row tiles}: the score tile T = X_t Y^T (X = Q, Y = K for gQ;
// X = K, Y = Q for gK, i.e. the TRANSPOSED score tile, so a thread again owns one row) goes to TMEM (24 MMAs); for gQ
// the four warps of the tile first take the row statistics (max, sum, D) straight from TMEM; then 32 columns at a time
// all eight warps turn T into G (exp2, one multiply), split it into three bf16 terms in a 128 x 32 operand chunk and an
// elected lane accumulates out += G_chunk Y_chunk (12 MMAs, B = the MN-major view of 32 rows of the Y tile).  In the gK
// rounds the thread's row sum of P is gV's weight.  Deterministic: fixed summation order everywhere.
// The SIMT kernel this replaces (attn_bwd_kernel, attn_pool.cu) spent 945 us per 512-graph batch on four fp32
// n x n x 64 products at 15 % of the FMA peak.
#include "tc_common.cuh"

namespace is {
namespace abt {
constexpr int NT = 256;
constexpr uint32_t LBO = 128, SBO = 8 * LBO;                             // Q / K tiles (unpadded canonical, K = features)
constexpr uint32_t G_LBO = 128, G_SBO = 4 * G_LBO, G_BYTES = 16 * G_SBO; // 128 x 32 chunk of G, one split term (8 KB)
constexpr uint32_t TM_T = 0, TM_O = 256;                                 // TMEM columns: score tile | output tile

struct Geom { uint32_t base, split, step, lbo, sbo; };  // operand view: start, split-term stride, K-step advance, LBO, SBO

// D (+)= A B^T in the bf16x3 split: six partial products, smallest first (ONE elected thread).  (The two further
// products a2 b3 + a3 b2 were tried: 201 -> 224 us, no measurable change in any gradient of the golden cases.)
__device__ __forceinline__ void issue6(uint32_t tmem_d, const Geom& a, const Geom& b, int nks, uint32_t idesc, uint32_t accumulate) {
    using namespace umma;
    const uint32_t ta[6] = {2, 0, 1, 1, 0, 0}, tb[6] = {0, 2, 1, 0, 1, 0};
    uint32_t acc = accumulate;
#pragma unroll
    for (int t = 0; t < 6; ++t)
#pragma unroll 4
        for (int ks = 0; ks < nks; ++ks) {
            mma_bf16(tmem_d, make_smem_desc(a.base + ta[t] * a.split + ks * a.step, a.lbo, a.sbo),
                     make_smem_desc(b.base + tb[t] * b.split + ks * b.step, b.lbo, b.sbo), idesc, acc);
            acc = 1;
        }
}
__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// balanced sum of 16 values: the row sums over <= 256 columns are taken as 16 chunk sums of this shape (a sequential
// fp32 sum of 200 terms carries ~1e-6 of rounding error, which scales a whole row of P and of gQ)
__device__ __forceinline__ float tree16(const float (&v)[16]) {
    float a[8], b[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = v[2 * i] + v[2 * i + 1];
#pragma unroll
    for (int i = 0; i < 4; ++i) b[i] = a[2 * i] + a[2 * i + 1];
    return (b[0] + b[1]) + (b[2] + b[3]);
}
}  // namespace abt

// NPAD = 128 or 256: padded node count per graph
template <int NPAD>
__global__ void __launch_bounds__(abt::NT, 1)
attn_pool_bwd_tc_kernel(const float* __restrict__ QKV, const int64_t* __restrict__ node_off, int n_graphs,
                        const float* __restrict__ g_pooled, float* __restrict__ gQKV) {
    using namespace abt;
    constexpr int MT = NPAD / 128;                              // 128-row tiles
    constexpr uint32_t KT_BYTES = (NPAD / 8) * SBO;             // Q or K tile, one split term
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint8_t* sQ = smem_raw;                                     // [3][KT_BYTES]  Q rows * log2(e)/8
    uint8_t* sK = sQ + 3 * KT_BYTES;                            // [3][KT_BYTES]
    uint8_t* sG = sK + 3 * KT_BYTES;                            // [3][G_BYTES]
    float* c_s = reinterpret_cast<float*>(sG + 3 * G_BYTES);    // [NPAD] c_j = V_j . g0
    float* mx_s = c_s + NPAD;                                   // [NPAD] maximum of score row i (log2 units)
    float* il_s = mx_s + NPAD;                                  // [NPAD] 1 / sum_j exp2(T_ij - mx_i)
    float* D_s = il_s + NPAD;                                   // [NPAD] D_i
    float* ps_s = D_s + NPAD;                                   // [2][128] row sums of P^T (two column halves)
    float* g0_s = ps_s + 256;                                   // [64]
    __shared__ __align__(8) uint64_t mbar_t, mbar_g;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) tmem_alloc(&s_tmem, 512);
    if (tid == 32) { mbar_init(&mbar_t, 1); mbar_init(&mbar_g, 1); }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = s_tmem;
    const uint32_t q_addr = smem_u32(sQ), k_addr = smem_u32(sK), g_addr = smem_u32(sG);
    const int r4 = lane & 3, kc = lane >> 2;                    // conflict-free staging map (see egnn_tc2.cu)
    const int q = warp & 3, half = warp >> 2;                   // TMEM lane quarter; column half of a 32-column chunk
    const int trow = 32 * q + lane;                             // this thread's row of the 128-row tile
    const uint32_t t_lane = tmem + ((uint32_t)(32 * q) << 16);
    const float qscale = 1.4426950408889634f * 0.125f;          // log2(e) / sqrt(64)
    const uint32_t id_t = umma::make_instr_desc(1u, 128, NPAD), id_o = umma::make_instr_desc(1u, 128, 64, 0, 1);
    uint32_t ph_t = 0, ph_g = 0;

    for (int g = blockIdx.x; g < n_graphs; g += gridDim.x) {
        const int64_t n0 = __ldg(node_off + g);
        const int n = min((int)(__ldg(node_off + g + 1) - n0), NPAD);      // clamped: see GraphBatch validation
        if (n <= 0) continue;
        if (tid < 64) g0_s[tid] = __ldg(g_pooled + (int64_t)g * 64 + tid) / (float)n;
        // ---- stage Q (scaled) and K rows; rows >= n are zero ---------------------------------------------------------
        for (int grp = warp; grp < NPAD / 8; grp += NT / 32) {
            float qv[2][8], kv[2][8];
#pragma unroll
            for (int p = 0; p < 2; ++p) {
                const int row = 8 * grp + r4 + 4 * ((kc & 1) ^ p);
#pragma unroll
                for (int i = 0; i < 8; ++i) qv[p][i] = kv[p][i] = 0.0f;
                if (row < n) {
                    const float* src = QKV + (n0 + row) * 192 + 8 * kc;
                    ldg256(src, qv[p]);
                    ldg256(src + 64, kv[p]);
                }
            }
#pragma unroll
            for (int p = 0; p < 2; ++p) {
                const int row = 8 * grp + r4 + 4 * ((kc & 1) ^ p);
                float vq[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) vq[i] = qv[p][i] * qscale;
                const uint32_t off = (row >> 3) * SBO + (row & 7) * 16 + kc * LBO;
                store_chunk8<PREC_BF16X3>(sQ + off, KT_BYTES, vq);
                store_chunk8<PREC_BF16X3>(sK + off, KT_BYTES, kv[p]);
            }
        }
        __syncthreads();                                        // g0_s
        for (int j = tid; j < NPAD; j += NT) {
            float c = 0.0f;
            if (j < n) {
                const float* vp = QKV + (n0 + j) * 192 + 128;
#pragma unroll
                for (int k8 = 0; k8 < 8; ++k8) {
                    float v[8];
                    ldg256(vp + 8 * k8, v);
#pragma unroll
                    for (int i = 0; i < 8; ++i) c = fmaf(g0_s[8 * k8 + i], v[i], c);
                }
            }
            c_s[j] = c;
        }
        fence_async_smem();
        fence_before_sync();
        __syncthreads();

        const int nch = (n + 31) >> 5;                          // 32-column chunks holding valid columns
        for (int pass = 0; pass < 2; ++pass) {                  // 0: gQ (rows i, X = Q, Y = K);  1: gK, gV (rows j, X = K, Y = Q)
            const uint32_t x_addr = pass == 0 ? q_addr : k_addr, y_addr = pass == 0 ? k_addr : q_addr;
            for (int t = 0; t < MT; ++t) {
                if (128 * t >= n) continue;
                // ---- T = X_t Y^T -> TMEM [128 rows] x [NPAD columns] -----------------------------------------------
                if (warp == 0) {
                    if (elect_one()) {
                        fence_after_sync();
                        issue6(tmem + TM_T, Geom{x_addr + t * 16 * SBO, KT_BYTES, 2 * LBO, LBO, SBO},
                               Geom{y_addr, KT_BYTES, 2 * LBO, LBO, SBO}, 4, id_t, 0);
                        mma_commit(&mbar_t);
                    }
                    __syncwarp();
                }
                mbar_wait(&mbar_t, ph_t);
                ph_t ^= 1;
                fence_after_sync();
                const int row = 128 * t + trow;
                const bool valid = row < n;
                if (pass == 0) {
                    // ---- row statistics (four warps, one score row per thread) -----------------------------------
                    if (half == 0) {
                        float mxr = 0.0f, il = 0.0f, D = 0.0f;
                        if (128 * t + 32 * q < n) {
                            const int nc16 = (n + 15) >> 4;
                            float mx = -INFINITY, l = 0.0f, ds = 0.0f;
                            for (int c = 0; c < nc16; ++c) {
                                float z[16];
                                tmem_ld<16>(t_lane + TM_T + 16 * c, z);
#pragma unroll
                                for (int i = 0; i < 16; ++i) mx = (16 * c + i < n) ? fmaxf(mx, z[i]) : mx;
                            }
                            for (int c = 0; c < nc16; ++c) {
                                float z[16];
                                tmem_ld<16>(t_lane + TM_T + 16 * c, z);
                                float e[16], ec[16];
#pragma unroll
                                for (int i = 0; i < 16; ++i) {
                                    e[i] = (16 * c + i < n) ? ex2(z[i] - mx) : 0.0f;
                                    ec[i] = e[i] * c_s[16 * c + i];
                                }
                                l += tree16(e);
                                ds += tree16(ec);
                            }
                            mxr = mx;
                            il = 1.0f / l;
                            D = ds * il;
                        }
                        mx_s[row] = mxr;
                        il_s[row] = il;
                        D_s[row] = D;
                    }
                    __syncthreads();
                }
                // P_ij = exp2(T_ij - mx_i) / l_i with the SAME exp2 values the statistics summed: sum_j G_ij cancels to
                // rounding, which matters because the K rows share a large common component (sum_j G_ij K_j)
                const float my_mx = mx_s[row], my_il = il_s[row], my_D = D_s[row], my_c = c_s[row];
                float psum = 0.0f;
                // ---- out += G_chunk Y_chunk, 32 columns at a time ---------------------------------------------------
                for (int c = 0; c < nch; ++c) {
                    const int j0 = 32 * c + 16 * half;
                    float z[16], gv[16], pv[16];
                    tmem_ld<16>(t_lane + TM_T + j0, z);
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int j = j0 + i;
                        const bool ok = valid && j < n;
                        if (pass == 0) {
                            const float p = ok ? ex2(z[i] - my_mx) * my_il : 0.0f;
                            gv[i] = ok ? p * (c_s[j] - my_D) : 0.0f;
                        } else {
                            const float p = ok ? ex2(z[i] - mx_s[j]) * il_s[j] : 0.0f;
                            pv[i] = p;
                            gv[i] = ok ? p * (my_c - D_s[j]) : 0.0f;
                        }
                    }
                    if (pass == 1) psum += tree16(pv);
                    if (c > 0) { mbar_wait(&mbar_g, ph_g); ph_g ^= 1; }      // the previous chunk's MMAs have read sG
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const float v8[8] = {gv[8 * u], gv[8 * u + 1], gv[8 * u + 2], gv[8 * u + 3],
                                             gv[8 * u + 4], gv[8 * u + 5], gv[8 * u + 6], gv[8 * u + 7]};
                        store_chunk8<PREC_BF16X3>(sG + (trow >> 3) * G_SBO + (trow & 7) * 16 + (2 * half + u) * G_LBO, G_BYTES, v8);
                    }
                    fence_async_smem();
                    fence_before_sync();
                    __syncthreads();
                    if (warp == 0) {
                        if (elect_one()) {
                            fence_after_sync();
                            issue6(tmem + TM_O, Geom{g_addr, G_BYTES, 2 * G_LBO, G_LBO, G_SBO},
                                   Geom{y_addr + 4 * c * SBO, KT_BYTES, 2 * SBO, SBO, LBO}, 2, id_o, c > 0 ? 1u : 0u);
                            mma_commit(&mbar_g);
                        }
                        __syncwarp();
                    }
                }
                mbar_wait(&mbar_g, ph_g);
                ph_g ^= 1;
                fence_after_sync();
                // ---- read the output tile: this thread's row, 32 of the 64 columns ---------------------------------
                {
                    float o[32];
                    tmem_ld<32>(t_lane + TM_O + 32 * half, o);
                    const float sc = pass == 0 ? 0.125f : 0.6931471805599453f;   // gK used the pre-scaled Q: 1/8 / (log2(e)/8)
                    if (valid) {
                        float* dst = gQKV + (n0 + row) * 192 + (pass == 0 ? 0 : 64) + 32 * half;
#pragma unroll
                        for (int k8 = 0; k8 < 4; ++k8) {
                            const float v8[8] = {o[8 * k8] * sc, o[8 * k8 + 1] * sc, o[8 * k8 + 2] * sc, o[8 * k8 + 3] * sc,
                                                 o[8 * k8 + 4] * sc, o[8 * k8 + 5] * sc, o[8 * k8 + 6] * sc, o[8 * k8 + 7] * sc};
                            stg256(dst + 8 * k8, v8);
                        }
                    }
                }
                if (pass == 1) ps_s[half * 128 + trow] = psum;
                fence_before_sync();
                __syncthreads();
                if (pass == 1 && valid) {                       // gV_j = (sum_i P_ij) g0: halves added in a fixed order
                    const float w = ps_s[trow] + ps_s[128 + trow];
                    float* dst = gQKV + (n0 + row) * 192 + 128 + 32 * half;
#pragma unroll
                    for (int k8 = 0; k8 < 4; ++k8) {
                        float v8[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) v8[i] = w * g0_s[32 * half + 8 * k8 + i];
                        stg256(dst + 8 * k8, v8);
                    }
                }
            }
        }
        __syncthreads();                                        // g0_s / ps_s / tiles are rewritten by the next graph
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

template <int NPAD>
static int launch_attn_bwd_tc(const float* QKV, const int64_t* node_off, int n_graphs, const float* g_pooled, float* gQKV,
                              cudaStream_t st) {
    using namespace abt;
    const size_t smem = 6 * (size_t)(NPAD / 8) * SBO + 3 * (size_t)G_BYTES + sizeof(float) * (4 * NPAD + 256 + 64);
    cudaError_t e = cudaFuncSetAttribute(attn_pool_bwd_tc_kernel<NPAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const int sms = current_num_sms();
    const int grid = n_graphs < sms ? n_graphs : sms;
    attn_pool_bwd_tc_kernel<NPAD><<<grid, NT, smem, st>>>(QKV, node_off, n_graphs, g_pooled, gQKV);
    e = cudaGetLastError();
    return e == cudaSuccess ? 0 : (int)e;
}

}  // namespace is

using namespace is;

extern "C" {

// Backward of is_attn_pool_infer_tc (single head, pooled rows only) in the bf16x3 (fp32-accurate) split:
// QKV [N_total, 192], node_off [n_graphs + 1], g_pooled [n_graphs, 64] -> gQKV [N_total, 192] (every row written).
// max_nodes <= 256.
int is_attn_pool_bwd_tc(const float* QKV, const int64_t* node_off, int n_graphs, int max_nodes, const float* g_pooled,
                        float* gQKV, void* stream) {
    if (n_graphs <= 0 || max_nodes <= 0) return IS_ERR_ARG;
    if (max_nodes > 256) return IS_ERR_UNSUPPORTED;
    if (((reinterpret_cast<uintptr_t>(QKV) | reinterpret_cast<uintptr_t>(gQKV)) & 31) != 0) return IS_ERR_ARG;   // 256-bit accesses
    cudaStream_t st = (cudaStream_t)stream;
    return max_nodes <= 128 ? launch_attn_bwd_tc<128>(QKV, node_off, n_graphs, g_pooled, gQKV, st)
                            : launch_attn_bwd_tc<256>(QKV, node_off, n_graphs, g_pooled, gQKV, st);
}

}  // extern "C"
