// Per-graph single-head attention + global mean pool on the tcgen05 tensor cores (inference, pooled rows only).
//
// Reference: MultiHeadAttention / SelfAttention applied per graph followed by global_mean_pool
// (models/layers.py:13-22, 29-48, 67-78; models/hybrid_models.py:92-97 / 326-331):
//     S = Q K^T / sqrt(64),  P = softmax_rows(S),  pooled = mean_i (P V)_i = sum_j u_j V_j,  u_j = (1/n) sum_i P_ij
// (the mean over query rows commutes with the product by V, so only the COLUMN SUMS of P are needed: no P V GEMM,
// nothing of size n x n leaves the SM).
//
// One CTA walks graphs.  Per graph: the Q rows (pre-multiplied by log2(e)/sqrt(64)) and K rows are split on the
// fly into bf16 x 3 (fp32-accurate) or rounded to bf16 and staged as canonical K-major operand tiles; S = Q K^T runs
// as 24 (bf16: 4) MMAs per 128-row tile into TMEM ([<= 256 rows] x [NPAD columns] fp32 = up to 512 columns); each
// thread then owns one score row in TMEM: three passes of tcgen05.ld (row maximum, row sum of exp2, normalised
// exp2) with the column sums taken across the warp's 32 rows by a 16-shuffle reduce-scatter per 16 columns; the
// eight warps' partial column sums are added in warp order (deterministic) and contracted with V straight from
// global memory.  The SIMT kernel this replaces (attn_pool.cu) spent 189 us per 512-graph batch on the fp32 Q K^T.
#include "tc_common.cuh"

namespace is {
namespace atc {
constexpr int NT = 256;
constexpr uint32_t LBO = 128, SBO = 8 * LBO;           // unpadded canonical tiles
constexpr uint32_t QT_BYTES = 16 * SBO;                // 128 query rows, one split term

// sum over the warp's 32 rows of 16 per-lane column values; afterwards lane L holds the total of column (L >> 1) & 15
__device__ __forceinline__ float colsum16(const float (&v)[16], int lane) {
    float w8[8], w4[4], w2[2];
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float send = b4 ? v[i] : v[i + 8], keep = b4 ? v[i + 8] : v[i];
        w8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float send = b3 ? w8[i] : w8[i + 4], keep = b3 ? w8[i + 4] : w8[i];
        w4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float send = b2 ? w4[i] : w4[i + 2], keep = b2 ? w4[i + 2] : w4[i];
        w2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    const float send = b1 ? w2[0] : w2[1], keep = b1 ? w2[1] : w2[0];
    float s = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    return s;
}
__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
}  // namespace atc

// NPAD = 128 or 256: padded node count per graph (key rows / score columns and, in 128-row tiles, query rows)
template <int PREC, int NPAD>
__global__ void __launch_bounds__(atc::NT, 1)
attn_pool_tc_kernel(const float* __restrict__ QKV, const int64_t* __restrict__ node_off, int n_graphs,
                    float* __restrict__ pooled) {
    using namespace atc;
    using C = TcCfg<PREC>;
    constexpr int NS = C::NSPLIT;
    constexpr int MT = NPAD / 128;                              // 128-row query tiles
    constexpr uint32_t KT_BYTES = (NPAD / 8) * SBO;             // key tile, one split term
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint8_t* sQ = smem_raw;                                     // [MT][NS][QT_BYTES]
    uint8_t* sK = sQ + MT * NS * QT_BYTES;                      // [NS][KT_BYTES]
    float* part = reinterpret_cast<float*>(sK + NS * KT_BYTES); // [8 warps][NPAD] partial column sums
    float* u = part + 8 * NPAD;                                 // [NPAD]
    float* red = u + NPAD;                                      // [4][64]
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) tmem_alloc(&s_tmem, MT * NPAD);
    if (tid == 32) mbar_init(&mbar, 1);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = s_tmem;
    const uint32_t q_addr = smem_u32(sQ), k_addr = smem_u32(sK);
    const int r4 = lane & 3, kc = lane >> 2;                    // conflict-free staging map (see egnn_tc2.cu)
    const int q = warp & 3, mt = warp >> 2;                     // softmax: TMEM lane quarter, query tile (mt < MT)
    const float qscale = 1.4426950408889634f * 0.125f;          // log2(e) / sqrt(64)
    uint32_t phase = 0;

    for (int g = blockIdx.x; g < n_graphs; g += gridDim.x) {
        const int64_t n0 = __ldg(node_off + g);
        const int n = min((int)(__ldg(node_off + g + 1) - n0), NPAD);      // clamped: see GraphBatch validation
        // ---- stage Q (scaled) and K rows; rows >= n are zero (zero scores, masked below) ----------------------
        for (int grp = warp; grp < NPAD / 8; grp += NT / 32) {
            float qv[2][8], kv[2][8];
#pragma unroll
            for (int p = 0; p < 2; ++p) {
                const int row = 8 * grp + r4 + 4 * ((kc & 1) ^ p);
#pragma unroll
                for (int i = 0; i < 8; ++i) qv[p][i] = kv[p][i] = 0.0f;
                if (row < n) {
                    const float* src = QKV + (n0 + row) * 192 + 8 * kc;      // 768-byte rows: every chunk is one 32-byte sector
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
                const float (&vk)[8] = kv[p];
                const int rq = row & 127;
                store_chunk8<PREC>(sQ + (row >> 7) * NS * QT_BYTES + (rq >> 3) * SBO + (rq & 7) * 16 + kc * LBO, QT_BYTES, vq);
                store_chunk8<PREC>(sK + (row >> 3) * SBO + (row & 7) * 16 + kc * LBO, KT_BYTES, vk);
            }
        }
        fence_async_smem();
        fence_before_sync();
        __syncthreads();
        // ---- S = Q K^T: one accumulator of NPAD columns per 128-row query tile --------------------------------
        if (warp == 0) {
            if (elect_one()) {
                fence_after_sync();
                const uint32_t idesc = make_instr_desc(C::FMT, 128, NPAD);
                constexpr int NTERM = PREC == PREC_BF16X3 ? 6 : PREC == PREC_FP16X2 ? 3 : 1;
                // (q-term, k-term), smallest products first -- bf16x3: six of the nine; fp16x2: lo hi, hi lo, hi hi
                const uint32_t ta[6] = {PREC == PREC_FP16X2 ? 1u : 2u, 0, PREC == PREC_FP16X2 ? 0u : 1u, 1, 0, 0};
                const uint32_t tw[6] = {0, PREC == PREC_FP16X2 ? 1u : 2u, PREC == PREC_FP16X2 ? 0u : 1u, 0, 1, 0};
#pragma unroll
                for (int m = 0; m < MT; ++m) {
                    uint32_t acc = 0;
#pragma unroll
                    for (int t = 0; t < NTERM; ++t)
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            const uint32_t at = NTERM > 1 ? ta[t] : 0, wt = NTERM > 1 ? tw[t] : 0;
                            mma_bf16(tmem + m * NPAD,
                                     make_smem_desc(q_addr + (m * NS + at) * QT_BYTES + ks * 2 * LBO, LBO, SBO),
                                     make_smem_desc(k_addr + wt * KT_BYTES + ks * 2 * LBO, LBO, SBO), idesc, acc);
                            acc = 1;
                        }
                }
                mma_commit(&mbar);
            }
            __syncwarp();
        }
        mbar_wait(&mbar, phase);
        phase ^= 1;
        fence_after_sync();
        // ---- softmax statistics and column sums: thread = one score row in TMEM ---------------------------------
        const int row = 128 * mt + 32 * q + lane;
        const bool active_warp = mt < MT && 128 * mt + 32 * q < n;           // warp has at least one valid row
        const bool valid = mt < MT && row < n;
        const uint32_t t_row = tmem + ((uint32_t)(32 * q) << 16) + (mt < MT ? mt : 0) * NPAD;
        const int nch = (n + 15) >> 4;
        if (active_warp) {
            float mx = -INFINITY, l = 0.0f;
            for (int c = 0; c < nch; ++c) {
                float z[16];
                tmem_ld<16>(t_row + 16 * c, z);
#pragma unroll
                for (int i = 0; i < 16; ++i) mx = (16 * c + i < n) ? fmaxf(mx, z[i]) : mx;
            }
            for (int c = 0; c < nch; ++c) {
                float z[16];
                tmem_ld<16>(t_row + 16 * c, z);
#pragma unroll
                for (int i = 0; i < 16; ++i) l += (16 * c + i < n) ? ex2(z[i] - mx) : 0.0f;
            }
            const float w = valid ? 1.0f / (l * (float)n) : 0.0f;
            for (int c = 0; c < nch; ++c) {
                float z[16], e[16];
                tmem_ld<16>(t_row + 16 * c, z);
#pragma unroll
                for (int i = 0; i < 16; ++i) e[i] = (valid && 16 * c + i < n) ? ex2(z[i] - mx) * w : 0.0f;
                const float s = colsum16(e, lane);
                if (!(lane & 1)) part[warp * NPAD + 16 * c + (lane >> 1)] = s;
            }
        } else {
            for (int j = lane; j < 16 * nch; j += 32) part[warp * NPAD + j] = 0.0f;
        }
        fence_before_sync();
        __syncthreads();
        for (int j = tid; j < n; j += NT) {
            float s = 0.0f;
#pragma unroll
            for (int w8 = 0; w8 < 8; ++w8) s += part[w8 * NPAD + j];
            u[j] = s;
        }
        __syncthreads();
        // ---- pooled[c] = sum_j u_j V[j][c]: four j-quarters x 64 columns ----------------------------------------
        {
            const int c = tid & 63, jq = tid >> 6;
            const int jb = (n * jq) >> 2, je = (n * (jq + 1)) >> 2;
            float s = 0.0f;
            for (int j = jb; j < je; ++j) s = fmaf(u[j], __ldg(QKV + (n0 + j) * 192 + 128 + c), s);
            red[jq * 64 + c] = s;
        }
        __syncthreads();
        if (tid < 64) pooled[(int64_t)g * 64 + tid] = (red[tid] + red[64 + tid]) + (red[128 + tid] + red[192 + tid]);
        // the next graph's staging / MMAs overwrite smem and TMEM: every read above is complete (barriers, wait::ld)
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, MT * NPAD);
}

template <int PREC, int NPAD>
static int launch_attn_tc(const float* QKV, const int64_t* node_off, int n_graphs, float* pooled, cudaStream_t st) {
    using namespace atc;
    constexpr int NS = TcCfg<PREC>::NSPLIT;
    const size_t smem = (size_t)(NPAD / 128) * NS * QT_BYTES + (size_t)NS * (NPAD / 8) * SBO + sizeof(float) * (8 * NPAD + NPAD + 256);
    cudaError_t e = cudaFuncSetAttribute(attn_pool_tc_kernel<PREC, NPAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const int sms = current_num_sms();
    const int grid = n_graphs < sms ? n_graphs : sms;
    attn_pool_tc_kernel<PREC, NPAD><<<grid, NT, smem, st>>>(QKV, node_off, n_graphs, pooled);
    e = cudaGetLastError();
    return e == cudaSuccess ? 0 : (int)e;
}

}  // namespace is

using namespace is;

extern "C" {

// Single-head variant of is_attn_pool_infer on the tensor cores: QKV [N_total, 192] (Q | K | V, 64 wide each),
// node_off [n_graphs + 1] -> pooled [n_graphs, 64].  max_nodes <= 256.  precision 0 = bf16, 3 = bf16x3, 4 = fp16x2 (both fp32-accurate).
int is_attn_pool_infer_tc(const float* QKV, const int64_t* node_off, int n_graphs, int max_nodes, int precision,
                          float* pooled, void* stream) {
    if (n_graphs <= 0 || max_nodes <= 0) return IS_ERR_ARG;
    if (max_nodes > 256) return IS_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(QKV) & 31) != 0) return IS_ERR_ARG;          // 256-bit row loads
    cudaStream_t st = (cudaStream_t)stream;
    if (precision == PREC_BF16)
        return max_nodes <= 128 ? launch_attn_tc<PREC_BF16, 128>(QKV, node_off, n_graphs, pooled, st)
                                : launch_attn_tc<PREC_BF16, 256>(QKV, node_off, n_graphs, pooled, st);
    if (precision == PREC_BF16X3)
        return max_nodes <= 128 ? launch_attn_tc<PREC_BF16X3, 128>(QKV, node_off, n_graphs, pooled, st)
                                : launch_attn_tc<PREC_BF16X3, 256>(QKV, node_off, n_graphs, pooled, st);
    if (precision == PREC_FP16X2)          // fp16 hi / lo pairs for the scores (no-grad forward)
        return max_nodes <= 128 ? launch_attn_tc<PREC_FP16X2, 128>(QKV, node_off, n_graphs, pooled, st)
                                : launch_attn_tc<PREC_FP16X2, 256>(QKV, node_off, n_graphs, pooled, st);
    return IS_ERR_ARG;
}

}  // extern "C"
