// EGNN edge forward on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
// Same contract as is::edge_fwd_kernel (egnn.cu): per tile of <= 128 in-edges (node-aligned, CSR
// order)   t1 = silu(P[src] + Q[dst] + w_r r + w_a a)  ->  m = silu(t1 W2^T + b2)  ->
// c = w4 . silu(m W3^T + b3)  ->  warp per destination node:  hn = sum m,  x' = x + mean(c dhat).
// The two 128x64x64 per-tile GEMMs run as tcgen05.mma (M = 128, N = 64) with the accumulators in
// TMEM; everything else (gather, SiLU, aggregation) stays on the SIMT pipes of the same CTA:
//
//   gather (all warps)  -> A operand tile in smem (canonical K-major layout, umma.cuh)
//   MMA 1 (one thread)  -> TMEM cols [0,64)      -> epilogue 1 (all warps: tcgen05.ld, +b2, SiLU)
//                                                   writes m as the next A operand and as fp32 rows
//   MMA 2 (one thread)  -> TMEM cols [64,128)    -> meanwhile: hn aggregation from the fp32 rows
//                                                -> epilogue 2 (tcgen05.ld, +b3, SiLU, . w4)
//   coordinate aggregation (warp per destination node)
//
// Two precisions (template PREC):
//   PREC_BF16   : bf16 operands, fp32 accumulate, fast SiLU        (bf16 mode, tolerance 1e-2)
//   PREC_TF32X3 : 3xTF32 split  a = hi + lo, D = lo*Bhi + hi*Blo + hi*Bhi, accurate SiLU
//                 (fp32 parity: the self-test measures 2.6e-7 relative error for the GEMM)
// Two CTAs per SM share the tensor core: while one waits on its MMA mbarrier the other runs its
// SIMT phases.
#include "common.cuh"
#include "egnn_common.cuh"
#include "umma.cuh"

#define PREC_BF16 0
#define PREC_TF32X3 2

namespace is {

using namespace umma;

template <int PREC>
struct TcCfg {
    static constexpr int EB = PREC == PREC_BF16 ? 2 : 4;              // operand element bytes
    static constexpr int KCH = 64 * EB / 16;                          // 16-byte chunks per 64-wide row
    static constexpr uint32_t SBO = KCH * kLBO;                       // bytes between 8-row groups
    static constexpr uint32_t A_BYTES = 16 * SBO;                     // 128-row operand tile
    static constexpr uint32_t W_BYTES = 8 * SBO;                      // 64-row operand tile
    static constexpr int NSPLIT = PREC == PREC_BF16 ? 1 : 2;          // hi (+ lo)
    static constexpr uint32_t FMT = PREC == PREC_BF16 ? 1u : 2u;
};

template <int PREC>
__device__ __forceinline__ float act(float z) {
    if (PREC == PREC_BF16) return z * __fdividef(1.0f, 1.0f + __expf(-z));
    return silu(z);
}

// store 8 consecutive K values (features 8*kc8 .. 8*kc8+7) of operand row `row`
template <int PREC>
__device__ __forceinline__ void store_operand8(uint8_t* __restrict__ tile, int row, int kc8, const float (&v)[8]) {
    using C = TcCfg<PREC>;
    const uint32_t rbase = (uint32_t)((row >> 3) * C::SBO + (row & 7) * 16);
    if (PREC == PREC_BF16) {
        uint4 q;
        q.x = pack_bf16x2(v[0], v[1]); q.y = pack_bf16x2(v[2], v[3]);
        q.z = pack_bf16x2(v[4], v[5]); q.w = pack_bf16x2(v[6], v[7]);
        *reinterpret_cast<uint4*>(tile + rbase + kc8 * kLBO) = q;
    } else {
        float hi[8], lo[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { hi[i] = tf32_round(v[i]); lo[i] = tf32_round(v[i] - hi[i]); }
        uint8_t* p0 = tile + rbase + (2 * kc8) * kLBO;
        *reinterpret_cast<float4*>(p0) = make_float4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<float4*>(p0 + kLBO) = make_float4(hi[4], hi[5], hi[6], hi[7]);
        *reinterpret_cast<float4*>(p0 + C::A_BYTES) = make_float4(lo[0], lo[1], lo[2], lo[3]);
        *reinterpret_cast<float4*>(p0 + C::A_BYTES + kLBO) = make_float4(lo[4], lo[5], lo[6], lo[7]);
    }
}

// stage a torch [64 out, 64 in] weight as the B operand (row n = output feature, K = input feature)
template <int PREC>
__device__ __forceinline__ void stage_weight(uint8_t* __restrict__ tile, const float* __restrict__ W, int tid) {
    using C = TcCfg<PREC>;
    for (int idx = tid; idx < 64 * 64; idx += IS_THREADS) {
        const int n = idx >> 6, k = idx & 63;
        const float w = __ldg(W + idx);
        const uint32_t off = canon_off<C::EB>(n, k, C::KCH);
        if (PREC == PREC_BF16) {
            *reinterpret_cast<__nv_bfloat16*>(tile + off) = __float2bfloat16_rn(w);
        } else {
            const float hi = tf32_round(w);
            *reinterpret_cast<float*>(tile + off) = hi;
            *reinterpret_cast<float*>(tile + C::W_BYTES + off) = tf32_round(w - hi);
        }
    }
}

// issue one 128x64x64 GEMM: D[tmem_d] = A_tile * W_tile^T   (called by ONE thread)
template <int PREC>
__device__ __forceinline__ void issue_gemm(uint32_t tmem_d, uint32_t a_addr, uint32_t w_addr) {
    using C = TcCfg<PREC>;
    const uint32_t idesc = make_instr_desc(C::FMT, 128, 64);
    uint32_t acc = 0;
#pragma unroll
    for (int term = (PREC == PREC_BF16 ? 2 : 0); term < 3; ++term) {        // lo*hi, hi*lo, hi*hi
        const uint32_t aoff = (term == 0) ? C::A_BYTES : 0, woff = (term == 1) ? C::W_BYTES : 0;
#pragma unroll
        for (int ks = 0; ks < C::KCH / 2; ++ks) {
            const uint64_t da = make_smem_desc(a_addr + aoff + ks * 2 * kLBO, kLBO, C::SBO);
            const uint64_t db = make_smem_desc(w_addr + woff + ks * 2 * kLBO, kLBO, C::SBO);
            if (PREC == PREC_BF16) mma_bf16(tmem_d, da, db, idesc, acc); else mma_tf32(tmem_d, da, db, idesc, acc);
            acc = 1;
        }
    }
}

template <int PREC, bool HAS_COORD>
__global__ void __launch_bounds__(IS_THREADS, PREC == PREC_BF16 ? 2 : 1)
edge_fwd_tc_kernel(EdgeCommon p, float* __restrict__ hn, float* __restrict__ x_out) {
    using C = TcCfg<PREC>;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint8_t* sA = smem_raw;                                            // [NSPLIT][A_BYTES] t1, then m
    uint8_t* sW2 = sA + C::NSPLIT * C::A_BYTES;                        // [NSPLIT][W_BYTES]
    uint8_t* sW3 = sW2 + C::NSPLIT * C::W_BYTES;
    float* M32 = reinterpret_cast<float*>(sW3 + C::NSPLIT * C::W_BYTES);   // [128][68] fp32 m
    float* vec = M32 + IS_TM * IS_LD;                                  // b2, b3, w4, wr, wa
    float* e_c = vec + 5 * 64;                                         // [2][128] partial c per column half
    float* e_dh = e_c + 2 * IS_TM;                                     // [128][3]
    __shared__ int s_tile[4];
    __shared__ __align__(8) uint64_t mbar[2];
    __shared__ uint32_t s_tmem;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ldw1 = 2 * p.F + 2;
    if (warp == 0) tmem_alloc(&s_tmem, 128);
    if (tid == 32) { mbar_init(&mbar[0], 1); mbar_init(&mbar[1], 1); }
    stage_weight<PREC>(sW2, p.W2, tid);
    if (HAS_COORD) stage_weight<PREC>(sW3, p.W3, tid);
    if (tid < 64) {
        vec[tid] = p.b2[tid];
        vec[64 + tid] = HAS_COORD ? p.b3[tid] : 0.0f;
        vec[128 + tid] = HAS_COORD ? p.w4[tid] : 0.0f;
        vec[192 + tid] = p.W1[tid * ldw1 + 2 * p.F];
        vec[256 + tid] = p.W1[tid * ldw1 + 2 * p.F + 1];
    }
    const int chunk = (p.n_nodes + gridDim.x - 1) / gridDim.x;
    int n0 = blockIdx.x * chunk;
    const int nend = min(p.n_nodes, n0 + chunk);
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = s_tmem;
    const uint32_t a_addr = smem_u32(sA), w2_addr = smem_u32(sW2), w3_addr = smem_u32(sW3);
    // epilogue mapping: TMEM lane quarter q = warp % 4 (rows 32q + lane), column half ch = warp / 4
    const int q = warp & 3, ch = warp >> 2, erow = 32 * q + lane;
    const uint32_t t_lane = tmem + ((uint32_t)(32 * q) << 16);
    // gather mapping: 4 edges per warp per pass, 8 lanes per edge, 8 features per lane
    const int esub = lane >> 3, kc8 = lane & 7;
    uint32_t phase = 0;

    while (n0 < nend) {
        select_tile(s_tile, p.indptr, n0, nend, p.status);
        __syncthreads();
        const int n1 = s_tile[1], p0 = s_tile[2], ne = s_tile[3];
        if (ne < 0) { n0 = n1; __syncthreads(); continue; }

        // ---- gather -> A operand (t1) -----------------------------------------------------------
        {
            const float4 wr0 = *reinterpret_cast<const float4*>(vec + 192 + 8 * kc8), wr1 = *reinterpret_cast<const float4*>(vec + 196 + 8 * kc8);
            const float4 wa0 = *reinterpret_cast<const float4*>(vec + 256 + 8 * kc8), wa1 = *reinterpret_cast<const float4*>(vec + 260 + 8 * kc8);
#pragma unroll
            for (int pass = 0; pass < IS_TM / 32; ++pass) {
                const int j = pass * 32 + warp * 4 + esub;
                float v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = 0.0f;
                if (j < ne) {
                    const int e = p0 + j;
                    const int s = __ldg(p.csr_src + e), d = __ldg(p.csr_dst + e);
                    const float a = __ldg(p.edge_attr + __ldg(p.csr_eid + e));
                    const float dx = __ldg(p.x + s * p.ldx + 0) - __ldg(p.x + d * p.ldx + 0);
                    const float dy = __ldg(p.x + s * p.ldx + 1) - __ldg(p.x + d * p.ldx + 1);
                    const float dz = __ldg(p.x + s * p.ldx + 2) - __ldg(p.x + d * p.ldx + 2);
                    const float r = dx * dx + dy * dy + dz * dz;
                    const float4* pp = reinterpret_cast<const float4*>(p.PQ + (size_t)s * 128 + 8 * kc8);
                    const float4* qp = reinterpret_cast<const float4*>(p.PQ + (size_t)d * 128 + 64 + 8 * kc8);
                    const float4 p0v = __ldg(pp), p1v = __ldg(pp + 1), q0v = __ldg(qp), q1v = __ldg(qp + 1);
                    v[0] = act<PREC>(p0v.x + q0v.x + wr0.x * r + wa0.x * a);
                    v[1] = act<PREC>(p0v.y + q0v.y + wr0.y * r + wa0.y * a);
                    v[2] = act<PREC>(p0v.z + q0v.z + wr0.z * r + wa0.z * a);
                    v[3] = act<PREC>(p0v.w + q0v.w + wr0.w * r + wa0.w * a);
                    v[4] = act<PREC>(p1v.x + q1v.x + wr1.x * r + wa1.x * a);
                    v[5] = act<PREC>(p1v.y + q1v.y + wr1.y * r + wa1.y * a);
                    v[6] = act<PREC>(p1v.z + q1v.z + wr1.z * r + wa1.z * a);
                    v[7] = act<PREC>(p1v.w + q1v.w + wr1.w * r + wa1.w * a);
                    if (HAS_COORD && kc8 == 0) {
                        const float inv = 1.0f / (sqrtf(r) + 1e-30f);
                        e_dh[j * 3 + 0] = dx * inv; e_dh[j * 3 + 1] = dy * inv; e_dh[j * 3 + 2] = dz * inv;
                    }
                }
                store_operand8<PREC>(sA, j, kc8, v);
            }
        }
        fence_async_smem();
        fence_before_sync();
        __syncthreads();
        if (tid == 0) {
            fence_after_sync();
            issue_gemm<PREC>(tmem, a_addr, w2_addr);
            mma_commit(&mbar[0]);
        }

        // ---- epilogue 1: m = silu(acc0 + b2) -> fp32 rows (+ next A operand) ----------------------
        mbar_wait(&mbar[0], phase);
        fence_after_sync();
        {
            float z[32];
            tmem_ld32(t_lane + 32 * ch, z);
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                float m8[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) m8[i] = act<PREC>(z[8 * g + i] + vec[32 * ch + 8 * g + i]);
                float* dst = M32 + erow * IS_LD + 32 * ch + 8 * g;
                *reinterpret_cast<float4*>(dst) = make_float4(m8[0], m8[1], m8[2], m8[3]);
                *reinterpret_cast<float4*>(dst + 4) = make_float4(m8[4], m8[5], m8[6], m8[7]);
                if (HAS_COORD) store_operand8<PREC>(sA, erow, 4 * ch + g, m8);     // MMA 1 has finished reading sA
            }
        }
        fence_async_smem();
        fence_before_sync();
        __syncthreads();
        if (HAS_COORD && tid == 0) {
            fence_after_sync();
            issue_gemm<PREC>(tmem + 64, a_addr, w3_addr);
            mma_commit(&mbar[1]);
        }

        // ---- hn aggregation (overlaps MMA 2): one warp per destination node -----------------------
        for (int node = n0 + warp; node < n1; node += IS_THREADS / 32) {
            const int jb = __ldg(p.indptr + node) - p0, je = __ldg(p.indptr + node + 1) - p0;
            float2 s = make_float2(0.0f, 0.0f);
            for (int j = jb; j < je; ++j) {
                const float2 v = *reinterpret_cast<const float2*>(M32 + j * IS_LD + 2 * lane);
                s.x += v.x; s.y += v.y;
            }
            *reinterpret_cast<float2*>(hn + (size_t)node * 64 + 2 * lane) = s;
        }

        if (HAS_COORD) {
            // ---- epilogue 2: c = w4 . silu(acc1 + b3) (each thread: one row, 32 columns) ----------
            mbar_wait(&mbar[1], phase);
            fence_after_sync();
            float z[32];
            tmem_ld32(t_lane + 64 + 32 * ch, z);
            float c = 0.0f;
#pragma unroll
            for (int i = 0; i < 32; ++i) c = fmaf(vec[128 + 32 * ch + i], act<PREC>(z[i] + vec[64 + 32 * ch + i]), c);
            e_c[ch * IS_TM + erow] = c;
            fence_before_sync();
            __syncthreads();
            for (int node = n0 + warp; node < n1; node += IS_THREADS / 32) {
                if (lane < 3) {
                    const int jb = __ldg(p.indptr + node) - p0, je = __ldg(p.indptr + node + 1) - p0;
                    float sx = 0.0f;
                    for (int j = jb; j < je; ++j) sx += (e_c[j] + e_c[IS_TM + j]) * e_dh[j * 3 + lane];
                    const int deg = je - jb;
                    x_out[(size_t)node * 3 + lane] = __ldg(p.x + node * p.ldx + lane) + sx / (float)max(deg, 1);
                }
            }
        }
        phase ^= 1;
        n0 = n1;
        fence_before_sync();
        __syncthreads();
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 128);
}

template <int PREC>
static size_t tc_smem_bytes() {
    using C = TcCfg<PREC>;
    return (size_t)C::NSPLIT * (C::A_BYTES + 2 * C::W_BYTES) + sizeof(float) * (IS_TM * IS_LD + 5 * 64 + 2 * IS_TM + 3 * IS_TM) + 128;
}

template <int PREC, bool HAS_COORD>
static int launch_tc(const EdgeCommon& c, float* hn, float* x_out, int grid, cudaStream_t st) {
    const size_t smem = tc_smem_bytes<PREC>();
    cudaError_t e = cudaFuncSetAttribute(edge_fwd_tc_kernel<PREC, HAS_COORD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    edge_fwd_tc_kernel<PREC, HAS_COORD><<<grid, IS_THREADS, smem, st>>>(c, hn, x_out);
    e = cudaGetLastError();
    return e == cudaSuccess ? 0 : (int)e;
}

}  // namespace is

using namespace is;

extern "C" {

// Tensor-core variant of is_egnn_edge_fwd.  precision: 0 = bf16 operands, 2 = 3xTF32 (fp32-accurate).
int is_egnn_edge_fwd_tc(const int* indptr, const int* csr_src, const int* csr_dst, const int* csr_eid,
                        const float* PQ, const float* x, int64_t ldx, const float* edge_attr,
                        const float* W1, int F, const float* W2, const float* b2,
                        const float* W3, const float* b3, const float* w4, int update_coords, int precision,
                        float* hn, float* x_out, int64_t n_nodes, int* status, void* stream) {
    if (!(F == 20 || F == 64) || n_nodes <= 0 || n_nodes > 0x7fffffff) return IS_ERR_ARG;
    if (precision != PREC_BF16 && precision != PREC_TF32X3) return IS_ERR_ARG;
    EdgeCommon c;
    c.indptr = indptr; c.csr_src = csr_src; c.csr_dst = csr_dst; c.csr_eid = csr_eid;
    c.PQ = PQ; c.x = x; c.ldx = ldx; c.edge_attr = edge_attr; c.W1 = W1; c.F = F;
    c.W2 = W2; c.b2 = b2; c.W3 = W3; c.b3 = b3; c.w4 = w4; c.n_nodes = (int)n_nodes; c.status = status;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int per_sm = precision == PREC_BF16 ? 2 : 1;
    int64_t g = (n_nodes + 31) / 32;
    if (g > (int64_t)sms * per_sm) g = (int64_t)sms * per_sm;
    const int grid = (int)(g < 1 ? 1 : g);
    cudaStream_t st = (cudaStream_t)stream;
    if (precision == PREC_BF16)
        return update_coords ? launch_tc<PREC_BF16, true>(c, hn, x_out, grid, st) : launch_tc<PREC_BF16, false>(c, hn, x_out, grid, st);
    return update_coords ? launch_tc<PREC_TF32X3, true>(c, hn, x_out, grid, st) : launch_tc<PREC_TF32X3, false>(c, hn, x_out, grid, st);
}

}  // extern "C"
