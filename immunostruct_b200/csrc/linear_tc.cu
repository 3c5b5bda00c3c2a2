// Dense Linear layer on the tcgen05 tensor cores:  C[M,N] = act(A[M,K] W[N,K]^T + b[N]),  fp32 in / out.
//
// Reference: the nn.Linear layers of the sequence VAE (models/hybrid_models.py:63-74: vae_fc1 5943 -> 512 with
// ReLU, vae_fc4 512 -> 5943) -- the only large dense contractions of the model (M = batch).  cuBLAS runs them as
// fp32 SIMT GEMMs (147 + 95 us at batch 512); here the operands are split on the fly into three bf16 terms
// (a = a1 + a2 + a3, six partial products, fp32 accumulate in TMEM: fp32-accurate) or rounded to bf16.
//
// One CTA = one 128 x 128 output tile over a K slice (split-K for the short-and-wide vae_fc1 so that 148 SMs have
// work; partial tiles go to a workspace and `linear_reduce_kernel` sums them in slice order: deterministic).
// Per 64-wide K block: all 16 warps load + split the A and W rows into a double-buffered pair of unpadded canonical
// operand tiles (conflict-free quarter-warp mapping of egnn_tc2.cu), one elected lane issues the 24 (bf16: 4) MMAs
// of the block; the loads of block k+1 are in flight while the MMAs of block k run.  Rows of A and W may have any
// stride (5943 floats: no 16-byte alignment) -- scalar loads unless the stride and base allow float4.
#include "tc_common.cuh"

namespace is {
namespace lin {
constexpr int NT = 512;
constexpr int BM = 128, BN = 128, BK = 64;
constexpr uint32_t LBO = 128, SBO = 8 * LBO, T_BYTES = 16 * SBO;      // one 128-row operand tile, one split term
}  // namespace lin

template <int PREC>
__global__ void __launch_bounds__(lin::NT, 1)
linear_tc_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ W, int64_t ldw,
                 const float* __restrict__ bias, float* __restrict__ C, int64_t ldc, float* __restrict__ part,
                 int64_t M, int64_t N, int64_t K, int relu, int split_k) {
    using namespace lin;
    using Cf = TcCfg<PREC>;
    constexpr int NS = Cf::NSPLIT;
    constexpr uint32_t OPB = NS * T_BYTES;                 // one operand, all split terms
    extern __shared__ __align__(128) uint8_t smem_raw[];   // [2 buffers][A | W][NS][T_BYTES]
    __shared__ __align__(8) uint64_t mbar[2];
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t m0 = (int64_t)blockIdx.y * BM, n0 = (int64_t)blockIdx.x * BN;
    const int nkb_all = (int)((K + BK - 1) / BK);
    const int kb0 = (int)((int64_t)nkb_all * blockIdx.z / split_k), kb1 = (int)((int64_t)nkb_all * (blockIdx.z + 1) / split_k);
    const int nkb = kb1 - kb0;

    if (warp == 0) tmem_alloc(&s_tmem, BN);
    if (tid == 32) { mbar_init(&mbar[0], 1); mbar_init(&mbar[1], 1); }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = s_tmem;
    const uint32_t base = smem_u32(smem_raw);
    const bool vec8 = ((lda | ldw) & 7) == 0 && ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(W)) & 31) == 0;
    const bool vec4 = ((lda | ldw) & 3) == 0 && ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(W)) & 15) == 0;
    // staging: warp w owns 8-row group w of both operands; a quarter-warp covers rows r4 + 4 * (chunk parity ^ pass
    // parity) for two adjacent 8-wide K chunks (kc = lane / 4): its eight 16-byte stores fill one 128-byte bank line
    const int r4 = lane & 3, kc = lane >> 2;

    auto load_block = [&](float (&va)[2][8], float (&vw)[2][8], int kb) {
        const int64_t k = (int64_t)kb * BK + 8 * kc;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int row = 8 * warp + r4 + 4 * ((kc & 1) ^ u);
            const int64_t m = m0 + row, n = n0 + row;
            const float* ap = A + m * lda + k;
            const float* wp = W + n * ldw + k;
            if (vec8 && k + 8 <= K) {                              // one sector-complete 256-bit access per operand
#pragma unroll
                for (int i = 0; i < 8; ++i) va[u][i] = vw[u][i] = 0.0f;
                if (m < M) ldg256(ap, va[u]);
                if (n < N) ldg256(wp, vw[u]);
            } else if (vec4 && k + 8 <= K) {
                float4 a0 = make_float4(0, 0, 0, 0), a1 = a0, w0 = a0, w1 = a0;
                if (m < M) { a0 = __ldg(reinterpret_cast<const float4*>(ap)); a1 = __ldg(reinterpret_cast<const float4*>(ap) + 1); }
                if (n < N) { w0 = __ldg(reinterpret_cast<const float4*>(wp)); w1 = __ldg(reinterpret_cast<const float4*>(wp) + 1); }
                va[u][0] = a0.x; va[u][1] = a0.y; va[u][2] = a0.z; va[u][3] = a0.w; va[u][4] = a1.x; va[u][5] = a1.y; va[u][6] = a1.z; va[u][7] = a1.w;
                vw[u][0] = w0.x; vw[u][1] = w0.y; vw[u][2] = w0.z; vw[u][3] = w0.w; vw[u][4] = w1.x; vw[u][5] = w1.y; vw[u][6] = w1.z; vw[u][7] = w1.w;
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const bool kin = k + i < K;
                    va[u][i] = (kin && m < M) ? __ldg(ap + i) : 0.0f;
                    vw[u][i] = (kin && n < N) ? __ldg(wp + i) : 0.0f;
                }
            }
        }
    };
    auto store_block = [&](const float (&va)[2][8], const float (&vw)[2][8], int buf) {
        uint8_t* sa = smem_raw + buf * 2 * OPB;
        uint8_t* sw = sa + OPB;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int row = 8 * warp + r4 + 4 * ((kc & 1) ^ u);
            const uint32_t off = (uint32_t)((row >> 3) * SBO + (row & 7) * 16 + kc * LBO);
            store_chunk8<PREC>(sa + off, T_BYTES, va[u]);
            store_chunk8<PREC>(sw + off, T_BYTES, vw[u]);
        }
    };
    auto issue_block = [&](int buf, uint32_t accumulate) {          // ONE elected lane
        const uint32_t a_addr = base + buf * 2 * OPB, w_addr = a_addr + OPB;
        const uint32_t idesc = make_instr_desc(1u, BM, BN);
        uint32_t acc = accumulate;
        constexpr int NTERM = PREC == PREC_BF16X3 ? 6 : 1;
        const uint32_t ta[6] = {2, 0, 1, 1, 0, 0}, tw[6] = {0, 2, 1, 0, 1, 0};   // smallest products first
#pragma unroll
        for (int t = 0; t < NTERM; ++t)
#pragma unroll
            for (int ks = 0; ks < BK / 16; ++ks) {
                const uint32_t at = PREC == PREC_BF16X3 ? ta[t] : 0, wt = PREC == PREC_BF16X3 ? tw[t] : 0;
                mma_bf16(tmem, make_smem_desc(a_addr + at * T_BYTES + ks * 2 * LBO, LBO, SBO),
                         make_smem_desc(w_addr + wt * T_BYTES + ks * 2 * LBO, LBO, SBO), idesc, acc);
                acc = 1;
            }
    };

    float va[2][8], vw[2][8];
    if (nkb > 0) {
        load_block(va, vw, kb0);
        store_block(va, vw, 0);
    }
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    for (int i = 0; i < nkb; ++i) {
        const int buf = i & 1;
        if (warp == 0) {
            if (elect_one()) {
                fence_after_sync();
                issue_block(buf, i > 0);
                mma_commit(&mbar[buf]);
            }
            __syncwarp();
        }
        if (i + 1 < nkb) {
            load_block(va, vw, kb0 + i + 1);                                   // in flight while the MMAs run
            if (i >= 1) mbar_wait(&mbar[buf ^ 1], ((i - 1) >> 1) & 1);         // MMAs of block i-1 are done with that buffer
            store_block(va, vw, buf ^ 1);
            fence_async_smem();
        }
        fence_before_sync();
        __syncthreads();
    }
    // ---- epilogue: thread = one row x 32 columns ---------------------------------------------------------------
    if (nkb > 0) {
        const int last = nkb - 1;
        mbar_wait(&mbar[last & 1], (last >> 1) & 1);
        if (nkb > 1) mbar_wait(&mbar[(last - 1) & 1], ((last - 1) >> 1) & 1);
    }
    fence_after_sync();
    {
        // TMEM -> registers (thread = one row x 32 columns) -> fp32 tile in the (now idle) operand buffers -> global
        // memory with a warp per row and a lane per column: full 128-byte segments whatever the row stride
        constexpr int LDT = BN + 1;                                   // odd stride: the row-per-lane writes are conflict free
        float* T = reinterpret_cast<float*>(smem_raw);                // 128 x 129 floats = 66 KB
        const int q = warp & 3, cq = warp >> 2, row = 32 * q + lane;
        float z[32];
        if (nkb > 0) {
            tmem_ld<32>(tmem + ((uint32_t)(32 * q) << 16) + 32 * cq, z);
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) z[i] = 0.0f;
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) T[row * LDT + 32 * cq + i] = z[i];
        __syncthreads();
        const bool direct = split_k == 1;
        float bv[BN / 32];
#pragma unroll
        for (int j = 0; j < BN / 32; ++j) {
            const int64_t n = n0 + lane + 32 * j;
            bv[j] = (direct && bias && n < N) ? __ldg(bias + n) : 0.0f;
        }
        for (int r = warp; r < BM; r += NT / 32) {
            const int64_t m = m0 + r;
            if (m >= M) break;
            float* dst = direct ? C + m * ldc : part + ((int64_t)blockIdx.z * M + m) * N;
#pragma unroll
            for (int j = 0; j < BN / 32; ++j) {
                const int64_t n = n0 + lane + 32 * j;
                if (n < N) {
                    float v = T[r * LDT + lane + 32 * j] + bv[j];
                    if (direct && relu) v = fmaxf(v, 0.0f);
                    dst[n] = v;
                }
            }
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, BN);
}

// C[m,n] = act(sum_s part[s][m][n] + b[n]) in slice order (deterministic)
__global__ void linear_reduce_kernel(const float* __restrict__ part, const float* __restrict__ bias, float* __restrict__ C,
                                     int64_t ldc, int64_t M, int64_t N, int split_k, int relu) {
    const int64_t total = M * N;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t m = idx / N, n = idx - m * N;
        float v = 0.0f;
        for (int s = 0; s < split_k; ++s) v += __ldg(part + (int64_t)s * total + idx);
        if (bias) v += __ldg(bias + n);
        if (relu) v = fmaxf(v, 0.0f);
        C[m * ldc + n] = v;
    }
}

template <int PREC>
static int launch_linear(const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias, float* C, int64_t ldc,
                         float* part, int64_t M, int64_t N, int64_t K, int relu, int split_k, cudaStream_t st) {
    using namespace lin;
    size_t smem = (size_t)2 * 2 * TcCfg<PREC>::NSPLIT * T_BYTES;
    const size_t epi = (size_t)BM * (BN + 1) * sizeof(float);          // the epilogue's fp32 tile reuses the operand buffers
    if (smem < epi) smem = epi;
    cudaError_t e = cudaFuncSetAttribute(linear_tc_kernel<PREC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    dim3 grid((unsigned)((N + BN - 1) / BN), (unsigned)((M + BM - 1) / BM), (unsigned)split_k);
    linear_tc_kernel<PREC><<<grid, NT, smem, st>>>(A, lda, W, ldw, bias, C, ldc, part, M, N, K, relu, split_k);
    e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    if (split_k > 1) {
        const int64_t total = M * N;
        int blocks = (int)((total + 255) / 256);
        if (blocks > 148 * 8) blocks = 148 * 8;
        linear_reduce_kernel<<<blocks, 256, 0, st>>>(part, bias, C, ldc, M, N, split_k, relu);
        e = cudaGetLastError();
    }
    return e == cudaSuccess ? 0 : (int)e;
}

}  // namespace is

using namespace is;

extern "C" {

// number of K slices that fills the GPU for an [M,N,K] problem (1 = no workspace needed)
int is_linear_tc_split_k(int64_t M, int64_t N, int64_t K) {
    const int sms = current_num_sms();
    const int64_t tiles = ((M + lin::BM - 1) / lin::BM) * ((N + lin::BN - 1) / lin::BN);
    const int64_t nkb = (K + lin::BK - 1) / lin::BK;
    int64_t s = tiles >= sms ? 1 : sms / tiles;
    if (s > nkb / 4) s = nkb / 4;               // at least four K blocks per slice
    return (int)(s < 1 ? 1 : s);
}

// C[M,N] = act(A[M,K] W[N,K]^T + bias) on the tensor cores.  lda / ldw / ldc = row strides in floats; bias may be
// NULL; relu != 0 applies max(.,0).  precision 0 = bf16 operands, 3 = bf16x3 (fp32-accurate).  split_k > 1 needs
// workspace >= split_k * M * N floats.
int is_linear_tc(const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias, float* C, int64_t ldc,
                 int64_t M, int64_t N, int64_t K, int relu, int precision, int split_k, float* workspace, void* stream) {
    if (M <= 0 || N <= 0 || K <= 0 || split_k < 1 || split_k > 65535 || (split_k > 1 && workspace == nullptr)) return IS_ERR_ARG;
    if ((M + lin::BM - 1) / lin::BM > 65535) return IS_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    if (precision == PREC_BF16) return launch_linear<PREC_BF16>(A, lda, W, ldw, bias, C, ldc, workspace, M, N, K, relu, split_k, st);
    if (precision == PREC_BF16X3) return launch_linear<PREC_BF16X3>(A, lda, W, ldw, bias, C, ldc, workspace, M, N, K, relu, split_k, st);
    return IS_ERR_ARG;
}

}  // extern "C"
