// Dense GEMM on the tcgen05 tensor cores fed by TMA, fp32-accurate through the bf16x3 operand split:
//     C[M, N] = act( A[M, K] * B[N, K]^T + bias[N] ),      fp32 in / out.
//
// Reference: the nn.Linear layers of the sequence VAE, forward AND backward (models/hybrid_models.py:63-74 /
// 297-308: vae_fc1 5943 -> 512 + ReLU, vae_fc4 512 -> 5943; 96 % of the model's parameters), which the reference runs
// as cuBLAS fp32 SIMT GEMMs.  All three products of a Linear layer are the same "both operands K-major" GEMM once the
// operands are stored the right way round:
//     forward   Y  = X  W^T        A = X   [B, in]      B = W    [out, in]
//     dgrad     gX = gY W          A = gY  [B, out]     B = W^T  [in, out]
//     wgrad     gW = gY^T X        A = gY^T [out, B]    B = X^T  [in, B]
// so a small streaming pre-pass (`split_planes_kernel`) converts an fp32 matrix ONCE into three bf16 planes
// (a = a1 + a2 + a3, round-to-nearest residuals: a1 + a2 + a3 == a to 2^-24) in row-major and / or transposed
// form, rows padded to a multiple of 8 elements (16-byte TMA strides, zero filled).  Weights are split once per
// parameter version (host cache), activations once per use -- the GEMM itself does no SIMT work on operands.
//
// GEMM kernel (one 128 x 128 output tile per CTA, optional split-K), warp-specialised:
//   warp 0   TMA producer: per 64-wide K block ONE cp.async.bulk.tensor per operand -- a 3-D box {64 k, 128 rows,
//            3 planes} = 48 KB, SWIZZLE_128B -- into a 2-stage ring (2 x 96 KB), completion on an mbarrier (expect_tx)
//   warp 1   MMA issuer: 6 partial products x 4 K-steps of tcgen05.mma (128 x 128 x 16, kind::f16) per stage,
//            smallest products first, fp32 accumulation in TMEM (128 columns); tcgen05.commit frees the stage
//   warps 2-5 epilogue: tcgen05.ld -> fp32 tile in shared memory -> bias / ReLU -> 128-byte coalesced stores (any
//            row stride: 5943 floats is not 16-byte aligned), or the split-K partial tile.
// If the pre-pass found an operand exactly representable in bf16 (one-hot sequence inputs: a2 = a3 = 0, device flag),
// the three products involving its second and third planes are skipped.
#include <cuda.h>

#include "tc_common.cuh"

namespace is {
namespace gt {

constexpr int BM = 128, BN = 128, BK = 64;
constexpr uint32_t PLANE_BYTES = BM * BK * 2;            // one 128 x 64 bf16 plane tile (16 KB, 128-byte swizzled rows)
constexpr int NSTAGE = 2;
constexpr int NT = 192;                                   // 6 warps

// 64-bit shared-memory matrix descriptor, K-major, SWIZZLE_128B: 8-row x 128-byte atoms, SBO = 1024 bytes
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;                                // LBO (ignored for swizzled K-major operands)
    d |= (uint64_t)(1024 >> 4) << 32;                      // SBO: bytes between 8-row groups
    d |= (uint64_t)1 << 46;                                // descriptor version 1 (Blackwell)
    d |= (uint64_t)2 << 61;                                // layout type SWIZZLE_128B
    return d;
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

template <int NPL>
__global__ void __launch_bounds__(NT, 1)
gemm_tma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                const int* __restrict__ a_resid_flag, const int* __restrict__ b_resid_flag, const float* __restrict__ bias, float* __restrict__ C, int64_t ldc,
                float* __restrict__ part, int64_t M, int64_t N, int nkb_all, int relu, int split_k) {
    constexpr uint32_t OP_BYTES = NPL * PLANE_BYTES, STAGE_BYTES = 2 * OP_BYTES;
    extern __shared__ uint8_t smem_dyn[];
    __shared__ __align__(8) uint64_t full[NSTAGE], empty[NSTAGE], acc_full;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;             // SWIZZLE_128B tiles: 1024-byte aligned
    uint8_t* base_ptr = smem_dyn + (base - smem_u32(smem_dyn));
    const int64_t m0 = (int64_t)blockIdx.y * BM, n0 = (int64_t)blockIdx.x * BN;
    const int kb0 = (int)((int64_t)nkb_all * blockIdx.z / split_k), kb1 = (int)((int64_t)nkb_all * (blockIdx.z + 1) / split_k);
    const int nkb = kb1 - kb0;

    if (warp == 0 && lane == 0) { prefetch_tmap(&mapA); prefetch_tmap(&mapB); }
    if (tid == 32) {
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(&acc_full, 1);
    }
    if (warp == 2) tmem_alloc(&s_tmem, NPL == 3 ? 2 * BN : BN);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = s_tmem;

    if (warp == 0) {
        // ================= TMA producer =================
        if (elect_one()) {
            for (int i = 0; i < nkb; ++i) {
                const int s = i % NSTAGE, n = i / NSTAGE;
                mbar_wait(&empty[s], (n & 1) ^ 1);                            // passes on a fresh barrier
                mbar_expect_tx(&full[s], STAGE_BYTES);
                const int k = (kb0 + i) * BK;
                tma_load_3d(base + s * STAGE_BYTES, &mapA, k, (int)m0, 0, &full[s]);
                tma_load_3d(base + s * STAGE_BYTES + OP_BYTES, &mapB, k, (int)n0, 0, &full[s]);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ================= MMA issuer =================
        const bool a_exact = (NPL == 3 && a_resid_flag != nullptr) ? (*a_resid_flag == 0) : false;
        const bool b_exact = (NPL == 3 && b_resid_flag != nullptr) ? (*b_resid_flag == 0) : false;
        const uint32_t idesc = make_instr_desc(1u, BM, BN);
        for (int i = 0; i < nkb; ++i) {
            const int s = i % NSTAGE, n = i / NSTAGE;
            mbar_wait(&full[s], n & 1);
            fence_after_sync();
            if (elect_one()) {
                const uint32_t a_addr = base + s * STAGE_BYTES, b_addr = a_addr + OP_BYTES;
                uint32_t acc = i > 0 ? 1u : 0u;
                if (NPL == 3) {
                    // (a-term, b-term), smallest products first: a3b1 a1b3 a2b2 a2b1 a1b2 | a1b1.  TWO accumulators: the
                    // five correction products (<= 2^-8 of the result) go to columns [BN, 2 BN), the leading product
                    // a1 b1 alone to [0, BN).  Every accumulation step rounds at the accumulator's magnitude; keeping the
                    // 20 correction steps per K block out of the leading accumulator leaves it 4 steps per block
                    // instead of 24, and the corrections' own rounding is 2^-8 smaller.  The epilogue adds the two.
                    const uint32_t ta[6] = {2, 0, 1, 1, 0, 0}, tb[6] = {0, 2, 1, 0, 1, 0};
                    uint32_t acc_c = acc;
                    bool any_c = false;
#pragma unroll
                    for (int t = 0; t < 5; ++t) {
                        if ((a_exact && ta[t] != 0) || (b_exact && tb[t] != 0)) continue;
                        any_c = true;
#pragma unroll
                        for (int ks = 0; ks < BK / 16; ++ks) {
                            mma_bf16(tmem + BN, make_sw128_desc(a_addr + ta[t] * PLANE_BYTES + ks * 32),
                                     make_sw128_desc(b_addr + tb[t] * PLANE_BYTES + ks * 32), idesc, acc_c);
                            acc_c = 1;
                        }
                    }
                    (void)any_c;
#pragma unroll
                    for (int ks = 0; ks < BK / 16; ++ks) {
                        mma_bf16(tmem, make_sw128_desc(a_addr + ks * 32), make_sw128_desc(b_addr + ks * 32), idesc, acc);
                        acc = 1;
                    }
                } else {
#pragma unroll
                    for (int ks = 0; ks < BK / 16; ++ks) {
                        mma_bf16(tmem, make_sw128_desc(a_addr + ks * 32), make_sw128_desc(b_addr + ks * 32), idesc, acc);
                        acc = 1;
                    }
                }
                mma_commit(&empty[s]);                                        // stage free once these MMAs have read it
                if (i == nkb - 1) mma_commit(&acc_full);
            }
            __syncwarp();
        }
    } else {
        // ================= epilogue (warps 2-5: TMEM lane quarter = warp % 4) =================
        constexpr int LDT = BN + 1;
        const bool has_corr = !((a_resid_flag != nullptr && *a_resid_flag == 0) && (b_resid_flag != nullptr && *b_resid_flag == 0));
        float* T = reinterpret_cast<float*>(base_ptr);                        // 128 x 129 floats over the (idle) stage ring
        const int q = warp & 3, row = 32 * q + lane;
        if (nkb > 0) {
            mbar_wait(&acc_full, 0);
            fence_after_sync();
        }
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
            float z[32];
            if (nkb > 0) {
                tmem_ld<32>(tmem + ((uint32_t)(32 * q) << 16) + 32 * c, z);
                if (NPL == 3 && has_corr) {                                   // + the correction accumulator
                    float zc[32];
                    tmem_ld<32>(tmem + ((uint32_t)(32 * q) << 16) + BN + 32 * c, zc);
#pragma unroll
                    for (int i = 0; i < 32; ++i) z[i] += zc[i];
                }
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) z[i] = 0.0f;
            }
#pragma unroll
            for (int i = 0; i < 32; ++i) T[row * LDT + 32 * c + i] = z[i];
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const bool direct = split_k == 1;
        float bv[BN / 32];
#pragma unroll
        for (int j = 0; j < BN / 32; ++j) {
            const int64_t n = n0 + lane + 32 * j;
            bv[j] = (direct && bias && n < N) ? __ldg(bias + n) : 0.0f;
        }
        for (int r = warp - 2; r < BM; r += 4) {
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
    if (warp == 2) tmem_dealloc(tmem, NPL == 3 ? 2 * BN : BN);
}

// C[m, n] = act(sum_s part[s][m][n] + b[n]) in slice order (deterministic)
__global__ void gemm_reduce_kernel(const float* __restrict__ part, const float* __restrict__ bias, float* __restrict__ C,
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

// ---- pre-pass: fp32 [R, C] -> bf16 planes, row-major [NPL][R][Cp] and / or transposed [NPL][C][Rp] ---------------
// 32 x 32 tile per 256-thread CTA.  Optional: ReLU mask (element kept where relu_src > 0), per-CTA column sums
// (bias gradient; summed over the row blocks by is_reduce_partials), residual flag (any a2 != 0).
__device__ __forceinline__ void split3(float a, __nv_bfloat16& p1, __nv_bfloat16& p2, __nv_bfloat16& p3) {
    p1 = __float2bfloat16_rn(a);
    const float r1 = a - __bfloat162float(p1);
    p2 = __float2bfloat16_rn(r1);
    p3 = __float2bfloat16_rn(r1 - __bfloat162float(p2));
}

template <int NPL>
__global__ void __launch_bounds__(256)
split_planes_kernel(const float* __restrict__ X, int64_t ld, int64_t R, int64_t Cn, const float* __restrict__ relu_src,
                    int64_t ld_relu, __nv_bfloat16* __restrict__ P, int64_t Cp, __nv_bfloat16* __restrict__ Tp, int64_t Rp,
                    float* __restrict__ colsum_part, int* __restrict__ resid_flag) {
    __shared__ float tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t c0 = (int64_t)blockIdx.x * 32, r0 = (int64_t)blockIdx.y * 32;
    bool resid = false;
    float csum = 0.0f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t r = r0 + ty + 8 * i, c = c0 + tx;
        float v = 0.0f;
        if (r < R && c < Cn) {
            v = __ldg(X + r * ld + c);
            if (relu_src != nullptr && !(__ldg(relu_src + r * ld_relu + c) > 0.0f)) v = 0.0f;
        }
        tile[ty + 8 * i][tx] = v;
        csum += v;
        if (P != nullptr && r < R && c < Cp) {
            __nv_bfloat16 p1, p2, p3;
            split3(v, p1, p2, p3);
            P[r * Cp + c] = p1;
            if (NPL == 3) {
                P[(R + r) * Cp + c] = p2;
                P[(2 * R + r) * Cp + c] = p3;
                resid |= (__bfloat162float(p2) != 0.0f);
            }
        }
    }
    __syncthreads();
    if (Tp != nullptr) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int64_t c = c0 + ty + 8 * i, r = r0 + tx;                   // transposed: lanes run along the rows
            if (c < Cn && r < Rp) {
                const float v = tile[tx][ty + 8 * i];                         // zero beyond R (tile rows were zero filled)
                __nv_bfloat16 p1, p2, p3;
                split3(v, p1, p2, p3);
                Tp[c * Rp + r] = p1;
                if (NPL == 3) {
                    Tp[(Cn + c) * Rp + r] = p2;
                    Tp[(2 * Cn + c) * Rp + r] = p3;
                    resid |= (__bfloat162float(p2) != 0.0f);
                }
            }
        }
    }
    if (colsum_part != nullptr) {
        __syncthreads();
        tile[ty][tx] = csum;                                                  // rows ty, ty + 8, ... of column tx
        __syncthreads();
        if (ty == 0 && c0 + tx < Cn) {
            float s = 0.0f;
#pragma unroll
            for (int l = 0; l < 8; ++l) s += tile[l][tx];
            colsum_part[(int64_t)blockIdx.y * Cn + c0 + tx] = s;
        }
    }
    if (resid_flag != nullptr && __syncthreads_or(resid ? 1 : 0) && threadIdx.x == 0) atomicOr(resid_flag, 1);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// planes [npl][rows][kp] bf16, K contiguous: 3-D map {kp, rows, npl}, box {64, 128, npl}, 128-byte swizzle, zero OOB fill
static int make_plane_map(CUtensorMap* map, const void* planes, int64_t rows, int64_t kp, int npl) {
    EncodeTiledFn enc = encode_fn();
    if (enc == nullptr) return IS_ERR_UNSUPPORTED;
    const cuuint64_t dims[3] = {(cuuint64_t)kp, (cuuint64_t)rows, (cuuint64_t)npl};
    const cuuint64_t strides[2] = {(cuuint64_t)kp * 2, (cuuint64_t)rows * (cuuint64_t)kp * 2};
    const cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)BM, (cuuint32_t)npl};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(planes), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? IS_OK : 1000 + (int)r;
}

template <int NPL>
static int launch_gemm(const void* Ap, const void* Bp, int64_t M, int64_t N, int64_t Kp, const int* flag, const int* flag_b,
                       const float* bias,
                       int relu, float* C, int64_t ldc, int split_k, float* ws, cudaStream_t st) {
    CUtensorMap mapA, mapB;
    int rc = make_plane_map(&mapA, Ap, M, Kp, NPL);
    if (rc != IS_OK) return rc;
    rc = make_plane_map(&mapB, Bp, N, Kp, NPL);
    if (rc != IS_OK) return rc;
    size_t smem = (size_t)NSTAGE * 2 * NPL * PLANE_BYTES;
    const size_t epi = (size_t)BM * (BN + 1) * sizeof(float);
    if (smem < epi) smem = epi;
    smem += 1024;                                                             // alignment slack
    cudaError_t e = cudaFuncSetAttribute(gemm_tma_kernel<NPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const int nkb = (int)((Kp + BK - 1) / BK);
    dim3 grid((unsigned)((N + BN - 1) / BN), (unsigned)((M + BM - 1) / BM), (unsigned)split_k);
    gemm_tma_kernel<NPL><<<grid, NT, smem, st>>>(mapA, mapB, flag, flag_b, bias, C, ldc, ws, M, N, nkb, relu, split_k);
    e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    if (split_k > 1) {
        const int64_t total = M * N;
        int blocks = (int)((total + 255) / 256);
        if (blocks > 148 * 8) blocks = 148 * 8;
        gemm_reduce_kernel<<<blocks, 256, 0, st>>>(ws, bias, C, ldc, M, N, split_k, relu);
        e = cudaGetLastError();
    }
    return e == cudaSuccess ? IS_OK : (int)e;
}

}  // namespace gt
}  // namespace is

using namespace is;
using namespace is::gt;

extern "C" {

// number of K slices that fills the GPU for an [M, N, Kp] problem (1 = no workspace needed)
int is_gemm_tma_split_k(int64_t M, int64_t N, int64_t Kp) {
    const int sms = current_num_sms();
    const int64_t tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
    const int64_t nkb = (Kp + BK - 1) / BK;
    int64_t s = tiles >= sms ? 1 : sms / tiles;
    if (s > nkb / 4) s = nkb / 4;                         // at least four K blocks per slice
    return (int)(s < 1 ? 1 : s);
}

// fp32 X [R, C] (row stride ld) -> bf16 planes.  n_planes 3 (a1 + a2 + a3) or 1 (rounded).  planes [n_planes][R][Cp]
// (Cp = C rounded up to 8, padding zero filled) and / or planes_t [n_planes][C][Rp] (Rp = R rounded up to 8); either may
// be NULL.  relu_src (row stride ld_relu): X is masked where relu_src <= 0.  colsum_part [ceil(R/32)][C]: per-row-block
// column sums of the (masked) X (sum them with is_reduce_partials) or NULL.  resid_flag: set to 1 if some a2 != 0
// (must be zeroed by the caller) or NULL.
int is_split_planes(const float* X, int64_t ld, int64_t R, int64_t C, const float* relu_src, int64_t ld_relu, int n_planes,
                    void* planes, void* planes_t, float* colsum_part, int* resid_flag, void* stream) {
    if (R <= 0 || C <= 0 || (n_planes != 1 && n_planes != 3) || (!planes && !planes_t && !colsum_part)) return IS_ERR_ARG;
    const int64_t Cp = (C + 7) / 8 * 8, Rp = (R + 7) / 8 * 8;
    // the grid covers the padded extents so that the padding is written (zeros)
    const int64_t gc = ((planes ? Cp : C) + 31) / 32, gr = ((planes_t ? Rp : R) + 31) / 32;
    if (gr > 65535) return IS_ERR_UNSUPPORTED;
    dim3 grid((unsigned)gc, (unsigned)gr);
    cudaStream_t st = (cudaStream_t)stream;
    if (n_planes == 3)
        split_planes_kernel<3><<<grid, 256, 0, st>>>(X, ld, R, C, relu_src, ld_relu, (__nv_bfloat16*)planes, Cp,
                                                     (__nv_bfloat16*)planes_t, Rp, colsum_part, resid_flag);
    else
        split_planes_kernel<1><<<grid, 256, 0, st>>>(X, ld, R, C, relu_src, ld_relu, (__nv_bfloat16*)planes, Cp,
                                                     (__nv_bfloat16*)planes_t, Rp, colsum_part, resid_flag);
    IS_LAUNCH_CHECK();
    return IS_OK;
}

// C[M, N] = act(A B^T + bias) from pre-split planes: A_planes [n_planes][M][Kp], B_planes [n_planes][N][Kp], bf16, Kp a
// multiple of 8, 16-byte aligned.  a_resid_flag / b_resid_flag (device ints, or NULL): 0 = that operand is exact in bf16
// (its second and third planes are zero: the products involving them are skipped).  split_k > 1 needs workspace >= split_k * M * N floats.  Return codes > 1000: CUresult of
// cuTensorMapEncodeTiled + 1000.
int is_gemm_planes_tma(const void* A_planes, int64_t M, const void* B_planes, int64_t N, int64_t Kp, int n_planes,
                       const int* a_resid_flag, const int* b_resid_flag, const float* bias, int relu, float* C, int64_t ldc,
                       int split_k, float* workspace, void* stream) {
    if (M <= 0 || N <= 0 || Kp <= 0 || (Kp & 7) != 0 || split_k < 1 || split_k > 65535 || (split_k > 1 && !workspace)) return IS_ERR_ARG;
    if (((reinterpret_cast<uintptr_t>(A_planes) | reinterpret_cast<uintptr_t>(B_planes)) & 15) != 0) return IS_ERR_ARG;
    if ((M + BM - 1) / BM > 65535) return IS_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    if (n_planes == 3) return launch_gemm<3>(A_planes, B_planes, M, N, Kp, a_resid_flag, b_resid_flag, bias, relu, C, ldc, split_k, workspace, st);
    if (n_planes == 1) return launch_gemm<1>(A_planes, B_planes, M, N, Kp, nullptr, nullptr, bias, relu, C, ldc, split_k, workspace, st);
    return IS_ERR_ARG;
}

}  // extern "C"
