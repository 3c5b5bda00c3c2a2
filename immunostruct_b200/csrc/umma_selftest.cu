// Self-test of the hand-written tcgen05 path: D[128,64] = A[128,64] * B[64,64]^T for one tile, with the
// exact device helpers (canonical layout, descriptors, MMA issue, commit, TMEM load) that the fused
// EGNN tensor-core kernels use.  mode 0: bf16 operands; 1: tf32; 2: 3xTF32 split (fp32-accurate);
// 3: bf16x3 split (a = a1 + a2 + a3 in bf16, six partial products, fp32-accurate, 6 bytes / element).
#include "common.cuh"
#include "umma.cuh"

namespace is {

template <int MODE>
__global__ void __launch_bounds__(128)
umma_selftest_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D) {
    using namespace umma;
    constexpr int EB = (MODE == 0 || MODE == 3) ? 2 : 4;  // operand element bytes
    constexpr int KCH = 64 * EB / 16;                     // 16-byte chunks per row (8 for bf16, 16 for tf32)
    constexpr uint32_t SBO = KCH * kLBO;
    constexpr uint32_t A_BYTES = 16 * SBO, B_BYTES = 8 * SBO;
    constexpr int NSPLIT = MODE == 2 ? 2 : MODE == 3 ? 3 : 1;   // operand copies (split terms)
    extern __shared__ __align__(128) uint8_t sm[];
    uint8_t* sA = sm;                                     // [NSPLIT][A_BYTES]
    uint8_t* sB = sm + NSPLIT * A_BYTES;                  // [NSPLIT][B_BYTES]
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (warp == 0) tmem_alloc(&tmem_base, 64);
    if (tid == 32) mbar_init(&mbar, 1);
    // stage operands in the canonical layout
    for (int idx = tid; idx < 128 * 64; idx += 128) {
        const int r = idx >> 6, k = idx & 63;
        const float a = A[idx];
        const uint32_t off = canon_off<EB>(r, k, KCH);
        if (MODE == 0) {
            *reinterpret_cast<__nv_bfloat16*>(sA + off) = __float2bfloat16_rn(a);
        } else if (MODE == 3) {
            const __nv_bfloat16 a1 = __float2bfloat16_rn(a);
            const float r1 = a - __bfloat162float(a1);
            const __nv_bfloat16 a2 = __float2bfloat16_rn(r1);
            const __nv_bfloat16 a3 = __float2bfloat16_rn(r1 - __bfloat162float(a2));
            *reinterpret_cast<__nv_bfloat16*>(sA + off) = a1;
            *reinterpret_cast<__nv_bfloat16*>(sA + A_BYTES + off) = a2;
            *reinterpret_cast<__nv_bfloat16*>(sA + 2 * A_BYTES + off) = a3;
        } else {
            const float hi = tf32_round(a);
            *reinterpret_cast<float*>(sA + off) = hi;
            if (MODE == 2) *reinterpret_cast<float*>(sA + A_BYTES + off) = tf32_round(a - hi);
        }
    }
    for (int idx = tid; idx < 64 * 64; idx += 128) {
        const int r = idx >> 6, k = idx & 63;
        const float b = B[idx];
        const uint32_t off = canon_off<EB>(r, k, KCH);
        if (MODE == 0) {
            *reinterpret_cast<__nv_bfloat16*>(sB + off) = __float2bfloat16_rn(b);
        } else if (MODE == 3) {
            const __nv_bfloat16 b1 = __float2bfloat16_rn(b);
            const float r1 = b - __bfloat162float(b1);
            const __nv_bfloat16 b2 = __float2bfloat16_rn(r1);
            const __nv_bfloat16 b3 = __float2bfloat16_rn(r1 - __bfloat162float(b2));
            *reinterpret_cast<__nv_bfloat16*>(sB + off) = b1;
            *reinterpret_cast<__nv_bfloat16*>(sB + B_BYTES + off) = b2;
            *reinterpret_cast<__nv_bfloat16*>(sB + 2 * B_BYTES + off) = b3;
        } else {
            const float hi = tf32_round(b);
            *reinterpret_cast<float*>(sB + off) = hi;
            if (MODE == 2) *reinterpret_cast<float*>(sB + B_BYTES + off) = tf32_round(b - hi);
        }
    }
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tbase = tmem_base;
    if (tid == 0) {
        const uint32_t idesc = make_instr_desc((MODE == 0 || MODE == 3) ? 1u : 2u, 128, 64);
        const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
        uint32_t acc = 0;
        if (MODE == 3) {
            // (a-term, b-term) pairs, smallest products first: a3b1 a1b3 a2b2 a2b1 a1b2 a1b1
            const int ta[6] = {2, 0, 1, 1, 0, 0}, tb[6] = {0, 2, 1, 0, 1, 0};
            for (int t = 0; t < 6; ++t)
                for (int ks = 0; ks < KCH / 2; ++ks) {
                    const uint64_t da = make_smem_desc(a0 + ta[t] * A_BYTES + ks * 2 * kLBO, kLBO, SBO);
                    const uint64_t db = make_smem_desc(b0 + tb[t] * B_BYTES + ks * 2 * kLBO, kLBO, SBO);
                    mma_bf16(tbase, da, db, idesc, acc);
                    acc = 1;
                }
        } else {
        // small terms first: A_lo*B_hi, A_hi*B_lo, then A_hi*B_hi
        for (int term = (MODE == 2 ? 0 : 2); term < 3; ++term) {
            const uint32_t aoff = (term == 0) ? A_BYTES : 0, boff = (term == 1) ? B_BYTES : 0;
            for (int ks = 0; ks < KCH / 2; ++ks) {
                const uint64_t da = make_smem_desc(a0 + aoff + ks * 2 * kLBO, kLBO, SBO);
                const uint64_t db = make_smem_desc(b0 + boff + ks * 2 * kLBO, kLBO, SBO);
                if (MODE == 0) mma_bf16(tbase, da, db, idesc, acc); else mma_tf32(tbase, da, db, idesc, acc);
                acc = 1;
            }
        }
        }
        mma_commit(&mbar);
    }
    mbar_wait(&mbar, 0);
    fence_after_sync();
    float v[32];
    for (int half = 0; half < 2; ++half) {
        tmem_ld32(tbase + ((uint32_t)(warp * 32) << 16) + half * 32, v);
        const int row = warp * 32 + lane;
#pragma unroll
        for (int c = 0; c < 32; ++c) D[row * 64 + half * 32 + c] = v[c];
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 64);
}


// Layout experiments for the tensor-core backward (weight gradients need transposed operands):
//  mode 4: B operand MN-major  -- B^T is staged as a K-major tile (rows = k, cols = n) and read through an
//          MN-major descriptor (SBO <- LBO, LBO <- SBO, b_major = 1): D must still equal A B^T.
//  mode 5: A operand MN-major  -- same trick for A (tile rows = k, cols = m).
//  mode 6: M = 64 accumulator  -- D[64,64] = A[0:64] B^T with an M = 64 instruction; all 128 TMEM lanes are
//          dumped so the lane mapping of the 64 rows can be read off.
//  mode 7: A operand from TENSOR MEMORY -- every thread packs its own row of A to bf16 pairs and stores it into its
//          TMEM lane (32 columns); the MMAs take A from TMEM (8 columns per K step of 16) and B from shared memory.
//  mode 8: A operand copied shared memory -> TENSOR MEMORY by the tensor core's own copy (tcgen05.cp.128x256b: 128 rows x
//          256 bits = one K step of the K-major bf16 tile per copy, same descriptor as the MMA's A operand), then as mode 7.
//          Copies and MMAs of one thread execute in issue order: no barrier between them.
template <int MODE>
__global__ void __launch_bounds__(128)
umma_layout_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D) {
    using namespace umma;
    constexpr int KCH8 = 8;                                  // 64 bf16 columns = 8 chunks
    constexpr int KCH16 = 16;                                // 128 bf16 columns
    extern __shared__ __align__(128) uint8_t sm[];
    uint8_t* sA = sm;                                        // up to 128 rows x 64 cols or 64 rows x 128 cols
    uint8_t* sB = sm + 16 * KCH16 * kLBO;
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) tmem_alloc(&tmem_base, 128);
    if (tid == 32) mbar_init(&mbar, 1);
    if (MODE == 7) {
        __syncthreads();                                     // tmem_base visible
        const uint32_t ta = tmem_base + ((uint32_t)(warp * 32) << 16) + 64;
        const float* arow = A + (size_t)tid * 64;
#pragma unroll
        for (int c = 0; c < 32; c += 4)
            tmem_st4(ta + c, pack_bf16x2(arow[2 * c], arow[2 * c + 1]), pack_bf16x2(arow[2 * c + 2], arow[2 * c + 3]),
                     pack_bf16x2(arow[2 * c + 4], arow[2 * c + 5]), pack_bf16x2(arow[2 * c + 6], arow[2 * c + 7]));
        tmem_st_wait();
    }
    for (int idx = tid; idx < 128 * 64; idx += 128) {
        const int m = idx >> 6, k = idx & 63;
        const __nv_bfloat16 v = __float2bfloat16_rn(A[idx]);
        if (MODE == 5) *reinterpret_cast<__nv_bfloat16*>(sA + canon_off<2>(k, m, KCH16)) = v;     // tile rows = k, cols = m
        else *reinterpret_cast<__nv_bfloat16*>(sA + canon_off<2>(m, k, KCH8)) = v;
    }
    for (int idx = tid; idx < 64 * 64; idx += 128) {
        const int n = idx >> 6, k = idx & 63;
        const __nv_bfloat16 v = __float2bfloat16_rn(B[idx]);
        if (MODE == 4) *reinterpret_cast<__nv_bfloat16*>(sB + canon_off<2>(k, n, KCH8)) = v;      // tile rows = k, cols = n
        else *reinterpret_cast<__nv_bfloat16*>(sB + canon_off<2>(n, k, KCH8)) = v;
    }
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tbase = tmem_base;
    if (tid == 0) {
        const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
        const uint32_t idesc = make_instr_desc(1u, MODE == 6 ? 64 : 128, 64, MODE == 5 ? 1u : 0u, MODE == 4 ? 1u : 0u);
        for (int ks = 0; ks < 4; ++ks) {                     // K = 64 = 4 steps of 16
            uint64_t da, db;
            if (MODE == 5)   // MN-major A: MN-block stride = kLBO, K-group stride = row-group stride of the tile
                da = make_smem_desc(a0 + ks * 2 * (KCH16 * kLBO), KCH16 * kLBO, kLBO);
            else
                da = make_smem_desc(a0 + ks * 2 * kLBO, kLBO, KCH8 * kLBO);
            if (MODE == 4)
                db = make_smem_desc(b0 + ks * 2 * (KCH8 * kLBO), KCH8 * kLBO, kLBO);
            else
                db = make_smem_desc(b0 + ks * 2 * kLBO, kLBO, KCH8 * kLBO);
            if (MODE == 8) tmem_cp_128x256b(tbase + 64 + 8 * ks, da);
            if (MODE == 7 || MODE == 8) mma_bf16_ts(tbase, tbase + 64 + 8 * ks, db, idesc, ks > 0);
            else mma_bf16(tbase, da, db, idesc, ks > 0);
        }
        mma_commit(&mbar);
    }
    mbar_wait(&mbar, 0);
    fence_after_sync();
    float v[32];
    for (int half = 0; half < 2; ++half) {
        tmem_ld32(tbase + ((uint32_t)(warp * 32) << 16) + half * 32, v);
        const int row = warp * 32 + lane;
#pragma unroll
        for (int c = 0; c < 32; ++c) D[row * 64 + half * 32 + c] = v[c];
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 128);
}

// Cycles per tcgen05.mma, issue to completion of 96 MMAs (N = 64, bf16, K = 16 each) issued by one elected thread from a
// fully unrolled sequence with compile-time descriptors -- the way the edge kernels issue them.  NACC = number of TMEM
// accumulators the sequence rotates over (1 = every MMA accumulates into the one before it: a dependent chain).
template <int M, int AMN, int BMN, int ALBO, int ASBO, int BLBO, int BSBO, int TS, int NACC>
__device__ __forceinline__ float umma_time_one(uint32_t tb, uint32_t a0, uint32_t b0, uint64_t* mbar, uint32_t& phase, int tid) {
    using namespace umma;
    long long t0 = 0;
    if (tid < 32) {
        if (elect_one()) {
            const uint32_t idesc = make_instr_desc(1u, M, 64, AMN, BMN);
            t0 = clock64();
#pragma unroll
            for (int i = 0; i < 96; ++i) {
                const int ks = i & 3;
                const uint64_t da = make_smem_desc(a0 + ks * 2 * ALBO, ALBO, ASBO);
                const uint64_t db = make_smem_desc(b0 + ks * 2 * BLBO, BLBO, BSBO);
                const uint32_t d = tb + 64 * (i % NACC);
                if (TS) mma_bf16_ts(d, tb + 256 + 8 * ks, db, idesc, 1);
                else mma_bf16(d, da, db, idesc, 1);
            }
            mma_commit(mbar);
        }
        __syncwarp();
    }
    mbar_wait(mbar, phase);
    phase ^= 1;
    const float r = (float)(clock64() - t0) / 96.0f;
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(128)
umma_timing_kernel(float* __restrict__ out) {
    using namespace umma;
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 96 * 1024 / 16; i += 128) reinterpret_cast<uint4*>(sm)[i] = make_uint4(0u, 0u, 0u, 0u);
    if (warp == 0) tmem_alloc(&tmem_base, 512);
    if (tid == 32) mbar_init(&mbar, 1);
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tb = tmem_base, a0 = smem_u32(sm), b0 = smem_u32(sm + 48 * 1024);
    uint32_t ph = 0;
    float r[16];
    //                 M   amn bmn albo  asbo  blbo  bsbo  ts nacc
    r[0] = umma_time_one<128, 0, 0, 128, 1024, 128, 1024, 0, 1>(tb, a0, b0, &mbar, ph, tid);    // unpadded, one accumulator
    r[1] = umma_time_one<128, 0, 0, 128, 1024, 128, 1024, 0, 2>(tb, a0, b0, &mbar, ph, tid);    // two accumulators
    r[2] = umma_time_one<128, 0, 0, 128, 1024, 128, 1024, 0, 4>(tb, a0, b0, &mbar, ph, tid);    // four
    r[3] = umma_time_one<128, 0, 0, 144, 1152, 128, 1024, 0, 1>(tb, a0, b0, &mbar, ph, tid);    // A padded
    r[4] = umma_time_one<128, 0, 0, 144, 1152, 144, 1152, 0, 1>(tb, a0, b0, &mbar, ph, tid);    // both padded
    r[5] = umma_time_one<128, 0, 0, 144, 1152, 144, 1152, 0, 4>(tb, a0, b0, &mbar, ph, tid);
    r[6] = umma_time_one<128, 0, 0, 128, 1040, 128, 1040, 0, 1>(tb, a0, b0, &mbar, ph, tid);    // SBO padded
    r[7] = umma_time_one<128, 0, 0, 128, 1024, 128, 1024, 1, 1>(tb, a0, b0, &mbar, ph, tid);    // A from TMEM
    r[8] = umma_time_one<128, 0, 0, 128, 1024, 128, 1024, 1, 4>(tb, a0, b0, &mbar, ph, tid);
    r[9] = umma_time_one<128, 0, 1, 128, 1024, 1024, 128, 0, 1>(tb, a0, b0, &mbar, ph, tid);    // dgrad: B MN-major view
    r[10] = umma_time_one<128, 0, 1, 144, 1152, 1152, 144, 0, 1>(tb, a0, b0, &mbar, ph, tid);
    r[11] = umma_time_one<64, 1, 1, 1024, 128, 1024, 128, 0, 1>(tb, a0, b0, &mbar, ph, tid);    // wgrad: M = 64, both MN-major
    r[12] = umma_time_one<64, 1, 1, 1024, 128, 1024, 128, 0, 2>(tb, a0, b0, &mbar, ph, tid);
    r[13] = umma_time_one<64, 1, 1, 1152, 144, 1152, 144, 0, 1>(tb, a0, b0, &mbar, ph, tid);
    r[14] = umma_time_one<64, 0, 1, 128, 2048, 1152, 144, 0, 1>(tb, a0, b0, &mbar, ph, tid);    // segment sum of the edge forward
    r[15] = umma_time_one<64, 0, 0, 128, 1024, 128, 1024, 0, 1>(tb, a0, b0, &mbar, ph, tid);    // M = 64 K-major
    if (tid == 0)
        for (int i = 0; i < 16; ++i) out[i] = r[i];
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tb, 512);
}

}  // namespace is

using namespace is;

// out[16]: cycles per MMA for 16 operand-layout / accumulator-rotation configurations (see umma_timing_kernel)
extern "C" int is_umma_timing(float* out, void* stream) {
    const size_t smem = 96 * 1024 + 4096;
    cudaError_t e = cudaFuncSetAttribute(umma_timing_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    umma_timing_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(out);
    IS_LAUNCH_CHECK();
    return IS_OK;
}
extern "C" int is_umma_selftest(const float* A, const float* B, float* D, int mode, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e;
    if (mode == 0) {
        size_t smem = 24 * 8 * umma::kLBO + 128;
        e = cudaFuncSetAttribute(umma_selftest_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        umma_selftest_kernel<0><<<1, 128, smem, st>>>(A, B, D);
    } else if (mode == 1) {
        size_t smem = 24 * 16 * umma::kLBO + 128;
        e = cudaFuncSetAttribute(umma_selftest_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        umma_selftest_kernel<1><<<1, 128, smem, st>>>(A, B, D);
    } else if (mode == 2) {
        size_t smem = 2 * 24 * 16 * umma::kLBO + 128;
        e = cudaFuncSetAttribute(umma_selftest_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        umma_selftest_kernel<2><<<1, 128, smem, st>>>(A, B, D);
    } else if (mode == 3) {
        size_t smem = 3 * 24 * 8 * umma::kLBO + 128;
        e = cudaFuncSetAttribute(umma_selftest_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        umma_selftest_kernel<3><<<1, 128, smem, st>>>(A, B, D);
    } else if (mode >= 4 && mode <= 8) {
        size_t smem = (16 * 16 + 8 * 16) * umma::kLBO + 128;
        if (mode == 4) { cudaFuncSetAttribute(umma_layout_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); umma_layout_kernel<4><<<1, 128, smem, st>>>(A, B, D); }
        if (mode == 5) { cudaFuncSetAttribute(umma_layout_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); umma_layout_kernel<5><<<1, 128, smem, st>>>(A, B, D); }
        if (mode == 6) { cudaFuncSetAttribute(umma_layout_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); umma_layout_kernel<6><<<1, 128, smem, st>>>(A, B, D); }
        if (mode == 8) { cudaFuncSetAttribute(umma_layout_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); umma_layout_kernel<8><<<1, 128, smem, st>>>(A, B, D); }
        if (mode == 7) { cudaFuncSetAttribute(umma_layout_kernel<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); umma_layout_kernel<7><<<1, 128, smem, st>>>(A, B, D); }
    } else {
        return IS_ERR_ARG;
    }
    IS_LAUNCH_CHECK();
    return IS_OK;
}

