// Shared device helpers for the immunostruct_b200 kernels (sm_100a).
//
// Conventions used by every kernel in this directory:
//   * launchers are extern "C", take raw device pointers + sizes + a cudaStream_t (as void*),
//     never allocate, never synchronise, and return 0 or a negative argument error /
//     positive cudaError_t (see include/immunostruct_b200.h);
//   * all reductions are performed in a fixed order (no floating-point atomics), so results are
//     bit-reproducible run to run on the same device (the reference enables
//     torch.use_deterministic_algorithms, utils/seed.py:18);
//   * the SIMT fp32 register-tiled GEMM below computes a 128x64 output tile with 256 threads
//     (8 rows x 4 columns per thread); operands live in shared memory:  A row-major with a
//     padded leading dimension (multiple of 4, == 4 mod 32), B k-major [K][64].
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define IS_OK 0
#define IS_ERR_ARG (-1)
#define IS_ERR_UNSUPPORTED (-2)

#define IS_THREADS 256
#define IS_TM 128          // rows (edges / nodes) per tile
#define IS_H 64            // hidden width of every EGNN MLP (hybrid_models.py:247 gat_hidden_channels)
#define IS_LD 68           // padded leading dimension of a [*,64] shared-memory tile

#define IS_LAUNCH_CHECK()                                   \
    do {                                                    \
        cudaError_t e__ = cudaGetLastError();               \
        if (e__ != cudaSuccess) return (int)e__;            \
    } while (0)

namespace is {

// SM count of the CURRENT device (cached per device: the launch grids and the per-CTA partial buffers that the host
// sizes with is_egnn_*_grid must agree, also when several devices are driven from one process)
inline int current_num_sms() {
    static int cached[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (cached[dev] == 0) {
        int n = 0;
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        cached[dev] = n > 0 ? n : 148;
    }
    return cached[dev];
}

// e^x with ONE MUFU and the product-rounding term compensated: x log2(e) is formed as t + lo (t = the rounded product,
// lo = its exact rounding error + x * (log2 e - fl(log2 e))), 2^t comes from ex2.approx (2^-22.5 relative), and the
// factor 2^lo = 1 + lo ln 2 is applied afterwards.  Six instructions against ~11 for expf(); the plain ex2.approx(x *
// log2e) of the inference kernels carries |x| * 6e-8 of relative error from the product rounding, which was measured to
// push parameter gradients past 1e-5 -- this form does not.  IS_EXP_LIBM restores expf().
__device__ __forceinline__ float exp_comp(float x) {
#ifdef IS_EXP_LIBM
    return expf(x);
#else
    const float c1 = 1.4426950216293335f;              // fl(log2 e)
    const float c2 = 1.9259629911266175e-8f;           // log2 e - c1
    const float t = x * c1;
    float lo = fmaf(x, c1, -t);
    lo = fmaf(x, c2, lo);
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(t));
    return e * fmaf(lo, 0.6931471805599453f, 1.0f);      // (e may be 0 or inf: a product keeps them, e + e * x would not)
#endif
}

// accurate exp + correctly rounded reciprocal: the fast intrinsics (__expf, __fdividef) cost ~1e-5 of
// gradient parity against the fp32 reference (measured on the B200: tests/test_models_gpu.py)
__device__ __forceinline__ float sigmoidf_fast(float z) { return __frcp_rn(1.0f + expf(-z)); }
__device__ __forceinline__ float silu(float z) { return z * sigmoidf_fast(z); }
// d/dz [z * sigmoid(z)] = s * (1 + z * (1 - s))
__device__ __forceinline__ float dsilu(float z) {
    float s = sigmoidf_fast(z);
    return s * (1.0f + z * (1.0f - s));
}
__device__ __forceinline__ void silu_both(float z, float& y, float& dy) {
    float s = sigmoidf_fast(z);
    y = z * s;
    dy = s * (1.0f + z * (1.0f - s));
}

// 256-bit global accesses (sm_100: LDG.256 / STG.256).  One lane moves a whole 32-byte sector per instruction, so a
// warp-wide access whose lanes touch different rows is still sector complete (two 128-bit accesses per lane hit every
// sector twice, half a sector each).  p must be 32-byte aligned.
__device__ __forceinline__ void ldg256(const float* __restrict__ p, float (&v)[8]) {
    // Two 128-bit loads, NOT one ld.global.nc.v8.f32: builds in which ptxas allocated the 256-bit form (LDG.E.ENL2.256) with
    // its destination registers overlapping the address pair faulted on the B200 with "misaligned address" at that
    // instruction (compute-sanitizer: an address inside the right allocation, off by 48 / 120 bytes; egnn_tc2.cu gather,
    // only in some register allocations -- the same source with other flags ran clean); with 128-bit loads the same builds
    // are clean and no slower (the two halves of a lane's 32-byte sector are requested back to back).
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void stg256(float* __restrict__ p, const float (&v)[8]) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// sum over the 16 lanes that share the same ty (lanes differ in their low 4 bits)
__device__ __forceinline__ float half_warp_sum(float v) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- 128x64 register-tiled GEMM ------------------------------------------------------------
// acc[i][c] (+)= sum_k A[(ty + 16 i) * lda + k] * B[k * 64 + 4 tx + c],  i<8, c<4.
// Rows are interleaved (ty + 16 i) so that the two ty values of one warp hit different banks.
__device__ __forceinline__ void gemm_128x64(float (&acc)[8][4], const float* __restrict__ A, int lda,
                                            const float* __restrict__ B, int K, int ty, int tx) {
    const float* a0 = A + ty * lda;
    const float* b0 = B + 4 * tx;
#pragma unroll 2
    for (int k = 0; k < K; k += 4) {
        float4 w0 = *reinterpret_cast<const float4*>(b0 + (k + 0) * 64);
        float4 w1 = *reinterpret_cast<const float4*>(b0 + (k + 1) * 64);
        float4 w2 = *reinterpret_cast<const float4*>(b0 + (k + 2) * 64);
        float4 w3 = *reinterpret_cast<const float4*>(b0 + (k + 3) * 64);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float4 a = *reinterpret_cast<const float4*>(a0 + (16 * i) * lda + k);
            acc[i][0] = fmaf(a.x, w0.x, acc[i][0]); acc[i][1] = fmaf(a.x, w0.y, acc[i][1]);
            acc[i][2] = fmaf(a.x, w0.z, acc[i][2]); acc[i][3] = fmaf(a.x, w0.w, acc[i][3]);
            acc[i][0] = fmaf(a.y, w1.x, acc[i][0]); acc[i][1] = fmaf(a.y, w1.y, acc[i][1]);
            acc[i][2] = fmaf(a.y, w1.z, acc[i][2]); acc[i][3] = fmaf(a.y, w1.w, acc[i][3]);
            acc[i][0] = fmaf(a.z, w2.x, acc[i][0]); acc[i][1] = fmaf(a.z, w2.y, acc[i][1]);
            acc[i][2] = fmaf(a.z, w2.z, acc[i][2]); acc[i][3] = fmaf(a.z, w2.w, acc[i][3]);
            acc[i][0] = fmaf(a.w, w3.x, acc[i][0]); acc[i][1] = fmaf(a.w, w3.y, acc[i][1]);
            acc[i][2] = fmaf(a.w, w3.z, acc[i][2]); acc[i][3] = fmaf(a.w, w3.w, acc[i][3]);
        }
    }
}

__device__ __forceinline__ void zero_acc(float (&acc)[8][4]) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[i][c] = 0.0f;
}

// ---- weight-gradient tile: W[o][i] += sum_r G[r][o] * A[r][i] over the 128 rows of a tile ----
// 256 threads own a 64 x 64 block as 4x4 sub-blocks: o = 4*(tid/16)+a, i = 4*(tid%16)+b.
__device__ __forceinline__ void wgrad_64x64(float (&w)[4][4], const float* __restrict__ G, int ldg,
                                            const float* __restrict__ A, int lda, int tid) {
    const float* g0 = G + 4 * (tid >> 4);
    const float* a0 = A + 4 * (tid & 15);
#pragma unroll 4
    for (int r = 0; r < IS_TM; ++r) {
        float4 g = *reinterpret_cast<const float4*>(g0 + r * ldg);
        float4 a = *reinterpret_cast<const float4*>(a0 + r * lda);
        w[0][0] = fmaf(g.x, a.x, w[0][0]); w[0][1] = fmaf(g.x, a.y, w[0][1]);
        w[0][2] = fmaf(g.x, a.z, w[0][2]); w[0][3] = fmaf(g.x, a.w, w[0][3]);
        w[1][0] = fmaf(g.y, a.x, w[1][0]); w[1][1] = fmaf(g.y, a.y, w[1][1]);
        w[1][2] = fmaf(g.y, a.z, w[1][2]); w[1][3] = fmaf(g.y, a.w, w[1][3]);
        w[2][0] = fmaf(g.z, a.x, w[2][0]); w[2][1] = fmaf(g.z, a.y, w[2][1]);
        w[2][2] = fmaf(g.z, a.z, w[2][2]); w[2][3] = fmaf(g.z, a.w, w[2][3]);
        w[3][0] = fmaf(g.w, a.x, w[3][0]); w[3][1] = fmaf(g.w, a.y, w[3][1]);
        w[3][2] = fmaf(g.w, a.z, w[3][2]); w[3][3] = fmaf(g.w, a.w, w[3][3]);
    }
}

__device__ __forceinline__ void zero_w(float (&w)[4][4]) {
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) w[a][b] = 0.0f;
}

// store a thread's 4x4 weight-gradient block into a row-major [64][ncols] partial (cols < ncols)
__device__ __forceinline__ void store_w(float* __restrict__ dst, int ncols, int col0, const float (&w)[4][4],
                                        int tid) {
    int o = 4 * (tid >> 4), i = col0 + 4 * (tid & 15);
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b)
            if (i + b < ncols) dst[(o + a) * ncols + i + b] = w[a][b];
}

// ---- shared-memory staging -------------------------------------------------------------------
// Bt[k][o] = W[o * ldw + col0 + k]   (k < K, o < 64): k-major copy of a torch [out,in] weight block.
__device__ __forceinline__ void load_w_kmajor(float* __restrict__ Bt, const float* __restrict__ W, int ldw,
                                              int col0, int K, int tid) {
    for (int idx = tid; idx < K * 64; idx += IS_THREADS) {
        int o = idx & 63, k = idx >> 6;
        Bt[k * 64 + o] = __ldg(W + (size_t)o * ldw + col0 + k);
    }
}
// B[o][c] = W[o * ldw + col0 + c]   (o < 64 rows = reduction index, c < ncols <= 64; zero padded to 64)
__device__ __forceinline__ void load_w_rowmajor(float* __restrict__ B, const float* __restrict__ W, int ldw,
                                                int col0, int ncols, int tid) {
    for (int idx = tid; idx < 64 * 64; idx += IS_THREADS) {
        int c = idx & 63, o = idx >> 6;
        B[o * 64 + c] = (c < ncols) ? __ldg(W + (size_t)o * ldw + col0 + c) : 0.0f;
    }
}
// S[r][c] = X[(m0 + r) * ldx + c]  for r < 128, c < ncols ; rows >= M are zero filled.
__device__ __forceinline__ void load_rows(float* __restrict__ S, int lds, const float* __restrict__ X,
                                          int64_t ldx, int64_t m0, int64_t M, int ncols, int tid) {
    for (int idx = tid; idx < IS_TM * ncols; idx += IS_THREADS) {
        int r = idx / ncols, c = idx - r * ncols;
        int64_t m = m0 + r;
        S[r * lds + c] = (m < M) ? __ldg(X + m * ldx + c) : 0.0f;
    }
}

// column sums of a thread's accumulator-shaped values, reduced over the CTA deterministically:
// every thread holds v[4] (its 4 columns, already summed over its 8 rows); result[col] =
// sum over the 16 ty values in ascending ty order.  scratch: 16*64 floats.
__device__ __forceinline__ void cta_colsum_store(float* __restrict__ scratch, const float (&v)[4], int ty, int tx) {
#pragma unroll
    for (int c = 0; c < 4; ++c) scratch[ty * 64 + 4 * tx + c] = v[c];
}
__device__ __forceinline__ float cta_colsum_read(const float* __restrict__ scratch, int col) {
    float s = 0.0f;
#pragma unroll
    for (int t = 0; t < 16; ++t) s += scratch[t * 64 + col];
    return s;
}

}  // namespace is
