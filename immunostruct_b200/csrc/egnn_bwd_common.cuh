// Helpers shared by the tensor-core EGNN edge-backward kernels (egnn_bwd_tc.cu: lock-step, one tile at a time;
// egnn_bwd_ws.cu: two tile streams per CTA): fixed-order column sums, the accurate SiLU pair, bf16x3 GEMM issue.
#pragma once
#include "egnn_common.cuh"
#include "tc_common.cuh"

// tile rows (edges) of the two-stream kernel: four bf16x3 operand buffers of 112 rows + the weights fill shared memory
#define IS_BWD_WS_TR 112

namespace is {

// sum over the warp's 32 rows of 16 per-lane column values; afterwards lane L holds the total of column
// (L >> 1) & 15 (both lanes of a pair hold the same value).  Fixed order -> deterministic.
__device__ __forceinline__ float warp_colsum16(const float (&v)[16], int lane) {
    float w8[8], w4[4], w2[2];
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float send = b4 ? v[i] : v[i + 8];
        const float keep = b4 ? v[i + 8] : v[i];
        w8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float send = b3 ? w8[i] : w8[i + 4];
        const float keep = b3 ? w8[i + 4] : w8[i];
        w4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float send = b2 ? w4[i] : w4[i + 2];
        const float keep = b2 ? w4[i + 2] : w4[i];
        w2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    const float send = b1 ? w2[0] : w2[1];
    const float keep = b1 ? w2[1] : w2[0];
    float s = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    return s;
}

// SiLU and its derivative with the accurate expf and an approximate (1 ulp) reciprocal: the same arithmetic as
// the training forward (tc_common.cuh act<PREC, false>), ~10 instructions instead of ~20 with a rounded reciprocal.
__device__ __forceinline__ float sig_acc(float z) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + exp_comp(-z)));
    return r;
}
__device__ __forceinline__ float silu_acc(float z) { return z * sig_acc(z); }
__device__ __forceinline__ float dsilu_acc(float z) { const float s = sig_acc(z); return s * (1.0f + z * (1.0f - s)); }
__device__ __forceinline__ void silu_both_acc(float z, float& y, float& dy) {
    const float s = sig_acc(z);
    y = z * s;
    dy = s * (1.0f + z * (1.0f - s));
}

// ---- bf16x3 GEMM issue with explicit operand geometry (ONE thread) ------------------------------------------
struct OpGeom {
    uint32_t base;      // shared-memory address of split term 0
    uint32_t split;     // bytes between split terms
    uint32_t step;      // start-address advance per K step of 16
    uint32_t lbo, sbo;  // descriptor fields (bytes)
};
// t0 = first of the six partial products (0: all; timing experiments drop the small ones)
template <int T0 = 0>
__device__ __forceinline__ void issue_x3(uint32_t tmem_d, const OpGeom& a, const OpGeom& b, int nks, uint32_t idesc,
                                         uint32_t accumulate) {
    const uint32_t ta[6] = {2, 0, 1, 1, 0, 0}, tb[6] = {0, 2, 1, 0, 1, 0};     // smallest products first
    uint32_t acc = accumulate;
#pragma unroll
    for (int t = T0; t < 6; ++t)
#pragma unroll 8
        for (int ks = 0; ks < nks; ++ks) {
            mma_bf16(tmem_d, make_smem_desc(a.base + ta[t] * a.split + ks * a.step, a.lbo, a.sbo),
                     make_smem_desc(b.base + tb[t] * b.split + ks * b.step, b.lbo, b.sbo), idesc, acc);
            acc = 1;
        }
}

}  // namespace is
