// Helpers shared by the tcgen05 kernels (egnn_tc.cu, egnn_node_tc.cu): precision configurations,
// SiLU variants, operand staging in the canonical K-major layout, GEMM issue, TMEM loads.
//
// Precisions
//   PREC_BF16   (0): bf16 operands, fp32 accumulate, hardware-tanh SiLU        -- bf16 mode (1e-2)
//   PREC_TF32X3 (2): a = hi + lo in tf32, D = lo*Bhi + hi*Blo + hi*Bhi          -- fp32 parity, 8 B/element
//   PREC_BF16X3 (3): a = a1 + a2 + a3 in bf16, six partial products             -- fp32 parity, 6 B/element
//   PREC_FP16X2 (4): a = hi + lo in fp16 (11 + 11 mantissa bits), D = lo*Bhi + hi*Blo + hi*Bhi -- fp32 parity (2^-22), 4 B/element:
//                    two thirds of the operand bytes and HALF the MMAs of bf16x3, and a 3-instruction split instead of 5.5.
//                    fp16 has 5 exponent bits: values beyond +-65504 become inf (-> NaN downstream, loudly), and a residual
//                    below 2^-14 is kept to 2^-25 ABSOLUTE -- fine for O(1) forward activations and weights, not for
//                    gradients: inference forward only ("fp16x2" mode of functional.set_precision).
#pragma once
#include <cuda_fp16.h>
#include "common.cuh"
#include "umma.cuh"

#define PREC_BF16 0
#define PREC_TF32X3 2
#define PREC_BF16X3 3
#define PREC_FP16X2 4

namespace is {

using namespace umma;

template <int PREC>
struct TcCfg {
    static constexpr int EB = PREC == PREC_TF32X3 ? 4 : 2;             // operand element bytes
    static constexpr int KCH = 64 * EB / 16;                          // 16-byte chunks per 64-wide K block
    static constexpr uint32_t SBO = KCH * kLBO;                       // activation tiles: bytes between 8-row groups
    static constexpr uint32_t SBO_W = KCH * kLBO_W;                   // weight tiles (unpadded)
    static constexpr uint32_t A_BYTES = 16 * SBO;                     // 128-row operand tile (one split term)
    static constexpr uint32_t W_BYTES = 8 * SBO_W;                    // 64-row operand tile (one split term)
    static constexpr int NSPLIT = PREC == PREC_BF16 ? 1 : (PREC == PREC_TF32X3 || PREC == PREC_FP16X2) ? 2 : 3;
    static constexpr uint32_t FMT = PREC == PREC_TF32X3 ? 2u : PREC == PREC_FP16X2 ? 0u : 1u;       // instruction-descriptor operand format: 0 f16, 1 bf16, 2 tf32
    static constexpr bool ACCURATE = PREC != PREC_BF16;
};

// SiLU.  bf16 mode: z*sigmoid(z) = h + h*tanh(h), h = z/2, with the hardware tanh (one MUFU, three
// instructions, 2^-11 relative error -- below the bf16 operand rounding that follows).  fp32-parity
// modes: z * rcp(1 + 2^(-z log2 e)) with ex2.approx / rcp.approx (2^-22 each); the rounding of the
// product z*log2(e) adds |z| * 6e-8 relative error to exp(-z), which only matters where sigmoid is far
// from saturation, i.e. |z| = O(1): total error ~4e-7 relative, five instructions.  That is enough for
// logits at 1e-5 (inference) but was measured to push a few parameter GRADIENTS to 1.1-1.4e-5 of their
// scale, so the training forward (FAST = false) keeps the accurate expf.
template <int PREC, bool FAST = true>
__device__ __forceinline__ float act(float z) {
    if (!TcCfg<PREC>::ACCURATE) {
        const float h = 0.5f * z;
        float t;
        asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
        return fmaf(h, t, h);
    }
    float e, r;
    if (FAST) asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(z * -1.4426950408889634f));
    else e = exp_comp(-z);  // training forward: gradient parity at 1e-5 needs an exp without the product-rounding term
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
    return z * r;
}

__device__ __forceinline__ uint4 pack8_bf16(const float (&v)[8]) {
    uint4 q;
    q.x = pack_bf16x2(v[0], v[1]); q.y = pack_bf16x2(v[2], v[3]);
    q.z = pack_bf16x2(v[4], v[5]); q.w = pack_bf16x2(v[6], v[7]);
    return q;
}

// fp16 hi / lo split of 8 consecutive K values: hi = rn(v), lo = rn(v - hi) (the residual is exact in fp32); one packed
// conversion per pair of values and term
__device__ __forceinline__ void split8_fp16(const float (&v)[8], uint4& hi, uint4& lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __half2 h2 = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
        const float2 hf = __half22float2(h2);
        const __half2 l2 = __floats2half2_rn(v[2 * i] - hf.x, v[2 * i + 1] - hf.y);
        h[i] = *reinterpret_cast<const uint32_t*>(&h2);
        l[i] = *reinterpret_cast<const uint32_t*>(&l2);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// the packed 16-byte bf16 chunks of 8 consecutive K values, one per split term (q[0..NSPLIT-1]): bf16 = round to nearest;
// bf16x3 = three terms by truncation, as store_chunk8 below
template <int PREC>
__device__ __forceinline__ void split_chunk8(const float (&v)[8], uint4 (&q)[3]) {
    static_assert(PREC == PREC_BF16 || PREC == PREC_BF16X3 || PREC == PREC_FP16X2, "16-bit operand tiles only");
    if (PREC == PREC_BF16) {
        q[0] = pack8_bf16(v);
    } else if (PREC == PREC_FP16X2) {
        split8_fp16(v, q[0], q[1]);
    } else {
        uint32_t q1[4], q2[4], q3[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t a1 = __float_as_uint(v[2 * i]) & 0xffff0000u, b1 = __float_as_uint(v[2 * i + 1]) & 0xffff0000u;
            const float ra = v[2 * i] - __uint_as_float(a1), rb = v[2 * i + 1] - __uint_as_float(b1);
            const uint32_t a2 = __float_as_uint(ra) & 0xffff0000u, b2 = __float_as_uint(rb) & 0xffff0000u;
            const float sa = ra - __uint_as_float(a2), sb = rb - __uint_as_float(b2);
            q1[i] = __byte_perm(a1, b1, 0x7632);
            q2[i] = __byte_perm(a2, b2, 0x7632);
            q3[i] = __byte_perm(__float_as_uint(sa), __float_as_uint(sb), 0x7632);
        }
        q[0] = make_uint4(q1[0], q1[1], q1[2], q1[3]);
        q[1] = make_uint4(q2[0], q2[1], q2[2], q2[3]);
        q[2] = make_uint4(q3[0], q3[1], q3[2], q3[3]);
    }
}

// store 8 consecutive K values (one 16-byte bf16 chunk per split term) at byte address p0 of split term 0;
// `split_bytes` = distance between the split-term copies of the tile.  bf16 / bf16x3 only.
template <int PREC>
__device__ __forceinline__ void store_chunk8(uint8_t* __restrict__ p0, uint32_t split_bytes, const float (&v)[8]) {
    static_assert(PREC == PREC_BF16 || PREC == PREC_BF16X3 || PREC == PREC_FP16X2, "16-bit operand tiles only");
    if (PREC == PREC_BF16) {
        *reinterpret_cast<uint4*>(p0) = pack8_bf16(v);
    } else if (PREC == PREC_FP16X2) {
        uint4 hi, lo;
        split8_fp16(v, hi, lo);
        *reinterpret_cast<uint4*>(p0) = hi;
        *reinterpret_cast<uint4*>(p0 + split_bytes) = lo;
    } else {
        // three bf16 terms by truncation (top 16 bits); every residual is exact in fp32, so the sum of
        // the terms differs from v only by the truncation of the last one (2^-24 relative)
        uint32_t q1[4], q2[4], q3[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t a1 = __float_as_uint(v[2 * i]) & 0xffff0000u, b1 = __float_as_uint(v[2 * i + 1]) & 0xffff0000u;
            const float ra = v[2 * i] - __uint_as_float(a1), rb = v[2 * i + 1] - __uint_as_float(b1);
            const uint32_t a2 = __float_as_uint(ra) & 0xffff0000u, b2 = __float_as_uint(rb) & 0xffff0000u;
            const float sa = ra - __uint_as_float(a2), sb = rb - __uint_as_float(b2);
            q1[i] = __byte_perm(a1, b1, 0x7632);          // {hi16(a1), hi16(b1)} = bf16x2(a, b)
            q2[i] = __byte_perm(a2, b2, 0x7632);
            q3[i] = __byte_perm(__float_as_uint(sa), __float_as_uint(sb), 0x7632);
        }
        *reinterpret_cast<uint4*>(p0) = make_uint4(q1[0], q1[1], q1[2], q1[3]);
        *reinterpret_cast<uint4*>(p0 + split_bytes) = make_uint4(q2[0], q2[1], q2[2], q2[3]);
        *reinterpret_cast<uint4*>(p0 + 2 * split_bytes) = make_uint4(q3[0], q3[1], q3[2], q3[3]);
    }
}

// store 8 consecutive K values (k = 8*kc8 .. 8*kc8+7 of a 64-wide K block) of operand row `row`;
// `split_bytes` = distance between the split-term copies of the tile
template <int PREC>
__device__ __forceinline__ void store_operand8(uint8_t* __restrict__ tile, uint32_t split_bytes, int row, int kc8,
                                               const float (&v)[8]) {
    using C = TcCfg<PREC>;
    const uint32_t rbase = (uint32_t)((row >> 3) * C::SBO + (row & 7) * 16);
    if constexpr (PREC != PREC_TF32X3) {
        store_chunk8<PREC>(tile + rbase + kc8 * kLBO, split_bytes, v);
    } else {
        float hi[8], lo[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { hi[i] = tf32_round(v[i]); lo[i] = tf32_round(v[i] - hi[i]); }
        uint8_t* p0 = tile + rbase + (2 * kc8) * kLBO;
        *reinterpret_cast<float4*>(p0) = make_float4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<float4*>(p0 + kLBO) = make_float4(hi[4], hi[5], hi[6], hi[7]);
        *reinterpret_cast<float4*>(p0 + split_bytes) = make_float4(lo[0], lo[1], lo[2], lo[3]);
        *reinterpret_cast<float4*>(p0 + split_bytes + kLBO) = make_float4(lo[4], lo[5], lo[6], lo[7]);
    }
}

// features (2*lane, 2*lane+1) of operand row `row`, reconstructed from the (split) bf16 operand tile:
// with LBO = 144 the 32 lanes of a warp hit 32 different banks
template <int PREC>
__device__ __forceinline__ float2 load_operand2(const uint8_t* __restrict__ tile, uint32_t split_bytes, int row, int lane) {
    using C = TcCfg<PREC>;
    static_assert(C::EB == 2, "bf16 operand tiles only");
    const uint8_t* p = tile + (uint32_t)((row >> 3) * C::SBO + (row & 7) * 16) + (lane >> 2) * kLBO + (lane & 3) * 4;
    float2 r = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p));
    if (PREC == PREC_BF16X3) {
        const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p + split_bytes));
        const float2 c = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p + 2 * split_bytes));
        r.x += b.x + c.x; r.y += b.y + c.y;
    }
    return r;
}

// stage one element of a weight block (row n = output feature, k = input feature within a 64-wide K block)
template <int PREC>
__device__ __forceinline__ void store_weight1(uint8_t* __restrict__ tile, uint32_t split_bytes, int n, int k, float w) {
    using C = TcCfg<PREC>;
    const uint32_t off = canon_off<C::EB>(n, k, C::KCH, kLBO_W);
    if (PREC == PREC_BF16) {
        *reinterpret_cast<__nv_bfloat16*>(tile + off) = __float2bfloat16_rn(w);
    } else if (PREC == PREC_BF16X3) {
        const __nv_bfloat16 w1 = __float2bfloat16_rn(w);
        const float r1 = w - __bfloat162float(w1);
        const __nv_bfloat16 w2 = __float2bfloat16_rn(r1);
        *reinterpret_cast<__nv_bfloat16*>(tile + off) = w1;
        *reinterpret_cast<__nv_bfloat16*>(tile + split_bytes + off) = w2;
        *reinterpret_cast<__nv_bfloat16*>(tile + 2 * split_bytes + off) = __float2bfloat16_rn(r1 - __bfloat162float(w2));
    } else {
        const float hi = tf32_round(w);
        *reinterpret_cast<float*>(tile + off) = hi;
        *reinterpret_cast<float*>(tile + split_bytes + off) = tf32_round(w - hi);
    }
}

// stage a whole 64 x 64 weight block (rows n = output feature, W[n * ldw + col0 + k], k < kvalid <= 64; zero beyond)
// with 16-byte chunk stores: 512 (row, 8-wide K chunk) items over `nthreads` threads; consecutive threads take
// consecutive rows of one chunk, so a warp's stores fill four 128-byte core matrices (conflict free).  Same split
// arithmetic as store_weight1 (round-to-nearest terms).  W == nullptr stages zeros.
template <int PREC>
__device__ __forceinline__ void stage_weight_block(uint8_t* __restrict__ tile, uint32_t split_bytes, const float* __restrict__ W,
                                                   int ldw, int col0, int kvalid, int tid, int nthreads) {
    static_assert(PREC == PREC_BF16 || PREC == PREC_BF16X3 || PREC == PREC_FP16X2, "16-bit weight tiles only");
    for (int idx = tid; idx < 64 * 8; idx += nthreads) {
        const int n = idx & 63, c = idx >> 6;
        float w[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) w[i] = 0.0f;
        if (W != nullptr) {
            const float* src = W + (size_t)n * ldw + col0 + 8 * c;
            if (8 * c + 8 <= kvalid && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(src)), b = __ldg(reinterpret_cast<const float4*>(src) + 1);
                w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) w[i] = (8 * c + i < kvalid) ? __ldg(src + i) : 0.0f;
            }
        }
        uint8_t* dst = tile + (uint32_t)((n >> 3) * (8 * kLBO_W) + c * kLBO_W + (n & 7) * 16);
        if (PREC == PREC_FP16X2) {
            uint4 hi, lo;
            split8_fp16(w, hi, lo);
            *reinterpret_cast<uint4*>(dst) = hi;
            *reinterpret_cast<uint4*>(dst + split_bytes) = lo;
            continue;
        }
        uint32_t q1[4], q2[4], q3[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const __nv_bfloat16 a1 = __float2bfloat16_rn(w[2 * i]), b1 = __float2bfloat16_rn(w[2 * i + 1]);
            q1[i] = (uint32_t)__bfloat16_as_ushort(a1) | ((uint32_t)__bfloat16_as_ushort(b1) << 16);
            if (PREC == PREC_BF16X3) {
                const float ra = w[2 * i] - __bfloat162float(a1), rb = w[2 * i + 1] - __bfloat162float(b1);
                const __nv_bfloat16 a2 = __float2bfloat16_rn(ra), b2 = __float2bfloat16_rn(rb);
                const __nv_bfloat16 a3 = __float2bfloat16_rn(ra - __bfloat162float(a2)), b3 = __float2bfloat16_rn(rb - __bfloat162float(b2));
                q2[i] = (uint32_t)__bfloat16_as_ushort(a2) | ((uint32_t)__bfloat16_as_ushort(b2) << 16);
                q3[i] = (uint32_t)__bfloat16_as_ushort(a3) | ((uint32_t)__bfloat16_as_ushort(b3) << 16);
            }
        }
        *reinterpret_cast<uint4*>(dst) = make_uint4(q1[0], q1[1], q1[2], q1[3]);
        if (PREC == PREC_BF16X3) {
            *reinterpret_cast<uint4*>(dst + split_bytes) = make_uint4(q2[0], q2[1], q2[2], q2[3]);
            *reinterpret_cast<uint4*>(dst + 2 * split_bytes) = make_uint4(q3[0], q3[1], q3[2], q3[3]);
        }
    }
}

// issue D[tmem_d] (+)= A * W^T over one 64-wide K block (called by ONE thread); N = 64 or 128 output
// columns (W tile rows); a_split / w_split = byte distance between split-term copies.
// FULL = true adds the two 2^-24 products a2 w3 + a3 w2 (eight partial products instead of six): the result is then
// limited by the fp32 accumulator only -- used by the training forward of the node kernel, whose outputs feed the
// gradient parity test at 1e-5.
template <int PREC, bool FULL = false>
__device__ __forceinline__ void issue_gemm(uint32_t tmem_d, uint32_t a_addr, uint32_t a_split, uint32_t w_addr,
                                           uint32_t w_split, uint32_t n_cols, uint32_t accumulate) {
    using C = TcCfg<PREC>;
    const uint32_t idesc = make_instr_desc(C::FMT, 128, n_cols);
    uint32_t acc = accumulate;
    if (PREC == PREC_BF16X3) {
        // (a-term, w-term), smallest products first: [a3w2 a2w3] a3w1 a1w3 a2w2 a2w1 a1w2 a1w1
        const uint32_t ta[8] = {2, 1, 2, 0, 1, 1, 0, 0}, tw[8] = {1, 2, 0, 2, 1, 0, 1, 0};
#pragma unroll
        for (int t = FULL ? 0 : 2; t < 8; ++t)
#pragma unroll
            for (int ks = 0; ks < C::KCH / 2; ++ks) {
                mma_bf16(tmem_d, make_smem_desc(a_addr + ta[t] * a_split + ks * 2 * kLBO, kLBO, C::SBO),
                         make_smem_desc(w_addr + tw[t] * w_split + ks * 2 * kLBO_W, kLBO_W, C::SBO_W), idesc, acc);
                acc = 1;
            }
    } else if (PREC == PREC_FP16X2) {
        // (a-term, w-term), smallest products first: [lo lo] lo hi, hi lo, hi hi
        const uint32_t ta[4] = {1, 1, 0, 0}, tw[4] = {1, 0, 1, 0};
#pragma unroll
        for (int t = FULL ? 0 : 1; t < 4; ++t)
#pragma unroll
            for (int ks = 0; ks < C::KCH / 2; ++ks) {
                mma_bf16(tmem_d, make_smem_desc(a_addr + ta[t] * a_split + ks * 2 * kLBO, kLBO, C::SBO),
                         make_smem_desc(w_addr + tw[t] * w_split + ks * 2 * kLBO_W, kLBO_W, C::SBO_W), idesc, acc);
                acc = 1;
            }
    } else {
#pragma unroll
        for (int term = (PREC == PREC_BF16 ? 2 : 0); term < 3; ++term) {        // lo*hi, hi*lo, hi*hi
            const uint32_t aoff = (term == 0) ? a_split : 0, woff = (term == 1) ? w_split : 0;
#pragma unroll
            for (int ks = 0; ks < C::KCH / 2; ++ks) {
                const uint64_t da = make_smem_desc(a_addr + aoff + ks * 2 * kLBO, kLBO, C::SBO);
                const uint64_t db = make_smem_desc(w_addr + woff + ks * 2 * kLBO_W, kLBO_W, C::SBO_W);
                if (PREC == PREC_BF16) mma_bf16(tmem_d, da, db, idesc, acc); else mma_tf32(tmem_d, da, db, idesc, acc);
                acc = 1;
            }
        }
    }
}

template <int N>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, float (&v)[N]);
template <>
__device__ __forceinline__ void tmem_ld<32>(uint32_t taddr, float (&v)[32]) { tmem_ld32(taddr, v); }
template <>
__device__ __forceinline__ void tmem_ld<16>(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

template <>
__device__ __forceinline__ void tmem_ld<8>(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

}  // namespace is
