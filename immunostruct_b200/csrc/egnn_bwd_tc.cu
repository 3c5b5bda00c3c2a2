// EGNN edge BACKWARD on the tensor cores (tcgen05 + TMEM), fp32-accurate (bf16x3 operand split).
//
// Same contract and outputs as is::edge_bwd_kernel (egnn.cu): per node-aligned tile of <= 128 in-edges the
// per-edge activations are recomputed and differentiated; gz1 ([E,64], CSR order), gQ, gD, gxd and one block
// of weight-gradient partials per CTA are written; no floating-point atomics.
// All six 128x64x64 GEMM-shaped products of a tile run as tcgen05.mma:
//
//   MMA 1  z2  = t1  W2^T          A = t1 tile (K-major)          B = W2 tile (K-major)
//   MMA 2  z3  = m   W3^T          A = m tile                     B = W3 tile
//   MMA 3  gm  = gz3 W3            A = gz3 tile                   B = W3 tile read MN-major (transposed view)
//   WG  3  gW3 += gz3^T m          A = gz3 tile read MN-major (M = 64)   B = m tile read MN-major
//   MMA 4  gt1 = gz2 W2            A = gz2 tile                   B = W2 tile read MN-major
//   WG  2  gW2 += gz2^T t1         A = gz2 tile read MN-major     B = t1 tile (re-gathered) read MN-major
//
// The transposed operands cost nothing: a K-major SWIZZLE_NONE tile read through a descriptor with the two
// strides swapped and the MN-major bit set IS its transpose (verified by is_umma_selftest modes 4 / 5).
// The weight gradients accumulate in TMEM (two M = 64 accumulators: row r of D sits in lane 32 (r/16) + r%16,
// probed by mode 6) across all tiles of the persistent CTA and are read out once at the end.  Bias-type
// gradients are column sums over edge rows: each warp reduce-scatters its 32 rows with 16 shuffles and keeps
// one running register per vector.
// One 512-thread CTA per SM; shared memory: two bf16x3 operand tiles (X, Y), W2 / W3, one fp32 tile.
#include "egnn_bwd_common.cuh"

namespace is {

constexpr int BT_NT = 512;
constexpr int BT_NW = BT_NT / 32;     // 16 warps
constexpr int BT_CW = 16;             // accumulator columns per thread (4 column quarters)
#define IS_EDGE_BWD_PARTIAL (4096 + 4096 + 5 * 64)

struct BwdMeta {                      // per-edge scalars of one tile (prefetched one tile ahead)
    int src[IS_TM];
    int dst[IS_TM];
    float r[IS_TM];
    float a[IS_TM];
    float inv[IS_TM];
    float gc[IS_TM];                  // dL/dc = v . dhat
    float dx[3 * IS_TM];              // raw difference x_src - x_dst
    float v[3 * IS_TM];               // gx_out[dst] / max(deg, 1)
};

template <bool HAS_COORD>
__device__ __forceinline__ void load_bwd_meta(const EdgeCommon& p, const float* __restrict__ gx_out, BwdMeta& m,
                                              int j, int p0, int ne) {
    int s = 0, d = 0;
    float r = 0.f, a = 0.f, inv = 0.f, gc = 0.f, dx = 0.f, dy = 0.f, dz = 0.f, v0 = 0.f, v1 = 0.f, v2 = 0.f;
    if (j < ne) {
        const int e = p0 + j;
        s = __ldg(p.csr_src + e); d = __ldg(p.csr_dst + e);
        a = __ldg(p.edge_attr + __ldg(p.csr_eid + e));
        dx = __ldg(p.x + s * p.ldx + 0) - __ldg(p.x + d * p.ldx + 0);
        dy = __ldg(p.x + s * p.ldx + 1) - __ldg(p.x + d * p.ldx + 1);
        dz = __ldg(p.x + s * p.ldx + 2) - __ldg(p.x + d * p.ldx + 2);
        r = dx * dx + dy * dy + dz * dz;
        inv = 1.0f / (sqrtf(r) + 1e-30f);
        if (HAS_COORD) {
            const int deg = __ldg(p.indptr + d + 1) - __ldg(p.indptr + d);
            const float sc = 1.0f / (float)max(deg, 1);
            v0 = __ldg(gx_out + (size_t)d * 3 + 0) * sc;
            v1 = __ldg(gx_out + (size_t)d * 3 + 1) * sc;
            v2 = __ldg(gx_out + (size_t)d * 3 + 2) * sc;
            gc = (v0 * dx + v1 * dy + v2 * dz) * inv;
        }
    }
    m.src[j] = s; m.dst[j] = d; m.r[j] = r; m.a[j] = a; m.inv[j] = inv; m.gc[j] = gc;
    m.dx[j] = dx; m.dx[IS_TM + j] = dy; m.dx[2 * IS_TM + j] = dz;
    m.v[j] = v0; m.v[IS_TM + j] = v1; m.v[2 * IS_TM + j] = v2;
}

template <bool HAS_COORD>
__global__ void __launch_bounds__(BT_NT, 1)
edge_bwd_tc_kernel(EdgeCommon p, const float* __restrict__ ghn, const float* __restrict__ gx_out,
                   float* __restrict__ gz1, float* __restrict__ gQ, float* __restrict__ gD, float* __restrict__ gxd,
                   float* __restrict__ partials, const int* __restrict__ gate) {
    // gate (device, optional) = the batch's maximum in-degree: when the two-stream kernel of egnn_bwd_ws.cu can take
    // the batch (every node fits one of its smaller tiles) this launch is its idle fallback and returns at once
    if (gate != nullptr && __ldg(gate) <= IS_BWD_WS_TR) return;
    using C = TcCfg<PREC_BF16X3>;
    constexpr uint32_t ASPL = C::A_BYTES, WSPL = C::W_BYTES;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint8_t* sX = smem_raw;                                   // [3][A_BYTES]  t1 -> gz3 -> gz2
    uint8_t* sY = sX + 3 * ASPL;                              // [3][A_BYTES]  m  -> t1 (re-gathered)
    uint8_t* sW2 = sY + 3 * ASPL;                             // [3][W_BYTES]
    uint8_t* sW3 = sW2 + 3 * WSPL;
    float* F32 = reinterpret_cast<float*>(sW3 + 3 * WSPL);    // [128][68] gz1 rows (dst-side sums)
    float* vec = F32 + IS_TM * IS_LD;                         // b2, b3, w4, wr, wa
    float* e_c = vec + 5 * 64;                                // [4][128] partial c per column quarter
    float* e_gr = e_c + 4 * IS_TM;                            // [4][128] partial gr per column quarter
    float* e_gd = e_gr + 4 * IS_TM;                           // [3][128]
    float* red = e_gd + 3 * IS_TM;                            // [16 warps][16] final vector reductions
    BwdMeta* meta = reinterpret_cast<BwdMeta*>(red + BT_NW * 16);   // [2]
    __shared__ int s_tile[2][4];
    __shared__ __align__(8) uint64_t mbar;       // data path: z2, z3, gm, gt1
    __shared__ __align__(8) uint64_t mbar_wg;    // weight-gradient MMAs (only their operand tiles are waited for)
    __shared__ uint32_t s_tmem;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ldw1 = 2 * p.F + 2;
    if (warp == 0) tmem_alloc(&s_tmem, 512);
    if (tid == 32) { mbar_init(&mbar, 1); mbar_init(&mbar_wg, 1); }
    for (int idx = tid; idx < 64 * 64; idx += BT_NT) {
        const int n = idx >> 6, k = idx & 63;
        store_weight1<PREC_BF16X3>(sW2, WSPL, n, k, __ldg(p.W2 + idx));
        store_weight1<PREC_BF16X3>(sW3, WSPL, n, k, HAS_COORD ? __ldg(p.W3 + idx) : 0.0f);
    }
    if (tid < 64) {
        vec[tid] = p.b2[tid];
        vec[64 + tid] = HAS_COORD ? p.b3[tid] : 0.0f;
        vec[128 + tid] = HAS_COORD ? p.w4[tid] : 0.0f;
        vec[192 + tid] = p.W1[tid * ldw1 + 2 * p.F];
        vec[256 + tid] = p.W1[tid * ldw1 + 2 * p.F + 1];
    }
    const int chunk = (p.n_nodes + gridDim.x - 1) / gridDim.x;
    const int nbeg = blockIdx.x * chunk;
    const int nend = min(p.n_nodes, nbeg + chunk);
    if (warp >= BT_NW - 4) {
        int tn0, tn1, tp0, tne;
        next_tile(p.indptr, nbeg, nend, p.status, lane, tn0, tn1, tp0, tne);
        if (warp == BT_NW - 4 && lane == 0) { s_tile[0][0] = tn0; s_tile[0][1] = tn1; s_tile[0][2] = tp0; s_tile[0][3] = tne; }
        load_bwd_meta<HAS_COORD>(p, gx_out, meta[0], (warp - (BT_NW - 4)) * 32 + lane, tp0, tne);
    }
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = s_tmem;
    // TMEM columns: D1 [0,64) z2 | D2 [64,128) z3 | D3 [128,192) gm | D4 [192,256) gt1 | DW2 [256,320) | DW3 [320,384)
    const int q = warp & 3, cq = warp >> 2, erow = 32 * q + lane;
    const uint32_t t_lane = tmem + ((uint32_t)(32 * q) << 16) + BT_CW * cq;
    const int esub = lane >> 3, kc8 = lane & 7;
    const float4 wr0 = *reinterpret_cast<const float4*>(vec + 192 + 8 * kc8), wr1 = *reinterpret_cast<const float4*>(vec + 196 + 8 * kc8);
    const float4 wa0 = *reinterpret_cast<const float4*>(vec + 256 + 8 * kc8), wa1 = *reinterpret_cast<const float4*>(vec + 260 + 8 * kc8);
    // operand geometries
    const OpGeom gX = {smem_u32(sX), ASPL, 2 * kLBO, kLBO, C::SBO};                 // K-major activation tile
    const OpGeom gY = {smem_u32(sY), ASPL, 2 * kLBO, kLBO, C::SBO};
    const OpGeom gXt = {smem_u32(sX), ASPL, 2 * C::SBO, C::SBO, kLBO};              // same tile, transposed (MN-major)
    const OpGeom gYt = {smem_u32(sY), ASPL, 2 * C::SBO, C::SBO, kLBO};
    const OpGeom gW2 = {smem_u32(sW2), WSPL, 2 * kLBO_W, kLBO_W, C::SBO_W};          // forward weight operand
    const OpGeom gW3 = {smem_u32(sW3), WSPL, 2 * kLBO_W, kLBO_W, C::SBO_W};
    const OpGeom gW2t = {smem_u32(sW2), WSPL, 2 * C::SBO_W, C::SBO_W, kLBO_W};       // transposed view (dgrad)
    const OpGeom gW3t = {smem_u32(sW3), WSPL, 2 * C::SBO_W, C::SBO_W, kLBO_W};
    const uint32_t id_fwd = make_instr_desc(1u, 128, 64, 0, 0);
    const uint32_t id_dgrad = make_instr_desc(1u, 128, 64, 0, 1);
    const uint32_t id_wgrad = make_instr_desc(1u, 64, 64, 1, 1);

    float acc_gb2 = 0.f, acc_gb3 = 0.f, acc_gw4 = 0.f, acc_gwr = 0.f, acc_gwa = 0.f;   // column (lane>>1)&15 of this warp
    uint32_t phase = 0, phase_wg = 0, wg_started = 0, wg_pending = 0;
    int cur = 0;

    // t1 = silu(P[src] + Q[dst] + wr r + wa a) for the tile's 128 rows -> operand tile `dst_tile`
    // t1 = silu(P[src] + Q[dst] + wr r + wa a) for this thread's two (row, chunk) items -> operand tile `dst_tile`; the
    // sixteen values stay in registers (t1v) so that the second use of t1 (weight gradient of W2) is a re-store, not a
    // second gather + SiLU pass
    float t1v[2][8];
    auto gather_t1 = [&](uint8_t* dst_tile, const BwdMeta& mt, int ne, auto&& before_stores) {
        float pv[2][8], qv[2][8];
        float rr[2], aa[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int j = u * 4 * BT_NW + warp * 4 + esub;
            const bool valid = j < ne;
            const int s = valid ? mt.src[j] : 0, d = valid ? mt.dst[j] : 0;
            rr[u] = mt.r[j]; aa[u] = mt.a[j];
            ldg256(p.PQ + (size_t)s * 128 + 8 * kc8, pv[u]);               // 256-bit loads: whole sectors per lane
            ldg256(p.PQ + (size_t)d * 128 + 64 + 8 * kc8, qv[u]);
        }
        const float wr[8] = {wr0.x, wr0.y, wr0.z, wr0.w, wr1.x, wr1.y, wr1.z, wr1.w};
        const float wa[8] = {wa0.x, wa0.y, wa0.z, wa0.w, wa1.x, wa1.y, wa1.z, wa1.w};
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int j = u * 4 * BT_NW + warp * 4 + esub;
            float d1[8];                                               // silu'(z1): parked in the fp32 tile for epilogue 4
#pragma unroll
            for (int i = 0; i < 8; ++i) silu_both_acc(pv[u][i] + qv[u][i] + wr[i] * rr[u] + wa[i] * aa[u], t1v[u][i], d1[i]);
            float* dd = F32 + j * IS_LD + 8 * kc8;
            *reinterpret_cast<float4*>(dd) = make_float4(d1[0], d1[1], d1[2], d1[3]);
            *reinterpret_cast<float4*>(dd + 4) = make_float4(d1[4], d1[5], d1[6], d1[7]);
        }
        // the row gathers and the SiLU evaluations above ran while the previous tile's weight-gradient MMAs (WG 2) were
        // still reading the operand tiles; only the stores below have to wait for them
        before_stores();
#pragma unroll
        for (int u = 0; u < 2; ++u)
            store_operand8<PREC_BF16X3>(dst_tile, ASPL, u * 4 * BT_NW + warp * 4 + esub, kc8, t1v[u]);   // rows >= ne: finite, x 0
    };
    auto restore_t1 = [&](uint8_t* dst_tile) {
#pragma unroll
        for (int u = 0; u < 2; ++u) store_operand8<PREC_BF16X3>(dst_tile, ASPL, u * 4 * BT_NW + warp * 4 + esub, kc8, t1v[u]);
    };
    // publish smem operand writes, let thread 0 issue `fn` + commit, wait for completion
    auto run_mma = [&](auto&& fn) {
        fence_async_smem();
        fence_before_sync();
        __syncthreads();
        if (warp == 0) {              // one elected lane of the converged warp issues (bare UTCHMMA, no election loop)
            if (elect_one()) {
                fence_after_sync();
                fn();
                mma_commit(&mbar);
            }
            __syncwarp();
            mbar_wait(&mbar, phase);
        }
        phase ^= 1;
        __syncthreads();
        fence_after_sync();
    };

    // wait until the weight-gradient MMAs issued last have finished reading their operand tiles
    auto wait_wg = [&]() {
        if (wg_pending) {
            if (tid == 0) mbar_wait(&mbar_wg, phase_wg);
            phase_wg ^= 1;
            wg_pending = 0;
            __syncthreads();
            fence_after_sync();
        }
    };

    while (true) {
        const int n0 = s_tile[cur][0], n1 = s_tile[cur][1], p0 = s_tile[cur][2], ne = s_tile[cur][3];
        if (n0 >= nend) break;
        const BwdMeta& mt = meta[cur];
        const bool row_valid = erow < ne;
        __syncthreads();        // the previous tile's destination-side sums are done with the fp32 tile

        // ---- t1 -> X ; MMA 1: z2 = t1 W2^T ; meanwhile prefetch the next tile's scalars ------------
        gather_t1(sX, mt, ne, [&] { wait_wg(); });     // WG 2 of the previous tile still reads X and Y
        fence_async_smem();
        fence_before_sync();
        __syncthreads();
        if (warp == 0) {
            if (elect_one()) {
                fence_after_sync();
                issue_x3(tmem + 0, gX, gW2, 4, id_fwd, 0);
                mma_commit(&mbar);
            }
            __syncwarp();
        }
        if (warp >= BT_NW - 4) {
            int tn0, tn1, tp0, tne;
            next_tile(p.indptr, n1, nend, p.status, lane, tn0, tn1, tp0, tne);
            if (warp == BT_NW - 4 && lane == 0) {
                s_tile[cur ^ 1][0] = tn0; s_tile[cur ^ 1][1] = tn1; s_tile[cur ^ 1][2] = tp0; s_tile[cur ^ 1][3] = tne;
            }
            load_bwd_meta<HAS_COORD>(p, gx_out, meta[cur ^ 1], (warp - (BT_NW - 4)) * 32 + lane, tp0, tne);
        }
        if (tid == 0) mbar_wait(&mbar, phase);
        phase ^= 1;
        __syncthreads();
        fence_after_sync();

        // ---- epilogue 1: m = silu(z2) -> Y, d2 = silu'(z2) stays in registers ---------------------------
        float d2[BT_CW];
        {
            float z[BT_CW];
            tmem_ld<BT_CW>(t_lane + 0, z);
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                float m8[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) silu_both_acc(z[8 * g + i] + vec[BT_CW * cq + 8 * g + i], m8[i], d2[8 * g + i]);
                if (HAS_COORD) store_operand8<PREC_BF16X3>(sY, ASPL, erow, 2 * cq + g, m8);
            }
        }
        float gm[BT_CW];
#pragma unroll
        for (int i = 0; i < BT_CW; ++i) gm[i] = 0.0f;
        if (HAS_COORD) {
            // ---- MMA 2: z3 = m W3^T ; epilogue 2: c, gz3 = gc w4 silu'(z3) -> X -----------------------------
            run_mma([&] { issue_x3(tmem + 64, gY, gW3, 4, id_fwd, 0); });
            {
                float z[BT_CW], g3[BT_CW], gu[BT_CW];
                tmem_ld<BT_CW>(t_lane + 64, z);
                const float gc = mt.gc[erow];
                float cpart = 0.0f;
#pragma unroll
                for (int i = 0; i < BT_CW; ++i) {
                    float u, d3;
                    silu_both_acc(z[i] + vec[64 + BT_CW * cq + i], u, d3);
                    const float w = vec[128 + BT_CW * cq + i];
                    cpart = fmaf(w, u, cpart);
                    g3[i] = gc * w * d3;            // gc == 0 on rows beyond the tile
                    gu[i] = gc * u;
                }
                e_c[cq * IS_TM + erow] = cpart;
                acc_gb3 += warp_colsum16(g3, lane);
                acc_gw4 += warp_colsum16(gu, lane);
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    float v8[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) v8[i] = g3[8 * g + i];
                    store_operand8<PREC_BF16X3>(sX, ASPL, erow, 2 * cq + g, v8);     // MMA 1 is done with X
                }
            }
            // ---- MMA 3: gm = gz3 W3 ; WG 3: gW3 += gz3^T m --------------------------------------------------
            fence_async_smem();
            fence_before_sync();
            __syncthreads();
            if (warp == 0) {
                if (elect_one()) {
                    fence_after_sync();
                    issue_x3(tmem + 128, gX, gW3t, 4, id_dgrad, 0);
                    mma_commit(&mbar);
                    issue_x3(tmem + 320, gXt, gYt, 8, id_wgrad, wg_started);
                    mma_commit(&mbar_wg);
                }
                __syncwarp();
                mbar_wait(&mbar, phase);
            }
            phase ^= 1;
            wg_pending = 1;
            __syncthreads();
            fence_after_sync();
            tmem_ld<BT_CW>(t_lane + 128, gm);
        }
        // ---- gz2 = (gm + ghn[dst]) silu'(z2) -> X ; t1 re-gathered -> Y ---------------------------------------
        {
            float g2[BT_CW];
            if (row_valid) {
                const float4* gh = reinterpret_cast<const float4*>(ghn + (size_t)mt.dst[erow] * 64 + BT_CW * cq);
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const float4 h4 = __ldg(gh + g);
                    g2[4 * g + 0] = (gm[4 * g + 0] + h4.x) * d2[4 * g + 0];
                    g2[4 * g + 1] = (gm[4 * g + 1] + h4.y) * d2[4 * g + 1];
                    g2[4 * g + 2] = (gm[4 * g + 2] + h4.z) * d2[4 * g + 2];
                    g2[4 * g + 3] = (gm[4 * g + 3] + h4.w) * d2[4 * g + 3];
                }
            } else {
#pragma unroll
                for (int i = 0; i < BT_CW; ++i) g2[i] = 0.0f;
            }
            acc_gb2 += warp_colsum16(g2, lane);
            wait_wg();          // WG 3 (if any) must be done with X (gz3) and Y (m) before they are overwritten
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                float v8[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) v8[i] = g2[8 * g + i];
                store_operand8<PREC_BF16X3>(sX, ASPL, erow, 2 * cq + g, v8);
            }
        }
        restore_t1(sY);
        // ---- MMA 4: gt1 = gz2 W2 ; WG 2: gW2 += gz2^T t1 -------------------------------------------------------
        fence_async_smem();
        fence_before_sync();
        __syncthreads();
        if (warp == 0) {
            if (elect_one()) {
                fence_after_sync();
                issue_x3(tmem + 192, gX, gW2t, 4, id_dgrad, 0);
                mma_commit(&mbar);
                issue_x3(tmem + 256, gXt, gYt, 8, id_wgrad, wg_started);
                mma_commit(&mbar_wg);
            }
            __syncwarp();
            mbar_wait(&mbar, phase);
        }
        phase ^= 1;
        wg_pending = 1;
        wg_started = 1;
        __syncthreads();
        fence_after_sync();
        // ---- epilogue 4: gz1 = gt1 silu'(z1) -> global + fp32 tile ; gr, gwr, gwa ------------------------------
        {
            float gt1[BT_CW], g1[BT_CW], gr_[BT_CW], ga_[BT_CW];
            tmem_ld<BT_CW>(t_lane + 192, gt1);
            float grpart = 0.0f;
            if (row_valid) {
                const float r = mt.r[erow], a = mt.a[erow];
                const float* dd = F32 + erow * IS_LD + BT_CW * cq;          // silu'(z1) of this row, written by the gather
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const float4 d4 = *reinterpret_cast<const float4*>(dd + 4 * g);
                    const float d1[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int c = BT_CW * cq + 4 * g + i;
                        const float wrc = vec[192 + c];
                        const float gz = gt1[4 * g + i] * d1[i];
                        g1[4 * g + i] = gz;
                        gr_[4 * g + i] = gz * r;
                        ga_[4 * g + i] = gz * a;
                        grpart = fmaf(wrc, gz, grpart);
                    }
                }
                float* go = gz1 + (size_t)(p0 + erow) * 64 + BT_CW * cq;
#pragma unroll
                for (int g = 0; g < 2; ++g) {                 // 256-bit stores: whole 32-byte sectors per lane
                    const float o[8] = {g1[8 * g], g1[8 * g + 1], g1[8 * g + 2], g1[8 * g + 3], g1[8 * g + 4], g1[8 * g + 5], g1[8 * g + 6], g1[8 * g + 7]};
                    stg256(go + 8 * g, o);
                }
            } else {
#pragma unroll
                for (int i = 0; i < BT_CW; ++i) { g1[i] = 0.0f; gr_[i] = 0.0f; ga_[i] = 0.0f; }
            }
            float* fo = F32 + erow * IS_LD + BT_CW * cq;
#pragma unroll
            for (int g = 0; g < 4; ++g) *reinterpret_cast<float4*>(fo + 4 * g) = make_float4(g1[4 * g], g1[4 * g + 1], g1[4 * g + 2], g1[4 * g + 3]);
            e_gr[cq * IS_TM + erow] = grpart;
            acc_gwr += warp_colsum16(gr_, lane);
            acc_gwa += warp_colsum16(ga_, lane);
        }
        fence_before_sync();
        __syncthreads();
        // ---- geometry backward: one thread per edge --------------------------------------------------------------
        if (tid < IS_TM) {
            const int j = tid;
            float g0 = 0.f, g1 = 0.f, g2 = 0.f;
            if (j < ne) {
                const float dx = mt.dx[j], dy = mt.dx[IS_TM + j], dz = mt.dx[2 * IS_TM + j];
                const float two_gr = 2.0f * (e_gr[j] + e_gr[IS_TM + j] + e_gr[2 * IS_TM + j] + e_gr[3 * IS_TM + j]);
                g0 = two_gr * dx; g1 = two_gr * dy; g2 = two_gr * dz;
                if (HAS_COORD) {
                    const float c = e_c[j] + e_c[IS_TM + j] + e_c[2 * IS_TM + j] + e_c[3 * IS_TM + j];
                    const float inv = mt.inv[j];
                    const float h0 = c * mt.v[j], h1 = c * mt.v[IS_TM + j], h2 = c * mt.v[2 * IS_TM + j];
                    const float rho = sqrtf(mt.r[j]);
                    const float k = (h0 * dx + h1 * dy + h2 * dz) * inv * inv / rho;
                    g0 += h0 * inv - k * dx; g1 += h1 * inv - k * dy; g2 += h2 * inv - k * dz;
                }
                gD[(size_t)(p0 + j) * 3 + 0] = g0;
                gD[(size_t)(p0 + j) * 3 + 1] = g1;
                gD[(size_t)(p0 + j) * 3 + 2] = g2;
            }
            e_gd[j] = g0; e_gd[IS_TM + j] = g1; e_gd[2 * IS_TM + j] = g2;
        }
        __syncthreads();
        // ---- destination-side sums: gQ[d] = sum gz1 ; gxd[d] = -sum g_diff ------------------------------------------
        for (int node = n0 + warp; node < n1; node += BT_NW) {
            const int jb = __ldg(p.indptr + node) - p0, je = __ldg(p.indptr + node + 1) - p0;
            float2 s = make_float2(0.0f, 0.0f);
            for (int j = jb; j < je; ++j) {
                const float2 v = *reinterpret_cast<const float2*>(F32 + j * IS_LD + 2 * lane);
                s.x += v.x; s.y += v.y;
            }
            *reinterpret_cast<float2*>(gQ + (size_t)node * 64 + 2 * lane) = s;
            if (lane < 3) {
                float sx = 0.0f;
                for (int j = jb; j < je; ++j) sx += e_gd[lane * IS_TM + j];
                gxd[(size_t)node * 3 + lane] = -sx;
            }
        }
        cur ^= 1;
        // the next iteration starts with wait_wg(): WG 2 is still reading X and Y; F32 / e_* are next written
        // after several barriers of the next iteration
    }

    // ---- per-CTA partials: weight gradients from TMEM, vector gradients from the running registers -----------
    wait_wg();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    float* P = partials + (size_t)blockIdx.x * IS_EDGE_BWD_PARTIAL;
    {
        float w[BT_CW];
        // M = 64 accumulators: row r lives in TMEM lane 32 (r / 16) + r % 16 -> this warp's lanes 0..15 hold rows 16 q + lane
        if (wg_started) tmem_ld<BT_CW>(t_lane + 256, w);
        if (lane < 16) {
#pragma unroll
            for (int i = 0; i < BT_CW; ++i) P[(16 * q + lane) * 64 + BT_CW * cq + i] = wg_started ? w[i] : 0.0f;
        }
        if (HAS_COORD && wg_started) tmem_ld<BT_CW>(t_lane + 320, w);
        if (lane < 16) {
#pragma unroll
            for (int i = 0; i < BT_CW; ++i) P[4096 + (16 * q + lane) * 64 + BT_CW * cq + i] = (HAS_COORD && wg_started) ? w[i] : 0.0f;
        }
    }
    const float accs[5] = {acc_gb2, acc_gb3, acc_gw4, acc_gwr, acc_gwa};
#pragma unroll
    for (int v = 0; v < 5; ++v) {
        __syncthreads();
        if ((lane & 1) == 0) red[warp * 16 + (lane >> 1)] = accs[v];
        __syncthreads();
        if (tid < 64) {
            const int cqq = tid >> 4, col = tid & 15;
            float s = 0.0f;
#pragma unroll
            for (int qq = 0; qq < 4; ++qq) s += red[(cqq * 4 + qq) * 16 + col];      // warps (q, cq): warp = 4 cq + q
            P[8192 + v * 64 + tid] = s;
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace is

// shared launcher: is_egnn_edge_bwd_tc (gate = nullptr) and the fallback leg of is_egnn_edge_bwd_ws (egnn_bwd_ws.cu)
namespace is {
int launch_edge_bwd_tc(const EdgeCommon& c, const float* ghn, const float* gx_out, float* gz1, float* gQ, float* gD,
                       float* gxd, float* partials, const int* gate, cudaStream_t st) {
    using C = TcCfg<PREC_BF16X3>;
    const size_t smem = 6 * (size_t)C::A_BYTES + 6 * (size_t)C::W_BYTES +
                        sizeof(float) * (IS_TM * IS_LD + 5 * 64 + 11 * IS_TM + BT_NW * 16) + 2 * sizeof(BwdMeta) + 128;
    const int sms = current_num_sms();
    int64_t g = ((int64_t)c.n_nodes + 31) / 32;
    if (g > sms) g = sms;
    const int grid = (int)(g < 1 ? 1 : g);
    cudaError_t e;
    if (gx_out) {
        e = cudaFuncSetAttribute(edge_bwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        edge_bwd_tc_kernel<true><<<grid, BT_NT, smem, st>>>(c, ghn, gx_out, gz1, gQ, gD, gxd, partials, gate);
    } else {
        e = cudaFuncSetAttribute(edge_bwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        edge_bwd_tc_kernel<false><<<grid, BT_NT, smem, st>>>(c, ghn, gx_out, gz1, gQ, gD, gxd, partials, gate);
    }
    IS_LAUNCH_CHECK();
    return IS_OK;
}
}  // namespace is

using namespace is;

extern "C" {

// Tensor-core variant of is_egnn_edge_bwd (same outputs, same partial layout, grid = is_egnn_edge_bwd_grid).
int is_egnn_edge_bwd_tc(const int* indptr, const int* csr_src, const int* csr_dst, const int* csr_eid,
                        const float* PQ, const float* x, int64_t ldx, const float* edge_attr,
                        const float* W1, int F, const float* W2, const float* b2,
                        const float* W3, const float* b3, const float* w4,
                        const float* ghn, const float* gx_out,
                        float* gz1, float* gQ, float* gD, float* gxd, float* partials,
                        int64_t n_nodes, int* status, void* stream) {
    if (!(F == 20 || F == 64) || n_nodes <= 0 || n_nodes > 0x7fffffff) return IS_ERR_ARG;
    if (((reinterpret_cast<uintptr_t>(PQ) | reinterpret_cast<uintptr_t>(gz1)) & 31) != 0) return IS_ERR_ARG;   // 256-bit accesses
    EdgeCommon c;
    c.indptr = indptr; c.csr_src = csr_src; c.csr_dst = csr_dst; c.csr_eid = csr_eid;
    c.PQ = PQ; c.x = x; c.ldx = ldx; c.edge_attr = edge_attr; c.W1 = W1; c.F = F;
    c.W2 = W2; c.b2 = b2; c.W3 = W3; c.b3 = b3; c.w4 = w4; c.n_nodes = (int)n_nodes; c.status = status;
    return launch_edge_bwd_tc(c, ghn, gx_out, gz1, gQ, gD, gxd, partials, nullptr, (cudaStream_t)stream);
}

}  // extern "C"
