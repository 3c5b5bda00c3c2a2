// EGNN edge BACKWARD on the tensor cores, TWO TILE STREAMS PER CTA (tcgen05 + TMEM, bf16x3 operand split, sm_100a).
//
// Same contract, outputs and per-CTA partial layout as is::edge_bwd_tc_kernel (egnn_bwd_tc.cu) and is::edge_bwd_kernel
// (egnn.cu).  The lock-step kernel spends 36 % of its warp time waiting: a tile is a strict chain
//     gather -> MMA 1 -> epilogue 1 -> MMA 2 -> epilogue 2 -> MMA 3 (+ WG 3) -> epilogue 3 -> MMA 4 (+ WG 2) -> epilogue 4
// and with one tile in flight the tensor pipe (192 MMAs per tile) and the SIMT epilogues strictly alternate (ncu:
// barrier stalls 5.4 of 13.4 cycles per issue, profiles/r02_edge_bwd_*).  Here the CTA runs two independent TEAMS of
// 8 warps; team T owns the tiles T, T + 2, ... of the CTA's node range and walks the same chain on its own operand
// buffers, accumulator columns and named barrier, so one team's MMAs run under the other team's epilogues and every
// latency inside a team (MMA completion, L2 gathers, the scalar prefetch chain) is covered by the other team's work.
//
// What makes two tiles fit one SM:
//   * tiles of <= 112 edges (IS_BWD_WS_TR): four bf16x3 operand buffers of 112 x 64 (42 KB each) + W2 / W3 (48 KB) =
//     216 KB.  The M = 128 data MMAs read 16 rows past a buffer (the next buffer: finite or not, those accumulator
//     rows are never read); the weight-gradient MMAs run over K = 112 edges exactly;
//   * a thread owns one tile row and 32 columns in EVERY phase (gather included), so silu'(z1) never passes through
//     shared memory: z1 is parked in the thread's own TMEM lane (64 spare columns per team) and t1 / silu'(z1) are
//     re-derived from it where needed (the second use of t1, as the B operand of the W2 gradient, is a re-store);
//   * the four data accumulators of a tile are consumed one after the other: they share ONE 64-column TMEM block;
//   * the fp32 copy of gz1 that the destination-side sums read is written over the team's X buffer once the last
//     weight-gradient MMA of the tile has completed;
//   * per-edge scalars are single-buffered per team (their load is covered by the other team).
// Nodes with more than 112 in-edges do not fit these tiles: the launcher is handed the batch's maximum in-degree on the
// DEVICE (GraphBatch.stats[0]) and enqueues this kernel and the lock-step kernel back to back; each returns at once
// when the batch is the other one's.
//
// TMEM (512 columns): ACC[team] 64 | DW2, DW3 [team] 2 x 64 (M = 64 weight-gradient accumulators) | PARK[team] 64.
#include <type_traits>

#include "egnn_bwd_common.cuh"

namespace is {

int launch_edge_bwd_tc(const EdgeCommon& c, const float* ghn, const float* gx_out, float* gz1, float* gQ, float* gD,
                       float* gxd, float* partials, const int* gate, cudaStream_t st);     // egnn_bwd_tc.cu

namespace bw {
constexpr int TR = IS_BWD_WS_TR;                    // edge rows per tile
constexpr uint32_t LBO = 128, SBO = 8 * LBO;        // unpadded K-major tiles (a thread stores its own row: conflict free)
constexpr uint32_t A_TERM = (TR / 8) * SBO;         // one split term of a 112-row operand tile (14 KB)
constexpr uint32_t A_BUF = 3 * A_TERM;
constexpr uint32_t W_TERM = 8 * 8 * umma::kLBO_W;   // one split term of a 64 x 64 weight tile (8 KB)
constexpr uint32_t W_BUF = 3 * W_TERM;
constexpr int F_LD = 68;                            // fp32 gz1 rows: padded leading dimension
constexpr uint32_t F32_OFF = 2 * SBO;               // the fp32 copy starts behind the 16 rows a neighbour's M = 128 MMAs read
constexpr uint32_t F32_BYTES = TR * F_LD * 4;
constexpr int NKW = TR / 16;                        // K steps of a weight-gradient MMA group
constexpr int MAXN = 32;                            // destination nodes per tile
constexpr uint32_t TM_ACC = 0, TM_DW2 = 128, TM_DW3 = 192, TM_PARK = 384;
static_assert(F32_OFF + F32_BYTES + 7 * TR * 4 <= A_BUF, "fp32 staging must fit the X buffer");
static_assert(TR % 16 == 0 && TR <= 128, "tile rows");

struct Meta {                   // per-edge scalars of the team's current tile
    int src[TR];
    int dst[TR];
    float r[TR];
    float a[TR];
    float gc[TR];               // dL/dc = v . dhat
    float dx[3 * TR];           // raw difference x_src - x_dst
    float vn[3 * MAXN];         // gx_out[node] / max(deg, 1) for the tile's destination nodes
    int nptr[MAXN + 1];         // indptr[n0 + i] - p0: first tile row of node i (the destination-side sums walk these)
};

template <int TT>
__device__ __forceinline__ void team_sync(int team) {
    asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "n"(TT) : "memory");
}

// sum over the warp's 32 rows of 8 per-lane column values; afterwards lane L holds the total of column (L >> 2) & 7
// (all four lanes of a quad hold the same value).  Fixed order -> deterministic.
__device__ __forceinline__ float warp_colsum8(const float (&v)[8], int lane) {
    float w4[4], w2[2];
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float send = b4 ? v[i] : v[i + 4];
        const float keep = b4 ? v[i + 4] : v[i];
        w4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float send = b3 ? w4[i] : w4[i + 2];
        const float keep = b3 ? w4[i + 2] : w4[i];
        w2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    const float send = b2 ? w2[0] : w2[1];
    const float keep = b2 ? w2[1] : w2[0];
    float s = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    return s;
}

// column chunks (8 features) per thread of column group g, and the group's first chunk, for a team of TWARPS warps
// (4 TMEM lane quarters x TWARPS / 4 column groups): 8 warps: 4 + 4 | 12 warps: 3 + 3 + 2 | 16 warps: 2 + 2 + 2 + 2
template <int TWARPS> __host__ __device__ constexpr int group_chunks(int g) { return TWARPS == 8 ? 4 : TWARPS == 16 ? 2 : (g < 2 ? 3 : 2); }
template <int TWARPS> __host__ __device__ constexpr int group_first(int g) { return TWARPS == 8 ? 4 * g : TWARPS == 16 ? 2 * g : 3 * g; }

// Optional timeline tracing (debug builds: bash scripts/build_debug_lib.sh; scripts/trace_edge_bwd.py): thread 0 of each
// team of CTA 0 records clock64() per (team, event, tile) into the buffer handed over by is_debug_set_trace_bwd.
#ifdef IS_TRACE
__device__ long long* g_trace_bwd = nullptr;
#define TRB(ev) do { if (trc && t == 0 && tile_idx < 64) trc[((team * 20) + (ev)) * 64 + tile_idx] = clock64(); } while (0)
#else
#define TRB(ev) do { } while (0)
#endif

#ifndef IS_EXP_WG_T0
#define IS_EXP_WG_T0 0          // timing experiment: first partial product of the weight-gradient MMAs (0 = all six)
#endif
#ifndef IS_EXP_DATA_T0
#define IS_EXP_DATA_T0 0        // same for the data MMAs
#endif
#ifndef IS_BW_WAIT_HINT_NS
#define IS_BW_WAIT_HINT_NS 1000
#endif
}  // namespace bw

template <bool HAS_COORD, int TWARPS>
__global__ void __launch_bounds__(2 * 32 * TWARPS, 1)
edge_bwd_ws_kernel(EdgeCommon p, const float* __restrict__ ghn, const float* __restrict__ gx_out,
                   float* __restrict__ gz1, float* __restrict__ gQ, float* __restrict__ gD, float* __restrict__ gxd,
                   float* __restrict__ partials, const int* __restrict__ gate) {
    using namespace bw;
    constexpr int TT = 32 * TWARPS, NT = 2 * TT;            // team threads, CTA threads
    constexpr int NP = TT / 64, RP = (TR + NP - 1) / NP;    // row blocks of the gwr / gwa pass
    if (gate != nullptr && __ldg(gate) > TR) return;        // the lock-step kernel (128-edge tiles) takes this batch
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint8_t* sBuf = smem_raw;                                 // [team][X, Y][A_BUF]
    uint8_t* sW2 = sBuf + 4 * A_BUF;                          // [3][W_TERM]
    uint8_t* sW3 = sW2 + W_BUF;
    float* vec = reinterpret_cast<float*>(sW3 + W_BUF);       // b2, b3, w4, wr, wa
    Meta* metas = reinterpret_cast<Meta*>(vec + 5 * 64);      // [team]
    __shared__ int s_tile[2][4];                              // the team's NEXT tile: n0, n1, p0, ne
    __shared__ int s_started[2];
    __shared__ __align__(8) uint64_t mbar_d[2];               // data MMAs (z2, z3, gm, gt1) of a team
    __shared__ __align__(8) uint64_t mbar_wg[2];              // weight-gradient MMAs of a team
    __shared__ __align__(8) uint64_t mbar_kick;               // team 0 -> team 1: start half a tile later
    __shared__ uint32_t s_tmem;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int team = tid >= TT ? 1 : 0, t = tid - team * TT, tw = t >> 5;
    const int q = tw & 3, grp = tw >> 2, erow = 32 * q + lane;     // TMEM lane quarter (= CTA warp % 4), column group, tile row
    const bool rowv = erow < TR;
    const int ldw1 = 2 * p.F + 2;
    if (warp == 0) tmem_alloc(&s_tmem, 512);
    if (tid == 32) { mbar_init(&mbar_d[0], 1); mbar_init(&mbar_d[1], 1); mbar_init(&mbar_wg[0], 1); mbar_init(&mbar_wg[1], 1); mbar_init(&mbar_kick, 1); }
    // every byte the M = 128 MMAs can read is a finite bf16 from the start (and stays one: the fp32 copy of gz1 keeps clear
    // of the first 16 rows of a buffer): accumulator rows beyond a tile (rows 112..127 come from the neighbouring
    // buffer) then hold finite garbage, which the zero factors of the epilogues annihilate
    for (int i = tid; i < (int)(4 * A_BUF / 16); i += NT) reinterpret_cast<uint4*>(sBuf)[i] = make_uint4(0u, 0u, 0u, 0u);
    stage_weight_block<PREC_BF16X3>(sW2, W_TERM, p.W2, 64, 0, 64, tid, NT);
    stage_weight_block<PREC_BF16X3>(sW3, W_TERM, HAS_COORD ? p.W3 : nullptr, 64, 0, 64, tid, NT);
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
    if (tw == TWARPS - 1) {                 // the team's first tile: tile `team` of the CTA's sequence
        int a0, a1, ap, ae;
        next_tile<TR>(p.indptr, nbeg, nend, p.status, lane, a0, a1, ap, ae);
        if (team == 1) next_tile<TR>(p.indptr, a1, nend, p.status, lane, a0, a1, ap, ae);
        if (lane == 0) { s_tile[team][0] = a0; s_tile[team][1] = a1; s_tile[team][2] = ap; s_tile[team][3] = ae; }
    }
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = s_tmem;

    float acc_gwr = 0.f, acc_gwa = 0.f;          // column t & 63, row block t >> 6 (fp32 staging pass)
    float acc_v[3][4];                           // gb2, gb3, gw4: lane L holds column (L >> 2) & 7 of chunk ch of this warp's rows
#pragma unroll
    for (int v = 0; v < 3; ++v)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc_v[v][c] = 0.0f;
    uint32_t started = 0;

    auto team_main = [&](auto nch_tag) {
    constexpr int NCH = decltype(nch_tag)::value;             // 8-column chunks of this thread
    const int kc0 = group_first<TWARPS>(grp);                 // its first chunk
    uint8_t* X = sBuf + (size_t)team * 2 * A_BUF;             // t1 -> gz3 -> gz2 ; afterwards the fp32 copy of gz1
    uint8_t* Y = X + A_BUF;                                   // m -> t1 (re-stored)
    uint8_t* T1 = HAS_COORD ? X : Y;                          // without the coordinate branch t1 can stay where it is gathered
    Meta& mt = metas[team];
    float* F32 = reinterpret_cast<float*>(X + F32_OFF);       // [TR][F_LD]; the first 2 KB of X always hold bf16 operand rows
    float* e_gr = reinterpret_cast<float*>(X + F32_OFF + F32_BYTES);    // [TWARPS / 4][TR] partial gr per column group (<= 4)
    float* e_c = e_gr + 4 * TR;                               // [<= 4][TR] partial c per column group
    float* e_gd = e_c + 4 * TR;                               // [3][TR]
    uint64_t* bar_d = &mbar_d[team];
    uint64_t* bar_wg = &mbar_wg[team];
    const uint32_t t_acc = tmem + ((uint32_t)(32 * q) << 16) + TM_ACC + 64 * team + 8 * kc0;
    const uint32_t t_park = tmem + ((uint32_t)(32 * q) << 16) + TM_PARK + 64 * team + 8 * kc0;
    const uint32_t d_acc = tmem + TM_ACC + 64 * team, d_w2 = tmem + TM_DW2 + 128 * team, d_w3 = tmem + TM_DW3 + 128 * team;
    uint8_t* rowX = X + (uint32_t)((erow >> 3) * SBO + (erow & 7) * 16) + kc0 * LBO;     // this thread's first chunk in X
    uint8_t* rowY = rowX + A_BUF;
    uint8_t* rowT1 = HAS_COORD ? rowX : rowY;
    // operand geometries
    const OpGeom gXk = {smem_u32(X), A_TERM, 2 * LBO, LBO, SBO};                   // K-major activation tile
    const OpGeom gYk = {smem_u32(Y), A_TERM, 2 * LBO, LBO, SBO};
    const OpGeom gT1k = {smem_u32(T1), A_TERM, 2 * LBO, LBO, SBO};
    const OpGeom gXt = {smem_u32(X), A_TERM, 2 * SBO, SBO, LBO};                   // same tile, transposed (MN-major)
    const OpGeom gYt = {smem_u32(Y), A_TERM, 2 * SBO, SBO, LBO};
    const OpGeom gW2k = {smem_u32(sW2), W_TERM, 2 * kLBO_W, kLBO_W, 8 * kLBO_W};   // forward weight operand
    const OpGeom gW3k = {smem_u32(sW3), W_TERM, 2 * kLBO_W, kLBO_W, 8 * kLBO_W};
    const OpGeom gW2t = {smem_u32(sW2), W_TERM, 2 * 8 * kLBO_W, 8 * kLBO_W, kLBO_W};   // transposed view (dgrad)
    const OpGeom gW3t = {smem_u32(sW3), W_TERM, 2 * 8 * kLBO_W, 8 * kLBO_W, kLBO_W};
    const uint32_t id_fwd = make_instr_desc(1u, 128, 64, 0, 0);
    const uint32_t id_dgrad = make_instr_desc(1u, 128, 64, 0, 1);
    const uint32_t id_wgrad = make_instr_desc(1u, 64, 64, 1, 1);
    uint32_t ph_d = 0, ph_wg = 0;
#ifdef IS_TRACE
    long long* trc = blockIdx.x == 0 ? bw::g_trace_bwd : nullptr;
    int tile_idx = -1;
#endif

    // one warp of the team polls the mbarrier, the others sleep at the team's hardware barrier
    auto wait_d = [&]() {
        if (tw == 0) mbar_wait_hint(bar_d, ph_d, IS_BW_WAIT_HINT_NS);
        ph_d ^= 1;
        team_sync<TT>(team);
        fence_after_sync();
    };
    auto wait_wg = [&]() {
        if (tw == 0) mbar_wait_hint(bar_wg, ph_wg, IS_BW_WAIT_HINT_NS);
        ph_wg ^= 1;
        team_sync<TT>(team);
        fence_after_sync();
    };
    // publish this team's operand stores, then one elected lane of team warp 0 issues `fn`
    auto publish_and_issue = [&](auto&& fn) {
        fence_async_smem();
        fence_before_sync();
        TRB(18);
        team_sync<TT>(team);
        TRB(19);
        if (tw == 0) {
            if (elect_one()) {
                fence_after_sync();
                fn();
            }
            __syncwarp();
        }
    };

    // The two teams must NOT run in phase (measured with -DIS_TRACE: started together they stay in lock-step, both in an
    // epilogue or both waiting for their MMAs, and nothing overlaps): team 1 starts when team 0 is half way through its
    // first tile; equal tile periods keep the offset.
    bool kicked = false;
    if (team == 1) {
        if (tw == 0) mbar_wait_hint(&mbar_kick, 0, IS_BW_WAIT_HINT_NS);
        team_sync<TT>(team);
    }
    // operands already published (a team barrier and fence_after_sync lie behind): the elected lane of team warp 0 issues
    auto issue_only = [&](auto&& fn) {
        if (tw == 0) {
            if (elect_one()) fn();
            __syncwarp();
        }
    };
    while (true) {
        const int n0 = s_tile[team][0], n1 = s_tile[team][1], p0 = s_tile[team][2], ne = s_tile[team][3];
        if (n0 >= nend) break;
#ifdef IS_TRACE
        ++tile_idx;
#endif
        TRB(0);
        // ---- per-edge scalars of this tile (the other team's work covers the load chain) ---------------------------
        if (t < TR) {
            const int j = t;
            int s = 0, d = 0;
            float r = 0.f, a = 0.f, dx = 0.f, dy = 0.f, dz = 0.f;
            if (j < ne) {
                const int e = p0 + j;
                s = __ldg(p.csr_src + e); d = __ldg(p.csr_dst + e);
                a = __ldg(p.edge_attr + __ldg(p.csr_eid + e));
                dx = __ldg(p.x + s * p.ldx + 0) - __ldg(p.x + d * p.ldx + 0);
                dy = __ldg(p.x + s * p.ldx + 1) - __ldg(p.x + d * p.ldx + 1);
                dz = __ldg(p.x + s * p.ldx + 2) - __ldg(p.x + d * p.ldx + 2);
                r = dx * dx + dy * dy + dz * dz;
            }
            mt.src[j] = s; mt.dst[j] = d; mt.r[j] = r; mt.a[j] = a;
            mt.dx[j] = dx; mt.dx[TR + j] = dy; mt.dx[2 * TR + j] = dz;
            if (j <= MAXN) mt.nptr[j] = n0 + j <= n1 ? __ldg(p.indptr + n0 + j) - p0 : ne;
        } else if (HAS_COORD && t >= 128 && t < 128 + 3 * MAXN) {
            const int i = t - 128, nl = i / 3, comp = i - 3 * nl, node = n0 + nl;
            float v = 0.0f;
            if (node < n1) {
                const int deg = __ldg(p.indptr + node + 1) - __ldg(p.indptr + node);
                v = __ldg(gx_out + (size_t)node * 3 + comp) * (1.0f / (float)max(deg, 1));
            }
            mt.vn[i] = v;
        }
        team_sync<TT>(team);
        if (HAS_COORD && t < TR) {
            float gc = 0.0f;
            if (t < ne) {
                const int dl = mt.dst[t] - n0;
                const float inv = 1.0f / (sqrtf(mt.r[t]) + 1e-30f);
                gc = (mt.vn[3 * dl] * mt.dx[t] + mt.vn[3 * dl + 1] * mt.dx[TR + t] + mt.vn[3 * dl + 2] * mt.dx[2 * TR + t]) * inv;
            }
            mt.gc[t] = gc;          // read in epilogue 2, several team barriers from here
        }
        const bool valid = erow < ne;
        TRB(1);

        // ---- gather: z1 = P[src] + Q[dst] + wr r + wa a -> parked in TMEM ; t1 = silu(z1) -> T1 ; MMA 1 ------------
        {
            float pv[NCH][8], qv[NCH][8];
            float rr = 0.f, aa = 0.f;
            if (valid) {
                const int s = mt.src[erow], d = mt.dst[erow];
                rr = mt.r[erow]; aa = mt.a[erow];
                const float* pp = p.PQ + (size_t)s * 128 + 8 * kc0;
                const float* qp = p.PQ + (size_t)d * 128 + 64 + 8 * kc0;
#pragma unroll
                for (int ch = 0; ch < NCH; ++ch) { ldg256(pp + 8 * ch, pv[ch]); ldg256(qp + 8 * ch, qv[ch]); }
            } else {
#pragma unroll
                for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
                    for (int i = 0; i < 8; ++i) { pv[ch][i] = 0.f; qv[ch][i] = 0.f; }
            }
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) {
                float z[8], v8[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int c = 8 * (kc0 + ch) + i;
                    z[i] = pv[ch][i] + qv[ch][i] + vec[192 + c] * rr + vec[256 + c] * aa;       // rows beyond the tile: 0
                    v8[i] = silu_acc(z[i]);                                                    // silu(0) = 0
                }
                tmem_st8(t_park + 8 * ch, z);
                if (rowv) store_chunk8<PREC_BF16X3>(rowT1 + ch * LBO, A_TERM, v8);
            }
            tmem_st_wait();
        }
        TRB(2);
        publish_and_issue([&] {
            issue_x3<IS_EXP_DATA_T0>(d_acc, gT1k, gW2k, 4, id_fwd, 0);
            mma_commit(bar_d);
        });
        TRB(3);
        if (tw == TWARPS - 1) {             // this team's next tile = the tile after the other team's next one
            int a0, a1, ap, ae;
            next_tile<TR>(p.indptr, n1, nend, p.status, lane, a0, a1, ap, ae);
            next_tile<TR>(p.indptr, a1, nend, p.status, lane, a0, a1, ap, ae);
            if (lane == 0) { s_tile[team][0] = a0; s_tile[team][1] = a1; s_tile[team][2] = ap; s_tile[team][3] = ae; }
        }
        wait_d();
        TRB(4);

        // ---- epilogue 1: m = silu(z2 + b2) -> Y ; d2 = silu'(z2 + b2) stays in registers ---------------------------
        float d2[NCH][8];
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
            float z[8], m8[8];
            tmem_ld<8>(t_acc + 8 * ch, z);
#pragma unroll
            for (int i = 0; i < 8; ++i) silu_both_acc(z[i] + vec[8 * (kc0 + ch) + i], m8[i], d2[ch][i]);
            if (!valid) {
#pragma unroll
                for (int i = 0; i < 8; ++i) d2[ch][i] = 0.0f;
            }
            if (HAS_COORD && rowv) store_chunk8<PREC_BF16X3>(rowY + ch * LBO, A_TERM, m8);
        }
        float cpart = 0.0f;
        TRB(5);
        if (HAS_COORD) {
            // ---- MMA 2: z3 = m W3^T ; epilogue 2: c, gz3 = gc w4 silu'(z3 + b3) -> X -------------------------------
            publish_and_issue([&] {
                issue_x3<IS_EXP_DATA_T0>(d_acc, gYk, gW3k, 4, id_fwd, 0);
                mma_commit(bar_d);
            });
            wait_d();
            TRB(6);
            const float gc = valid ? mt.gc[erow] : 0.0f;
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) {
                float z[8], g3[8], gu[8];
                tmem_ld<8>(t_acc + 8 * ch, z);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int c = 8 * (kc0 + ch) + i;
                    float u, d3;
                    silu_both_acc(z[i] + vec[64 + c], u, d3);
                    const float w = vec[128 + c];
                    cpart = fmaf(w, u, cpart);
                    g3[i] = gc * w * d3;            // gc == 0 on rows beyond the tile (whose accumulator rows are finite)
                    gu[i] = gc * u;
                }
                acc_v[1][ch] += warp_colsum8(g3, lane);
                acc_v[2][ch] += warp_colsum8(gu, lane);
                if (rowv) store_chunk8<PREC_BF16X3>(rowX + ch * LBO, A_TERM, g3);     // MMA 1 is done with X
            }
            TRB(7);
            // ---- MMA 3: gm = gz3 W3 ; WG 3: gW3 += gz3^T m ---------------------------------------------------------
            // (the weight-gradient MMAs are issued only AFTER the data MMAs have completed: the two teams run nearly in
            //  phase, and a team's 24 data MMAs must not queue behind the other team's 42 weight-gradient MMAs; the
            //  weight gradients then run under the next epilogue's arithmetic)
            publish_and_issue([&] {
                issue_x3<IS_EXP_DATA_T0>(d_acc, gXk, gW3t, 4, id_dgrad, 0);
                mma_commit(bar_d);
            });
            wait_d();
            issue_only([&] {
                issue_x3<IS_EXP_WG_T0>(d_w3, gXt, gYt, NKW, id_wgrad, started);
                mma_commit(bar_wg);
            });
            TRB(8);
        }
        if (team == 0 && !kicked) {
            if (t == 0) mbar_arrive(&mbar_kick);
            kicked = true;
        }
        // ---- epilogue 3: gz2 = (gm + ghn[dst]) silu'(z2) -> X ; t1 re-derived from the parked z1 -> Y ---------------
        float d1[NCH][8];
        {
            const float* ghrow = ghn + (size_t)(valid ? mt.dst[erow] : 0) * 64 + 8 * kc0;
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) {
                float gm[8], gh[8];
                if (HAS_COORD) {
                    tmem_ld<8>(t_acc + 8 * ch, gm);
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) gm[i] = 0.0f;
                }
                if (valid) {
                    ldg256(ghrow + 8 * ch, gh);
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) gh[i] = 0.0f;
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) d2[ch][i] = (gm[i] + gh[i]) * d2[ch][i];       // gz2 (0 on rows beyond the tile)
                acc_v[0][ch] += warp_colsum8(d2[ch], lane);
            }
            TRB(9);
            if (HAS_COORD) wait_wg();       // WG 3 must be done with X (gz3) and Y (m) before they are overwritten
            TRB(10);
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch)
                if (rowv) store_chunk8<PREC_BF16X3>(rowX + ch * LBO, A_TERM, d2[ch]);
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) {
                float z[8], v8[8];
                tmem_ld<8>(t_park + 8 * ch, z);
#pragma unroll
                for (int i = 0; i < 8; ++i) silu_both_acc(z[i], v8[i], d1[ch][i]);
                if (HAS_COORD && rowv) store_chunk8<PREC_BF16X3>(rowY + ch * LBO, A_TERM, v8);
            }
        }
        TRB(11);
        // ---- MMA 4: gt1 = gz2 W2 ; WG 2: gW2 += gz2^T t1 ---------------------------------------------------------------
        publish_and_issue([&] {
            issue_x3<IS_EXP_DATA_T0>(d_acc, gXk, gW2t, 4, id_dgrad, 0);
            mma_commit(bar_d);
        });
        wait_d();
        issue_only([&] {
            issue_x3<IS_EXP_WG_T0>(d_w2, gXt, gYt, NKW, id_wgrad, started);
            mma_commit(bar_wg);
        });
        started = 1;
        TRB(12);
        // ---- epilogue 4: gz1 = gt1 silu'(z1) -> global ; gr ; then (WG 2 done) the fp32 copy over X ------------------------
        {
            float grpart = 0.0f;
            float* go = gz1 + (size_t)(p0 + erow) * 64 + 8 * kc0;
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) {
                float gt1[8];
                tmem_ld<8>(t_acc + 8 * ch, gt1);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float gz = gt1[i] * d1[ch][i];      // rows ne..111: gz2 rows are zero, so gt1 == 0; rows >= 112 are never stored
                    d1[ch][i] = gz;
                    grpart = fmaf(vec[192 + 8 * (kc0 + ch) + i], gz, grpart);
                }
                if (valid) stg256(go + 8 * ch, d1[ch]);
            }
            TRB(13);
            wait_wg();                      // WG 2 has read X and Y: both are free
            TRB(14);
            if (rowv) {
                float* fo = F32 + erow * F_LD + 8 * kc0;
#pragma unroll
                for (int ch = 0; ch < NCH; ++ch) {
                    *reinterpret_cast<float4*>(fo + 8 * ch) = make_float4(d1[ch][0], d1[ch][1], d1[ch][2], d1[ch][3]);
                    *reinterpret_cast<float4*>(fo + 8 * ch + 4) = make_float4(d1[ch][4], d1[ch][5], d1[ch][6], d1[ch][7]);
                }
                e_gr[grp * TR + erow] = grpart;
                if (HAS_COORD) e_c[grp * TR + erow] = cpart;
            }
        }
        fence_before_sync();
        team_sync<TT>(team);
        TRB(15);
        // ---- geometry backward (one thread per edge) ; gwr / gwa from the fp32 copy (column t & 63, rows of block t >> 6) ---
        if (t < TR) {
            const int j = t;
            float g0 = 0.f, g1 = 0.f, g2 = 0.f;
            if (j < ne) {
                const float dx = mt.dx[j], dy = mt.dx[TR + j], dz = mt.dx[2 * TR + j];
                float grs = 0.0f, cs = 0.0f;
#pragma unroll
                for (int gg = 0; gg < TWARPS / 4; ++gg) { grs += e_gr[gg * TR + j]; if (HAS_COORD) cs += e_c[gg * TR + j]; }
                const float two_gr = 2.0f * grs;
                g0 = two_gr * dx; g1 = two_gr * dy; g2 = two_gr * dz;
                if (HAS_COORD) {
                    const float rho = sqrtf(mt.r[j]);
                    const float inv = 1.0f / (rho + 1e-30f);
                    const int dl = mt.dst[j] - n0;
                    const float h0 = cs * mt.vn[3 * dl], h1 = cs * mt.vn[3 * dl + 1], h2 = cs * mt.vn[3 * dl + 2];
                    const float k = (h0 * dx + h1 * dy + h2 * dz) * inv * inv / rho;
                    g0 += h0 * inv - k * dx; g1 += h1 * inv - k * dy; g2 += h2 * inv - k * dz;
                }
                gD[(size_t)(p0 + j) * 3 + 0] = g0;
                gD[(size_t)(p0 + j) * 3 + 1] = g1;
                gD[(size_t)(p0 + j) * 3 + 2] = g2;
            }
            e_gd[j] = g0; e_gd[TR + j] = g1; e_gd[2 * TR + j] = g2;
        }
        {
            const int c = t & 63, rb = (t >> 6) * RP, re = min(ne, rb + RP);
            for (int row = rb; row < re; ++row) {
                const float gz = F32[row * F_LD + c];
                acc_gwr = fmaf(gz, mt.r[row], acc_gwr);
                acc_gwa = fmaf(gz, mt.a[row], acc_gwa);
            }
        }
        team_sync<TT>(team);
        TRB(16);
        // ---- destination-side sums: gQ[d] = sum gz1 ; gxd[d] = -sum g_diff ----------------------------------------------
        for (int node = n0 + tw; node < n1; node += TWARPS) {
            const int jb = mt.nptr[node - n0], je = mt.nptr[node - n0 + 1];
            float2 s = make_float2(0.0f, 0.0f);
#pragma unroll 4
            for (int j = jb; j < je; ++j) {          // ascending row order, one accumulator: the sum order of the other kernels
                const float2 v = *reinterpret_cast<const float2*>(F32 + j * F_LD + 2 * lane);
                s.x += v.x; s.y += v.y;
            }
            *reinterpret_cast<float2*>(gQ + (size_t)node * 64 + 2 * lane) = s;
            if (lane < 3) {
                float sx = 0.0f;
                for (int j = jb; j < je; ++j) sx += e_gd[lane * TR + j];
                gxd[(size_t)node * 3 + lane] = -sx;
            }
        }
        team_sync<TT>(team);        // X, Y and the scalars are free for the team's next tile
        TRB(17);
    }
    if (team == 0 && !kicked && t == 0) mbar_arrive(&mbar_kick);      // no tile at all: release team 1
    };   // team_main
    if (group_chunks<TWARPS>(0) != group_chunks<TWARPS>(2) && grp == 2) team_main(std::integral_constant<int, group_chunks<TWARPS>(2)>{});
    else team_main(std::integral_constant<int, group_chunks<TWARPS>(0)>{});

    // ---- per-CTA partials: weight gradients from TMEM (both teams), vector gradients from the running registers -----
    if (t == 0) s_started[team] = (int)started;
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    float* P = partials + (size_t)blockIdx.x * (8192 + 5 * 64);
    if (warp < 16) {
        // M = 64 accumulators: row r lives in TMEM lane 32 (r / 16) + r % 16 -> lanes 0..15 of CTA warp w hold rows 16 (w & 3) + lane
        const int wq = warp & 3, part = warp >> 2;
        const bool st0 = s_started[0] != 0, st1 = s_started[1] != 0;
        const uint32_t ta = tmem + ((uint32_t)(32 * wq) << 16) + 16 * part;
        float w0[16], w1[16];
        tmem_ld<16>(ta + TM_DW2, w0);
        tmem_ld<16>(ta + TM_DW2 + 128, w1);
        if (lane < 16) {
#pragma unroll
            for (int i = 0; i < 16; ++i) P[(16 * wq + lane) * 64 + 16 * part + i] = (st0 ? w0[i] : 0.0f) + (st1 ? w1[i] : 0.0f);
        }
        if (HAS_COORD) {
            tmem_ld<16>(ta + TM_DW3, w0);
            tmem_ld<16>(ta + TM_DW3 + 128, w1);
        }
        if (lane < 16) {
#pragma unroll
            for (int i = 0; i < 16; ++i)
                P[4096 + (16 * wq + lane) * 64 + 16 * part + i] = HAS_COORD ? (st0 ? w0[i] : 0.0f) + (st1 ? w1[i] : 0.0f) : 0.0f;
        }
    }
    float* red = reinterpret_cast<float*>(sBuf);              // [3 vectors][2 teams][TWARPS][4 chunks][8]
    float* red2 = red + 3 * 2 * TWARPS * 32;                  // [2 vectors][2 teams][NP row blocks][64]
    if ((lane & 3) == 0) {
        const int k = lane >> 2;
#pragma unroll
        for (int v = 0; v < 3; ++v)
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) red[((v * 2 + team) * TWARPS + tw) * 32 + 8 * ch + k] = acc_v[v][ch];
    }
    red2[((0 * 2 + team) * NP + (t >> 6)) * 64 + (t & 63)] = acc_gwr;
    red2[((1 * 2 + team) * NP + (t >> 6)) * 64 + (t & 63)] = acc_gwa;
    __syncthreads();
    if (tid < 5 * 64) {
        const int v = tid >> 6, col = tid & 63;
        float s = 0.0f;
        if (v < 3) {
            const int kc = col >> 3, k = col & 7;
            int gg = 0;                                        // column group that owns chunk kc, and the chunk's index in it
#pragma unroll
            for (int g2 = 1; g2 < TWARPS / 4; ++g2) if (kc >= group_first<TWARPS>(g2)) gg = g2;
            const int ch = kc - group_first<TWARPS>(gg);
#pragma unroll
            for (int tm = 0; tm < 2; ++tm)
#pragma unroll
                for (int qq = 0; qq < 4; ++qq) s += red[((v * 2 + tm) * TWARPS + 4 * gg + qq) * 32 + 8 * ch + k];
        } else {
#pragma unroll
            for (int tm = 0; tm < 2; ++tm)
#pragma unroll
                for (int rb = 0; rb < NP; ++rb) s += red2[(((v - 3) * 2 + tm) * NP + rb) * 64 + col];
        }
        P[8192 + v * 64 + col] = s;
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// warps per team of the two-stream kernel: 8 (32 columns per thread, 128 registers), 12 (24 / 24 / 16 columns, 80
// registers) or 16 (16 columns, 64 registers); is_egnn_set_bwd_ws_warps
int g_bwd_ws_warps = 16;

template <bool HAS_COORD, int TWARPS>
static int launch_bwd_ws(const EdgeCommon& c, const float* ghn, const float* gx_out, float* gz1, float* gQ, float* gD,
                         float* gxd, float* partials, const int* gate, int grid, cudaStream_t st) {
    const size_t smem = 4 * (size_t)bw::A_BUF + 2 * (size_t)bw::W_BUF + sizeof(float) * 5 * 64 + 2 * sizeof(bw::Meta);
    cudaError_t e = cudaFuncSetAttribute(edge_bwd_ws_kernel<HAS_COORD, TWARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    edge_bwd_ws_kernel<HAS_COORD, TWARPS><<<grid, 2 * 32 * TWARPS, smem, st>>>(c, ghn, gx_out, gz1, gQ, gD, gxd, partials, gate);
    e = cudaGetLastError();
    return e == cudaSuccess ? 0 : (int)e;
}

}  // namespace is

using namespace is;

extern "C" {

// Two-stream tensor-core edge backward (same outputs and partial layout as is_egnn_edge_bwd / is_egnn_edge_bwd_tc).
// max_in_degree = DEVICE pointer to the batch's maximum in-degree (GraphBatch.stats); the 112-edge-tile kernel runs when
// it is <= 112, the 128-edge lock-step kernel otherwise (both are enqueued; the idle one returns at once).  NULL, or
// pointers that miss the 32-byte alignment of the 256-bit row accesses: lock-step kernel only.
int is_egnn_edge_bwd_ws(const int* indptr, const int* csr_src, const int* csr_dst, const int* csr_eid,
                        const float* PQ, const float* x, int64_t ldx, const float* edge_attr,
                        const float* W1, int F, const float* W2, const float* b2,
                        const float* W3, const float* b3, const float* w4,
                        const float* ghn, const float* gx_out,
                        float* gz1, float* gQ, float* gD, float* gxd, float* partials,
                        const int* max_in_degree, int64_t n_nodes, int* status, void* stream) {
    if (!(F == 20 || F == 64) || n_nodes <= 0 || n_nodes > 0x7fffffff) return IS_ERR_ARG;
    if (((reinterpret_cast<uintptr_t>(PQ) | reinterpret_cast<uintptr_t>(gz1)) & 31) != 0) return IS_ERR_ARG;   // 256-bit accesses
    EdgeCommon c;
    c.indptr = indptr; c.csr_src = csr_src; c.csr_dst = csr_dst; c.csr_eid = csr_eid;
    c.PQ = PQ; c.x = x; c.ldx = ldx; c.edge_attr = edge_attr; c.W1 = W1; c.F = F;
    c.W2 = W2; c.b2 = b2; c.W3 = W3; c.b3 = b3; c.w4 = w4; c.n_nodes = (int)n_nodes; c.status = status;
    cudaStream_t st = (cudaStream_t)stream;
    if (max_in_degree == nullptr || (reinterpret_cast<uintptr_t>(ghn) & 31) != 0)
        return launch_edge_bwd_tc(c, ghn, gx_out, gz1, gQ, gD, gxd, partials, nullptr, st);
    const int sms = current_num_sms();
    int64_t g = (n_nodes + 31) / 32;
    if (g > sms) g = sms;
    const int grid = (int)(g < 1 ? 1 : g);
    int rc;
    const int tw = g_bwd_ws_warps;
#define IS_BW_LAUNCH(HC)                                                                                                   \
    (tw == 8 ? launch_bwd_ws<HC, 8>(c, ghn, gx_out, gz1, gQ, gD, gxd, partials, max_in_degree, grid, st)                  \
             : tw == 16 ? launch_bwd_ws<HC, 16>(c, ghn, gx_out, gz1, gQ, gD, gxd, partials, max_in_degree, grid, st)       \
                        : launch_bwd_ws<HC, 12>(c, ghn, gx_out, gz1, gQ, gD, gxd, partials, max_in_degree, grid, st))
    rc = gx_out ? IS_BW_LAUNCH(true) : IS_BW_LAUNCH(false);
#undef IS_BW_LAUNCH
    if (rc != 0) return rc;
    return launch_edge_bwd_tc(c, ghn, gx_out, gz1, gQ, gD, gxd, partials, max_in_degree, st);
}

#ifdef IS_TRACE
// debug builds only: 2 teams x 20 events x 64 tiles of clock64() stamps written by CTA 0 of the two-stream edge backward
int is_debug_set_trace_bwd(long long* buf) {
    cudaError_t e = cudaMemcpyToSymbol(is::bw::g_trace_bwd, &buf, sizeof(buf));
    return e == cudaSuccess ? IS_OK : (int)e;
}
#endif

// warps per team of the two-stream edge backward: 8, 12 or 16 (default)
int is_egnn_set_bwd_ws_warps(int n) {
    if (n != 8 && n != 12 && n != 16) return IS_ERR_ARG;
    is::g_bwd_ws_warps = n;
    return IS_OK;
}

}  // extern "C"
