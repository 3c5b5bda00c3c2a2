// On-device collation of a graph batch: the device-side replacement of ``dgl.batch`` as called by
// ``collate`` / ``collate_amino_acid`` (immunostruct/data/utils.py:160-176 / 178-196) plus the
// destination-sorted CSR that DGL builds lazily on first ``update_all`` and its CSC transpose
// (SURVEY.md section 8(a) row 1).  All outputs are integers and must be BIT-EXACT against the
// oracle (oracle/reference_ops.py: dgl_batch, batch_vector, csr_from_coo).
//
// Input  : graph-local COO endpoints concatenated graph-major (src_local/dst_local int64 [E]),
//          per-graph node and edge counts (int64 [B]).
// Output : node_off/edge_off int64 [B+1] (segment offsets), edge_index int64 [2,E] (global ids,
//          edge order preserved), batch int64 [N]; CSR by destination (stable in edge id):
//          indptr int32 [N+1], csr_src/csr_dst/csr_eid int32 [E]; CSC transpose by source:
//          outptr int32 [N+1], csc_pos int32 [E] (CSR position of every out-edge, stable in edge id);
//          stats int32 [4] = {max in-degree, #endpoints out of range, max nodes per graph, #empty graphs}.
#include "common.cuh"

namespace is {

// ---- one warp walks one graph (graphs with more than CB_CAP nodes): counters / cursors in global scratch --------
// globalise endpoints, histogram, scan, stable fill of CSR and CSC.  Integer arithmetic only.
__device__ void collate_graph_warp(const int64_t* __restrict__ src_local, const int64_t* __restrict__ dst_local,
                                   int64_t n0, int64_t n1, int64_t e0, int64_t e1, int g, bool last,
                                   int64_t* __restrict__ edge_index, int64_t E, int64_t* __restrict__ batch,
                                   int* __restrict__ indptr, int* __restrict__ csr_src, int* __restrict__ csr_dst,
                                   int* __restrict__ csr_eid, int* __restrict__ outptr, int* __restrict__ csc_pos,
                                   int* __restrict__ cur_in, int* __restrict__ cur_out, int* __restrict__ stats, int lane) {
    const int ng = (int)(n1 - n0);
    for (int64_t n = n0 + lane; n < n1; n += 32) { batch[n] = g; cur_in[n] = 0; cur_out[n] = 0; }
    __syncwarp();
    int bad = 0;
    for (int64_t e = e0 + lane; e < e1; e += 32) {
        int64_t s = src_local[e], d = dst_local[e];
        if (s < 0 || s >= ng || d < 0 || d >= ng) { bad++; s = 0; d = 0; }
        s += n0; d += n0;
        edge_index[e] = s;
        edge_index[E + e] = d;
        atomicAdd(cur_in + d, 1);
        atomicAdd(cur_out + s, 1);
    }
    if (bad) atomicAdd(stats + 1, bad);
    __syncwarp();
    int run_in = (int)e0, run_out = (int)e0, maxdeg = 0;
    for (int64_t nb = n0; nb < n1; nb += 32) {
        const int64_t n = nb + lane;
        const int di = (n < n1) ? cur_in[n] : 0, dout = (n < n1) ? cur_out[n] : 0;
        maxdeg = max(maxdeg, di);
        int si = di, so = dout;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int ti = __shfl_up_sync(0xffffffffu, si, o), to = __shfl_up_sync(0xffffffffu, so, o);
            if (lane >= o) { si += ti; so += to; }
        }
        if (n < n1) {
            indptr[n] = run_in + si - di; outptr[n] = run_out + so - dout;
            cur_in[n] = run_in + si - di; cur_out[n] = run_out + so - dout;
        }
        run_in += __shfl_sync(0xffffffffu, si, 31);
        run_out += __shfl_sync(0xffffffffu, so, 31);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) maxdeg = max(maxdeg, __shfl_xor_sync(0xffffffffu, maxdeg, o));
    if (lane == 0) {
        atomicMax(stats, maxdeg);
        if (last) { indptr[n1] = (int)e1; outptr[n1] = (int)e1; }
    }
    __syncwarp();
    for (int64_t eb = e0; eb < e1; eb += 32) {
        const int64_t e = eb + lane;
        const bool act = e < e1;
        int s = -1 - lane, d = -1 - lane;
        if (act) {
            int64_t sl = src_local[e], dl = dst_local[e];
            if (sl < 0 || sl >= ng || dl < 0 || dl >= ng) { sl = 0; dl = 0; }
            s = (int)(sl + n0); d = (int)(dl + n0);
        }
        const unsigned md = __match_any_sync(0xffffffffu, d), ms = __match_any_sync(0xffffffffu, s);
        const unsigned lt = (1u << lane) - 1u;
        int pos_csr = 0;
        if (act) {
            pos_csr = cur_in[d] + __popc(md & lt);
            csr_src[pos_csr] = s; csr_dst[pos_csr] = d; csr_eid[pos_csr] = (int)e;
        }
        __syncwarp();
        if (act && (md & lt) == 0) cur_in[d] += __popc(md);
        if (act) csc_pos[cur_out[s] + __popc(ms & lt)] = pos_csr;
        __syncwarp();
        if (act && (ms & lt) == 0) cur_out[s] += __popc(ms);
        __syncwarp();
    }
}

// ---- one CTA (CB_WARPS warps) per graph: block-level STABLE counting sort -----------------------------------------
// The graph's edge list is cut into CB_WARPS contiguous chunks (ascending edge id).  Every warp histograms its chunk
// into its OWN per-node counters in shared memory; a block scan turns the per-node totals into indptr / outptr and
// the per-warp counters into per-(warp, node) start cursors (node start + edges of that node in earlier chunks);
// then every warp fills its chunk 32 edges per step with __match_any_sync ranks.  Chunks are ordered by edge id and
// each warp's fill is stable, so the result is the stable sort (bit-exact against the single-warp walk) -- with an
// 8x shorter dependent chain per graph (8 steps instead of 63 at 2 000 edges) and 8x more warps in flight.
#define CB_WARPS 8
#define CB_CAP 512
__global__ void __launch_bounds__(32 * CB_WARPS)
collate_kernel(const int64_t* __restrict__ src_local, const int64_t* __restrict__ dst_local,
               const int64_t* __restrict__ node_counts, const int64_t* __restrict__ edge_counts, int B,
               int64_t* __restrict__ node_off, int64_t* __restrict__ edge_off,
               int64_t* __restrict__ edge_index /* [2,E] */, int64_t E, int64_t* __restrict__ batch,
               int* __restrict__ indptr, int* __restrict__ csr_src, int* __restrict__ csr_dst, int* __restrict__ csr_eid,
               int* __restrict__ outptr, int* __restrict__ csc_pos,
               int* __restrict__ g_cur_in, int* __restrict__ g_cur_out /* scratch int32 [N] each */,
               int* __restrict__ stats) {
    __shared__ int s_cnt[2][CB_WARPS][CB_CAP];          // [in | out][warp][local node]: counts, then cursors (32 KB)
    __shared__ int s_wsum[2][CB_WARPS];
    __shared__ int64_t s_off[2][CB_WARPS];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int g = blockIdx.x;
    // segment offsets of this graph = sums of the counts of the graphs before it (every CTA reduces its own prefix:
    // B is a few hundred to a few thousand, so this is cheaper than a separate scan launch)
    {
        int64_t sn = 0, se = 0;
        for (int i = tid; i < g; i += 32 * CB_WARPS) { sn += node_counts[i]; se += edge_counts[i]; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { sn += __shfl_xor_sync(0xffffffffu, sn, o); se += __shfl_xor_sync(0xffffffffu, se, o); }
        if (lane == 0) { s_off[0][w] = sn; s_off[1][w] = se; }
        __syncthreads();
    }
    int64_t n0 = 0, e0 = 0;
#pragma unroll
    for (int ww = 0; ww < CB_WARPS; ++ww) { n0 += s_off[0][ww]; e0 += s_off[1][ww]; }
    const int64_t n1 = n0 + node_counts[g], e1 = e0 + edge_counts[g];
    if (tid == 0) {
        node_off[g] = n0; edge_off[g] = e0;
        if (g == B - 1) { node_off[B] = n1; edge_off[B] = e1; }
    }
    const int ng = (int)(n1 - n0);
    if (tid == 0) {
        atomicMax(stats + 2, ng);
        if (ng == 0) atomicAdd(stats + 3, 1);
    }
    if (ng > CB_CAP) {                                  // large graph: single-warp walk over the global scratch
        if (w == 0)
            collate_graph_warp(src_local, dst_local, n0, n1, e0, e1, g, g == B - 1, edge_index, E, batch, indptr, csr_src,
                               csr_dst, csr_eid, outptr, csc_pos, g_cur_in, g_cur_out, stats, lane);
        return;
    }
    for (int n = tid; n < ng; n += 32 * CB_WARPS) batch[n0 + n] = g;
    for (int n = tid; n < ng; n += 32 * CB_WARPS)
#pragma unroll
        for (int kw = 0; kw < 2 * CB_WARPS; ++kw) (&s_cnt[0][0][0])[kw * CB_CAP + n] = 0;
    __syncthreads();
    // this warp's chunk (multiple of 32 edges so that every step is a full warp except the graph's last)
    const int64_t ne = e1 - e0;
    const int64_t per = ((ne + CB_WARPS - 1) / CB_WARPS + 31) / 32 * 32;
    const int64_t ws = min(e1, e0 + w * per), we = min(e1, ws + per);
    // phase 1: globalise + per-warp degree histograms
    int bad = 0;
    for (int64_t eb = ws; eb < we; eb += 128) {
        int64_t sl[4], dl[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t e = eb + 32 * u + lane;
            sl[u] = e < we ? src_local[e] : 0; dl[u] = e < we ? dst_local[e] : 0;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t e = eb + 32 * u + lane;
            if (e < we) {
                int64_t s = sl[u], d = dl[u];
                if (s < 0 || s >= ng || d < 0 || d >= ng) { bad++; s = 0; d = 0; }
                edge_index[e] = s + n0;
                edge_index[E + e] = d + n0;
                atomicAdd(&s_cnt[0][w][d], 1);
                atomicAdd(&s_cnt[1][w][s], 1);
            }
        }
    }
    if (bad) atomicAdd(stats + 1, bad);
    __syncthreads();
    // phase 2: per-node totals (two consecutive nodes per thread), block exclusive scan, cursors
    int tot[2][2], maxdeg = 0;                          // [in | out][node 2 tid, 2 tid + 1]
#pragma unroll
    for (int k = 0; k < 2; ++k)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int n = 2 * tid + j;
            int t = 0;
            if (n < ng) {
#pragma unroll
                for (int ww = 0; ww < CB_WARPS; ++ww) {           // counts -> exclusive prefix over the warps
                    const int c = s_cnt[k][ww][n];
                    s_cnt[k][ww][n] = t;
                    t += c;
                }
            }
            tot[k][j] = t;
        }
    maxdeg = max(tot[0][0], tot[0][1]);
    int incl[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        int v = tot[k][0] + tot[k][1];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += t;
        }
        incl[k] = v;
        if (lane == 31) s_wsum[k][w] = v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) maxdeg = max(maxdeg, __shfl_xor_sync(0xffffffffu, maxdeg, o));
    if (lane == 0) atomicMax(stats, maxdeg);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        int base = (int)e0;
        for (int ww = 0; ww < w; ++ww) base += s_wsum[k][ww];
        const int ex0 = base + incl[k] - tot[k][0] - tot[k][1];       // start of node 2 tid
        const int ex1 = ex0 + tot[k][0];
        int* ptr = k == 0 ? indptr : outptr;
        if (2 * tid < ng) ptr[n0 + 2 * tid] = ex0;
        if (2 * tid + 1 < ng) ptr[n0 + 2 * tid + 1] = ex1;
#pragma unroll
        for (int ww = 0; ww < CB_WARPS; ++ww) {
            if (2 * tid < ng) s_cnt[k][ww][2 * tid] += ex0;
            if (2 * tid + 1 < ng) s_cnt[k][ww][2 * tid + 1] += ex1;
        }
    }
    if (tid == 0 && g == B - 1) { indptr[n1] = (int)e1; outptr[n1] = (int)e1; }
    __syncthreads();
    // phase 3: stable fill of this warp's chunk, 32 edges per step in ascending edge id
    int* cur_in = s_cnt[0][w];
    int* cur_out = s_cnt[1][w];
    for (int64_t eb4 = ws; eb4 < we; eb4 += 128) {
        int64_t sl4[4], dl4[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t e = eb4 + 32 * u + lane;
            sl4[u] = e < we ? src_local[e] : 0; dl4[u] = e < we ? dst_local[e] : 0;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t eb = eb4 + 32 * u;
            if (eb >= we) break;                                  // warp-uniform
            const int64_t e = eb + lane;
            const bool act = e < we;
            int s = -1 - lane, d = -1 - lane;                     // local node ids; inactive lanes match nobody
            if (act) {
                int64_t sl = sl4[u], dl = dl4[u];
                if (sl < 0 || sl >= ng || dl < 0 || dl >= ng) { sl = 0; dl = 0; }
                s = (int)sl; d = (int)dl;
            }
            const unsigned md = __match_any_sync(0xffffffffu, d), ms = __match_any_sync(0xffffffffu, s);
            const unsigned lt = (1u << lane) - 1u;
            int pos_csr = 0;
            if (act) {
                pos_csr = cur_in[d] + __popc(md & lt);
                csr_src[pos_csr] = s + (int)n0; csr_dst[pos_csr] = d + (int)n0; csr_eid[pos_csr] = (int)e;
            }
            __syncwarp();
            if (act && (md & lt) == 0) cur_in[d] += __popc(md);       // group leader advances the cursor
            if (act) csc_pos[cur_out[s] + __popc(ms & lt)] = pos_csr;
            __syncwarp();
            if (act && (ms & lt) == 0) cur_out[s] += __popc(ms);
            __syncwarp();
        }
    }
}

}  // namespace is

using namespace is;

extern "C" {

// Replaces dgl.batch (data/utils.py:163,169-170) + DGL's lazy CSC build.  scratch: int32 [2*N].
int is_collate_csr(const int64_t* src_local, const int64_t* dst_local, const int64_t* node_counts,
                   const int64_t* edge_counts, int n_graphs, int64_t n_nodes, int64_t n_edges,
                   int64_t* node_off, int64_t* edge_off, int64_t* edge_index, int64_t* batch,
                   int* indptr, int* csr_src, int* csr_dst, int* csr_eid, int* outptr, int* csc_pos,
                   int* scratch, int* stats, void* stream) {
    if (n_graphs <= 0 || n_nodes < 0 || n_edges < 0 || n_nodes > 0x7fffffff || n_edges > 0x7fffffff) return IS_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(stats, 0, 4 * sizeof(int), st);
    if (e != cudaSuccess) return (int)e;
    collate_kernel<<<n_graphs, 32 * CB_WARPS, 0, st>>>(
        src_local, dst_local, node_counts, edge_counts, n_graphs, node_off, edge_off, edge_index, n_edges, batch, indptr, csr_src, csr_dst,
        csr_eid, outptr, csc_pos, scratch, scratch + n_nodes, stats);
    IS_LAUNCH_CHECK();
    return IS_OK;
}

}  // extern "C"
