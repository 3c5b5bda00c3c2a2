// Segment (per-graph) mean / max / sum pooling over node rows, forward and backward.
//
// Reference: torch_geometric.nn.global_mean_pool / global_max_pool (PyG 2.5.3, un-vendored) at
// models/hybrid_models.py:97,331 and models/ablation_models.py:296-297 -- `scatter(x, batch, dim=0,
// dim_size=batch.max()+1, reduce='mean'|'max')`.  Segments are the contiguous node ranges of the batch
// (node_off), so no index vector and no atomics are needed: one CTA per graph, one thread per (row lane,
// column), rows combined in a fixed order (bit-reproducible).  Semantics restated from the published algorithm:
//   mean : segment sum / max(count, 1)        sum : segment sum        max : segment amax, 0 for an empty segment
//   max backward (scatter_reduce 'amax'): the gradient is split EVENLY among the rows that attain the maximum
//   (padded nodes of a graph produce identical rows, so ties are real).
// HBM-bound: reads X once (256 B / node at 64 columns), writes [B, C].
#include "common.cuh"

namespace is {

constexpr int SP_THREADS = 256;

// mode 0 = mean, 1 = max, 2 = sum.  grid = n_graphs, block = SP_THREADS (C <= 256 columns per pass)
__global__ void __launch_bounds__(SP_THREADS)
segment_pool_fwd_kernel(const float* __restrict__ X, int64_t ldx, int C, const int64_t* __restrict__ node_off,
                        int mode, float* __restrict__ out) {
    __shared__ float red[SP_THREADS];
    const int g = blockIdx.x;
    const int64_t r0 = node_off[g], r1 = node_off[g + 1];
    for (int c0 = 0; c0 < C; c0 += SP_THREADS) {
        const int cw = min(C - c0, SP_THREADS);
        const int rl = threadIdx.x / cw, c = threadIdx.x - rl * cw;
        const int nl = SP_THREADS / cw;                      // row lanes actually used for this pass
        float acc = mode == 1 ? -INFINITY : 0.0f;
        if (rl < nl) {
            for (int64_t r = r0 + rl; r < r1; r += nl) {
                const float v = __ldg(X + r * ldx + c0 + c);
                acc = mode == 1 ? fmaxf(acc, v) : acc + v;
            }
        }
        red[threadIdx.x] = acc;
        __syncthreads();
        if (threadIdx.x < cw) {
            float s = red[threadIdx.x];
            for (int l = 1; l < nl; ++l) {                   // ascending row-lane order: deterministic
                const float v = red[l * cw + threadIdx.x];
                s = mode == 1 ? fmaxf(s, v) : s + v;
            }
            const int64_t n = r1 - r0;
            if (mode == 0) s = s / (float)(n > 0 ? n : 1);
            if (mode == 1 && n == 0) s = 0.0f;
            out[(int64_t)g * C + c0 + threadIdx.x] = s;
        }
        __syncthreads();
    }
}

// gX[r][c] for the rows of graph g.  max: pooled = the forward output (the segment maxima).
__global__ void __launch_bounds__(SP_THREADS)
segment_pool_bwd_kernel(const float* __restrict__ X, int64_t ldx, int C, const int64_t* __restrict__ node_off,
                        int mode, const float* __restrict__ pooled, const float* __restrict__ g_out,
                        float* __restrict__ gX, int64_t ldg) {
    __shared__ int cnt[SP_THREADS];
    const int g = blockIdx.x;
    const int64_t r0 = node_off[g], r1 = node_off[g + 1];
    const int64_t n = r1 - r0;
    for (int c0 = 0; c0 < C; c0 += SP_THREADS) {
        const int cw = min(C - c0, SP_THREADS);
        const int rl = threadIdx.x / cw, c = threadIdx.x - rl * cw;
        const int nl = SP_THREADS / cw;
        const bool active = rl < nl;
        const float go = active ? __ldg(g_out + (int64_t)g * C + c0 + c) : 0.0f;
        if (mode != 1) {
            const float v = mode == 0 ? go / (float)(n > 0 ? n : 1) : go;
            if (active)
                for (int64_t r = r0 + rl; r < r1; r += nl) gX[r * ldg + c0 + c] = v;
            continue;
        }
        const float mx = active ? __ldg(pooled + (int64_t)g * C + c0 + c) : 0.0f;
        int k = 0;
        if (active)
            for (int64_t r = r0 + rl; r < r1; r += nl) k += (__ldg(X + r * ldx + c0 + c) == mx) ? 1 : 0;
        cnt[threadIdx.x] = k;
        __syncthreads();
        int total = 0;
        if (active)
            for (int l = 0; l < nl; ++l) total += cnt[l * cw + c];
        __syncthreads();
        if (active) {
            const float share = go / (float)max(total, 1);
            for (int64_t r = r0 + rl; r < r1; r += nl)
                gX[r * ldg + c0 + c] = (__ldg(X + r * ldx + c0 + c) == mx) ? share : 0.0f;
        }
    }
}

}  // namespace is

using namespace is;

extern "C" {

// out[g, :] = pool over rows node_off[g] .. node_off[g+1] of X [n, C] (row stride ldx floats); mode 0 mean, 1 max, 2 sum
int is_segment_pool_fwd(const float* X, int64_t ldx, int C, const int64_t* node_off, int n_graphs, int mode,
                        float* out, void* stream) {
    if (n_graphs < 0 || C <= 0 || mode < 0 || mode > 2) return IS_ERR_ARG;
    if (n_graphs == 0) return IS_OK;
    segment_pool_fwd_kernel<<<n_graphs, SP_THREADS, 0, (cudaStream_t)stream>>>(X, ldx, C, node_off, mode, out);
    IS_LAUNCH_CHECK();
    return IS_OK;
}

// gX [n, C] (row stride ldg) from g_out [B, C]; `pooled` = the forward output (needed for mode 1 only)
int is_segment_pool_bwd(const float* X, int64_t ldx, int C, const int64_t* node_off, int n_graphs, int mode,
                        const float* pooled, const float* g_out, float* gX, int64_t ldg, void* stream) {
    if (n_graphs < 0 || C <= 0 || mode < 0 || mode > 2 || (mode == 1 && (!pooled || !X))) return IS_ERR_ARG;
    if (n_graphs == 0) return IS_OK;
    segment_pool_bwd_kernel<<<n_graphs, SP_THREADS, 0, (cudaStream_t)stream>>>(X, ldx, C, node_off, mode, pooled, g_out, gX, ldg);
    IS_LAUNCH_CHECK();
    return IS_OK;
}

}  // extern "C"
