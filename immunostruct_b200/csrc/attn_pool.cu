// Per-graph dense self-attention over the node embeddings + global mean pooling, fused.
//
// Reference: ``SelfAttention.forward`` (immunostruct/models/layers.py:13-22),
// ``MultiHeadAttention.forward`` / ``ScaleDotProductAttention`` (layers.py:29-48,67-78) applied per
// graph after ``view(B, -1, 64)`` (models/hybrid_models.py:92-94 / :326-328) followed by
// ``global_mean_pool`` (hybrid_models.py:97 / :331; torch_geometric scatter-mean).
// Each graph is one softmax segment (the "segment softmax" of the north star): one CTA per graph,
// K and V of the graph staged in shared memory, one warp per query row, scores kept in registers.
// The Q/K/V projections and the output projection are ordinary Linear layers applied outside.
//
// Layout: QKV [N_total,192] = [Q | K | V] per node; heads split the 64 channels contiguously
// (layers.py:80-93).  Outputs: O [N_total,64] (concatenated heads, before w_concat), LSE
// [N_total,H] (log-sum-exp of the scaled scores, for the backward pass), pooled [B,64] = per-graph
// mean of O rows (mean pooling commutes with the affine w_concat), and optionally the attention
// weights [sum_g H*n_g*n_g] when ``return_attention`` is requested.  Padded nodes take part in
// both the softmax and the mean exactly as in the reference (no mask: hybrid_models.py:327).
#include "common.cuh"

#define IS_ATT_NMAX 256        // nodes per graph supported by the shared-memory staging
#define IS_ATT_LD 65

namespace is {

__global__ void __launch_bounds__(IS_THREADS)
attn_fwd_kernel(const float* __restrict__ QKV, const int64_t* __restrict__ node_off, int H, float scale,
                float* __restrict__ O, float* __restrict__ LSE, float* __restrict__ pooled,
                float* __restrict__ attn /* or null */, const int64_t* __restrict__ attn_off /* [B] or null */) {
    extern __shared__ __align__(16) float smem[];
    const int g = blockIdx.x;
    const int64_t n0 = node_off[g];
    const int n = (int)(node_off[g + 1] - n0);
    float* Ks = smem;                               // [n][65]
    float* Vs = Ks + IS_ATT_NMAX * IS_ATT_LD;       // [n][65]
    float* qrow = Vs + IS_ATT_NMAX * IS_ATT_LD;     // [8 warps][64]
    float* prow = qrow + 8 * 64;                    // [8 warps][NMAX]
    float* red = prow + 8 * IS_ATT_NMAX;            // [4][64]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int dh = 64 / H;
    for (int idx = tid; idx < n * 64; idx += IS_THREADS) {
        const int r = idx >> 6, c = idx & 63;
        Ks[r * IS_ATT_LD + c] = __ldg(QKV + (n0 + r) * 192 + 64 + c);
        Vs[r * IS_ATT_LD + c] = __ldg(QKV + (n0 + r) * 192 + 128 + c);
    }
    __syncthreads();
    float* q = qrow + warp * 64;
    float* pr = prow + warp * IS_ATT_NMAX;
    for (int i = warp; i < n; i += IS_THREADS / 32) {
        __syncwarp();
        q[lane] = __ldg(QKV + (n0 + i) * 192 + lane);
        q[lane + 32] = __ldg(QKV + (n0 + i) * 192 + 32 + lane);
        __syncwarp();
        for (int h = 0; h < H; ++h) {
            float s[IS_ATT_NMAX / 32];
            float mx = -INFINITY;
#pragma unroll
            for (int jj = 0; jj < IS_ATT_NMAX / 32; ++jj) {
                const int j = jj * 32 + lane;
                float a = -INFINITY;
                if (j < n) {
                    a = 0.0f;
                    const float* kr = Ks + j * IS_ATT_LD + h * dh;
                    const float* qh = q + h * dh;
                    for (int k = 0; k < dh; ++k) a = fmaf(qh[k], kr[k], a);
                    a *= scale;
                }
                s[jj] = a;
                mx = fmaxf(mx, a);
            }
            mx = warp_max(mx);
            float z = 0.0f;
#pragma unroll
            for (int jj = 0; jj < IS_ATT_NMAX / 32; ++jj) {
                const int j = jj * 32 + lane;
                const float e = (j < n) ? expf(s[jj] - mx) : 0.0f;
                s[jj] = e;
                z += e;
            }
            z = warp_sum(z);
            const float rz = 1.0f / z;
            __syncwarp();
#pragma unroll
            for (int jj = 0; jj < IS_ATT_NMAX / 32; ++jj) {
                const int j = jj * 32 + lane;
                if (j < n) {
                    const float pj = s[jj] * rz;
                    pr[j] = pj;
                    if (attn) attn[attn_off[g] + ((int64_t)h * n + i) * n + j] = pj;
                }
            }
            __syncwarp();
            if (lane == 0) LSE[(n0 + i) * H + h] = mx + logf(z);
            // O[i][h*dh + k] = sum_j p_j V[j][h*dh + k]; lanes over k (two k per lane when dh = 64)
            for (int k = lane; k < dh; k += 32) {
                float o = 0.0f;
                const float* vc = Vs + h * dh + k;
                for (int j = 0; j < n; ++j) o = fmaf(pr[j], vc[j * IS_ATT_LD], o);
                O[(n0 + i) * 64 + h * dh + k] = o;
            }
        }
    }
    __syncthreads();
    // mean pooling of the O rows this CTA just wrote (fixed order: 4 row quarters, then combine)
    {
        const int c = tid & 63, part = tid >> 6;
        const int per = (n + 3) / 4, r0 = min(n, part * per), r1 = min(n, r0 + per);
        float sacc = 0.0f;
        for (int r = r0; r < r1; ++r) sacc += O[(n0 + r) * 64 + c];
        red[part * 64 + c] = sacc;
    }
    __syncthreads();
    if (tid < 64) pooled[(int64_t)g * 64 + tid] = (red[tid] + red[64 + tid] + red[128 + tid] + red[192 + tid]) / (float)max(n, 1);
}

// Backward.  gO[i] = gO_full[i] (optional) + g_pooled[g] / n.  Two passes with recomputed scores:
// rows (gQ) then columns (gK, gV), both in fixed summation order.
__global__ void __launch_bounds__(IS_THREADS)
attn_bwd_kernel(const float* __restrict__ QKV, const float* __restrict__ O, const float* __restrict__ LSE,
                const int64_t* __restrict__ node_off, int H, float scale,
                const float* __restrict__ g_pooled /* [B,64] or null */, const float* __restrict__ gO_full /* or null */,
                float* __restrict__ gQKV) {
    extern __shared__ __align__(16) float smem[];
    const int g = blockIdx.x;
    const int64_t n0 = node_off[g];
    const int n = (int)(node_off[g + 1] - n0);
    float* S0 = smem;                               // pass A: K   ; pass B: Q
    float* S1 = S0 + IS_ATT_NMAX * IS_ATT_LD;       // pass A: V   ; pass B: gO
    float* row0 = S1 + IS_ATT_NMAX * IS_ATT_LD;     // [8][64]  pass A: q row  ; pass B: k row
    float* row1 = row0 + 8 * 64;                    // [8][64]  pass A: gO row ; pass B: v row
    float* prow = row1 + 8 * 64;                    // [8][NMAX] gs row / column
    float* prow2 = prow + 8 * IS_ATT_NMAX;          // [8][NMAX] p column (pass B)
    float* Dn = prow2 + 8 * IS_ATT_NMAX;            // [NMAX][8] D[i][h] = gO_i . O_i per head
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int dh = 64 / H;
    const float inv_n = 1.0f / (float)max(n, 1);

    // ---------------- pass A: rows ----------------
    for (int idx = tid; idx < n * 64; idx += IS_THREADS) {
        const int r = idx >> 6, c = idx & 63;
        S0[r * IS_ATT_LD + c] = __ldg(QKV + (n0 + r) * 192 + 64 + c);
        S1[r * IS_ATT_LD + c] = __ldg(QKV + (n0 + r) * 192 + 128 + c);
    }
    __syncthreads();
    float* q = row0 + warp * 64;
    float* go = row1 + warp * 64;
    float* pr = prow + warp * IS_ATT_NMAX;
    for (int i = warp; i < n; i += IS_THREADS / 32) {
        __syncwarp();
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int c = lane + 32 * half;
            q[c] = __ldg(QKV + (n0 + i) * 192 + c);
            float v = g_pooled ? __ldg(g_pooled + (int64_t)g * 64 + c) * inv_n : 0.0f;
            if (gO_full) v += __ldg(gO_full + (n0 + i) * 64 + c);
            go[c] = v;
        }
        __syncwarp();
        for (int h = 0; h < H; ++h) {
            // D = sum_k gO[i][k] O[i][k] over this head's channels
            float dpart = 0.0f;
            for (int k = lane; k < dh; k += 32) dpart += go[h * dh + k] * __ldg(O + (n0 + i) * 64 + h * dh + k);
            const float D = warp_sum(dpart);
            if (lane == 0) Dn[i * 8 + h] = D;
            const float lse = __ldg(LSE + (n0 + i) * H + h);
            __syncwarp();
            for (int j = lane; j < n; j += 32) {
                const float* kr = S0 + j * IS_ATT_LD + h * dh;
                const float* vr = S1 + j * IS_ATT_LD + h * dh;
                float a = 0.0f, gp = 0.0f;
                for (int k = 0; k < dh; ++k) { a = fmaf(q[h * dh + k], kr[k], a); gp = fmaf(go[h * dh + k], vr[k], gp); }
                const float pj = expf(a * scale - lse);
                pr[j] = pj * (gp - D);
            }
            __syncwarp();
            for (int k = lane; k < dh; k += 32) {
                float acc = 0.0f;
                const float* kc = S0 + h * dh + k;
                for (int j = 0; j < n; ++j) acc = fmaf(pr[j], kc[j * IS_ATT_LD], acc);
                gQKV[(n0 + i) * 192 + h * dh + k] = acc * scale;
            }
        }
    }
    __syncthreads();
    // ---------------- pass B: columns ----------------
    for (int idx = tid; idx < n * 64; idx += IS_THREADS) {
        const int r = idx >> 6, c = idx & 63;
        S0[r * IS_ATT_LD + c] = __ldg(QKV + (n0 + r) * 192 + c);
        float v = g_pooled ? __ldg(g_pooled + (int64_t)g * 64 + c) * inv_n : 0.0f;
        if (gO_full) v += __ldg(gO_full + (n0 + r) * 64 + c);
        S1[r * IS_ATT_LD + c] = v;
    }
    __syncthreads();
    float* kk = row0 + warp * 64;
    float* vv = row1 + warp * 64;
    float* pc = prow2 + warp * IS_ATT_NMAX;
    for (int j = warp; j < n; j += IS_THREADS / 32) {
        __syncwarp();
        kk[lane] = __ldg(QKV + (n0 + j) * 192 + 64 + lane);
        kk[lane + 32] = __ldg(QKV + (n0 + j) * 192 + 96 + lane);
        vv[lane] = __ldg(QKV + (n0 + j) * 192 + 128 + lane);
        vv[lane + 32] = __ldg(QKV + (n0 + j) * 192 + 160 + lane);
        __syncwarp();
        for (int h = 0; h < H; ++h) {
            for (int i = lane; i < n; i += 32) {
                const float* qr = S0 + i * IS_ATT_LD + h * dh;
                const float* gr = S1 + i * IS_ATT_LD + h * dh;
                float a = 0.0f, gp = 0.0f;
                for (int k = 0; k < dh; ++k) { a = fmaf(qr[k], kk[h * dh + k], a); gp = fmaf(gr[k], vv[h * dh + k], gp); }
                const float pij = expf(a * scale - __ldg(LSE + (n0 + i) * H + h));
                pc[i] = pij;
                pr[i] = pij * (gp - Dn[i * 8 + h]);
            }
            __syncwarp();
            for (int k = lane; k < dh; k += 32) {
                float gk = 0.0f, gv = 0.0f;
                const float* qc = S0 + h * dh + k;
                const float* gc = S1 + h * dh + k;
                for (int i = 0; i < n; ++i) {
                    gk = fmaf(pr[i], qc[i * IS_ATT_LD], gk);
                    gv = fmaf(pc[i], gc[i * IS_ATT_LD], gv);
                }
                gQKV[(n0 + j) * 192 + 64 + h * dh + k] = gk * scale;
                gQKV[(n0 + j) * 192 + 128 + h * dh + k] = gv;
            }
            __syncwarp();
        }
    }
}


// ---------------------------------------------------------------------------------------------------
// Inference fast path: only the pooled output is needed (no_grad, no attention weights).  Because the
// mean over rows commutes with the P V product,
//     pooled[h*dh + c] = sum_j w^h_j V[j][h*dh + c],   w^h_j = (1/n) sum_i softmax_j(q_i . k_j / sqrt(dh)),
// so only the column sums of the attention matrix are accumulated (in registers, per warp, fixed order)
// and the P V product disappears.  Q K^T is register blocked: each warp handles 4 query rows at a time,
// each lane up to 8 key columns -> 128 FMAs per 12 shared-memory loads.
#define IS_ATT_LD4 68
__global__ void __launch_bounds__(IS_THREADS, 2)
attn_pool_infer_kernel(const float* __restrict__ QKV, const int64_t* __restrict__ node_off, int H, float scale,
                       float* __restrict__ pooled) {
    extern __shared__ __align__(16) float smem[];
    const int g = blockIdx.x;
    const int64_t n0 = node_off[g];
    const int n = (int)(node_off[g + 1] - n0);
    float* Ks = smem;                               // [NMAX][68]
    float* qs = Ks + IS_ATT_NMAX * IS_ATT_LD4;      // [8 warps][4][64]
    float* part = qs + 8 * 4 * 64;                  // [8 warps][NMAX] column-sum partials
    float* wcol = part + 8 * IS_ATT_NMAX;           // [NMAX]
    float* red = wcol + IS_ATT_NMAX;                // [4][64]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int dh = 64 / H;
    for (int idx = tid; idx < n * 16; idx += IS_THREADS) {
        const int r = idx >> 4, c4 = idx & 15;
        *reinterpret_cast<float4*>(Ks + r * IS_ATT_LD4 + 4 * c4) =
            __ldg(reinterpret_cast<const float4*>(QKV + (n0 + r) * 192 + 64) + c4);
    }
    __syncthreads();
    float* q = qs + warp * 256;
    const float inv_n = 1.0f / (float)max(n, 1);
    for (int h = 0; h < H; ++h) {
        float colacc[IS_ATT_NMAX / 32];
#pragma unroll
        for (int jj = 0; jj < IS_ATT_NMAX / 32; ++jj) colacc[jj] = 0.0f;
        for (int i0 = 4 * warp; i0 < n; i0 += 4 * (IS_THREADS / 32)) {
            __syncwarp();
            // stage this warp's 4 query rows (this head's dh channels), zero beyond n
            for (int t = lane; t < 4 * dh; t += 32) {
                const int r = t / dh, c = t - r * dh;
                q[r * 64 + c] = (i0 + r < n) ? __ldg(QKV + (n0 + i0 + r) * 192 + h * dh + c) : 0.0f;
            }
            __syncwarp();
            float acc[4][IS_ATT_NMAX / 32];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int jj = 0; jj < IS_ATT_NMAX / 32; ++jj) acc[r][jj] = 0.0f;
            for (int k = 0; k < dh; k += 4) {
                float4 q4[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) q4[r] = *reinterpret_cast<const float4*>(q + r * 64 + k);
#pragma unroll
                for (int jj = 0; jj < IS_ATT_NMAX / 32; ++jj) {
                    const int j = jj * 32 + lane;
                    if (jj * 32 < n) {      // warp-uniform skip of empty column blocks
                        const float4 kv = *reinterpret_cast<const float4*>(Ks + (j < n ? j : 0) * IS_ATT_LD4 + h * dh + k);
#pragma unroll
                        for (int r = 0; r < 4; ++r)
                            acc[r][jj] = fmaf(q4[r].x, kv.x, fmaf(q4[r].y, kv.y, fmaf(q4[r].z, kv.z, fmaf(q4[r].w, kv.w, acc[r][jj]))));
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                float mx = -INFINITY;
#pragma unroll
                for (int jj = 0; jj < IS_ATT_NMAX / 32; ++jj) {
                    const float a = (jj * 32 + lane < n) ? acc[r][jj] * scale : -INFINITY;
                    acc[r][jj] = a;
                    mx = fmaxf(mx, a);
                }
                mx = warp_max(mx);
                float z = 0.0f;
#pragma unroll
                for (int jj = 0; jj < IS_ATT_NMAX / 32; ++jj) {
                    const float e = (jj * 32 + lane < n) ? expf(acc[r][jj] - mx) : 0.0f;
                    acc[r][jj] = e;
                    z += e;
                }
                z = warp_sum(z);
                const float rz = (i0 + r < n) ? 1.0f / z : 0.0f;      // rows beyond n contribute nothing
#pragma unroll
                for (int jj = 0; jj < IS_ATT_NMAX / 32; ++jj) colacc[jj] = fmaf(acc[r][jj], rz, colacc[jj]);
            }
        }
#pragma unroll
        for (int jj = 0; jj < IS_ATT_NMAX / 32; ++jj) part[warp * IS_ATT_NMAX + jj * 32 + lane] = colacc[jj];
        __syncthreads();
        for (int j = tid; j < n; j += IS_THREADS) {
            float s = 0.0f;
#pragma unroll
            for (int w = 0; w < IS_THREADS / 32; ++w) s += part[w * IS_ATT_NMAX + j];
            wcol[j] = s * inv_n;
        }
        __syncthreads();
        // pooled[h*dh + c] = sum_j wcol[j] V[j][h*dh + c]: 4 row quarters x dh channels, fixed order
        {
            const int c = tid & 63, quarter = tid >> 6;
            float sacc = 0.0f;
            if (c < dh) {
                const int per = (n + 3) / 4, j0 = min(n, quarter * per), j1 = min(n, j0 + per);
                for (int j = j0; j < j1; ++j) sacc = fmaf(wcol[j], __ldg(QKV + (n0 + j) * 192 + 128 + h * dh + c), sacc);
            }
            red[quarter * 64 + c] = sacc;
        }
        __syncthreads();
        if (tid < dh) pooled[(int64_t)g * 64 + h * dh + tid] = red[tid] + red[64 + tid] + red[128 + tid] + red[192 + tid];
        __syncthreads();
    }
}

}  // namespace is

using namespace is;

extern "C" {

int is_attn_max_nodes(void) { return IS_ATT_NMAX; }

// max_nodes: the host's upper bound on nodes per graph (validated <= 256; larger graphs are rejected)
int is_attn_pool_fwd(const float* QKV, const int64_t* node_off, int n_graphs, int n_head, int max_nodes,
                     float* O, float* LSE, float* pooled, float* attn, const int64_t* attn_off, void* stream) {
    if (n_graphs <= 0 || !(n_head == 1 || n_head == 2 || n_head == 4 || n_head == 8)) return IS_ERR_ARG;
    if (max_nodes > IS_ATT_NMAX) return IS_ERR_UNSUPPORTED;
    const float scale = 1.0f / sqrtf((float)(64 / n_head));
    size_t smem = sizeof(float) * (2 * IS_ATT_NMAX * IS_ATT_LD + 8 * 64 + 8 * IS_ATT_NMAX + 4 * 64);
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    attn_fwd_kernel<<<n_graphs, IS_THREADS, smem, (cudaStream_t)stream>>>(QKV, node_off, n_head, scale, O, LSE, pooled, attn, attn_off);
    IS_LAUNCH_CHECK();
    return IS_OK;
}

// Inference-only variant: pooled [B,64] only (nothing saved for a backward pass, no weights output).
int is_attn_pool_infer(const float* QKV, const int64_t* node_off, int n_graphs, int n_head, int max_nodes,
                       float* pooled, void* stream) {
    if (n_graphs <= 0 || !(n_head == 1 || n_head == 2 || n_head == 4 || n_head == 8)) return IS_ERR_ARG;
    if (max_nodes > IS_ATT_NMAX) return IS_ERR_UNSUPPORTED;
    const float scale = 1.0f / sqrtf((float)(64 / n_head));
    size_t smem = sizeof(float) * (IS_ATT_NMAX * IS_ATT_LD4 + 8 * 4 * 64 + 8 * IS_ATT_NMAX + IS_ATT_NMAX + 4 * 64);
    cudaError_t e = cudaFuncSetAttribute(attn_pool_infer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    attn_pool_infer_kernel<<<n_graphs, IS_THREADS, smem, (cudaStream_t)stream>>>(QKV, node_off, n_head, scale, pooled);
    IS_LAUNCH_CHECK();
    return IS_OK;
}

int is_attn_pool_bwd(const float* QKV, const float* O, const float* LSE, const int64_t* node_off, int n_graphs,
                     int n_head, int max_nodes, const float* g_pooled, const float* gO_full, float* gQKV, void* stream) {
    if (n_graphs <= 0 || !(n_head == 1 || n_head == 2 || n_head == 4 || n_head == 8)) return IS_ERR_ARG;
    if (max_nodes > IS_ATT_NMAX) return IS_ERR_UNSUPPORTED;
    const float scale = 1.0f / sqrtf((float)(64 / n_head));
    size_t smem = sizeof(float) * (2 * IS_ATT_NMAX * IS_ATT_LD + 2 * 8 * 64 + 2 * 8 * IS_ATT_NMAX + IS_ATT_NMAX * 8);
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    attn_bwd_kernel<<<n_graphs, IS_THREADS, smem, (cudaStream_t)stream>>>(QKV, O, LSE, node_off, n_head, scale, g_pooled, gO_full, gQKV);
    IS_LAUNCH_CHECK();
    return IS_OK;
}

}  // extern "C"
