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

#define IS_ATT_LD4 68
#define IS_ATT_JB (IS_ATT_NMAX / 32)

// ---- register-blocked building blocks (one warp, 4 rows at a time) ------------------------------------
// acc[r][jj] = sum_{k < dh} rows[r*64 + c0 + k] * Mat[(jj*32 + lane) * 68 + c0 + k]
// rows: this warp's [4][64] staging area; Mat: [n][68] in shared memory.  128 FMAs per 12 LDS.128.
__device__ __forceinline__ void blocked_dots(float (&acc)[4][IS_ATT_JB], const float* __restrict__ rows,
                                             const float* __restrict__ Mat, int c0, int dh, int n, int lane) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int jj = 0; jj < IS_ATT_JB; ++jj) acc[r][jj] = 0.0f;
    for (int k = 0; k < dh; k += 4) {
        float4 q4[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) q4[r] = *reinterpret_cast<const float4*>(rows + r * 64 + c0 + k);
#pragma unroll
        for (int jj = 0; jj < IS_ATT_JB; ++jj) {
            if (jj * 32 < n) {          // warp-uniform
                const int j = jj * 32 + lane;
                const float4 kv = *reinterpret_cast<const float4*>(Mat + (j < n ? j : 0) * IS_ATT_LD4 + c0 + k);
#pragma unroll
                for (int r = 0; r < 4; ++r)
                    acc[r][jj] = fmaf(q4[r].x, kv.x, fmaf(q4[r].y, kv.y, fmaf(q4[r].z, kv.z, fmaf(q4[r].w, kv.w, acc[r][jj]))));
            }
        }
    }
}

// out[r] (float2: channels c0 + 2*lane, +1; lanes with 2*lane >= dh idle) = sum_j W[r*NMAX + j] * Mat[j*68 + c0 + 2*lane ..]
__device__ __forceinline__ void blocked_wsum(float2 (&out)[4], const float* __restrict__ W, const float* __restrict__ Mat,
                                             int c0, int dh, int n, int lane) {
#pragma unroll
    for (int r = 0; r < 4; ++r) out[r] = make_float2(0.0f, 0.0f);
    if (2 * lane < dh) {
        const float* mc = Mat + c0 + 2 * lane;
        int j = 0;
        for (; j + 4 <= n; j += 4) {
            float4 w4[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) w4[r] = *reinterpret_cast<const float4*>(W + r * IS_ATT_NMAX + j);
            const float2 v0 = *reinterpret_cast<const float2*>(mc + (j + 0) * IS_ATT_LD4);
            const float2 v1 = *reinterpret_cast<const float2*>(mc + (j + 1) * IS_ATT_LD4);
            const float2 v2 = *reinterpret_cast<const float2*>(mc + (j + 2) * IS_ATT_LD4);
            const float2 v3 = *reinterpret_cast<const float2*>(mc + (j + 3) * IS_ATT_LD4);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                out[r].x = fmaf(w4[r].x, v0.x, fmaf(w4[r].y, v1.x, fmaf(w4[r].z, v2.x, fmaf(w4[r].w, v3.x, out[r].x))));
                out[r].y = fmaf(w4[r].x, v0.y, fmaf(w4[r].y, v1.y, fmaf(w4[r].z, v2.y, fmaf(w4[r].w, v3.y, out[r].y))));
            }
        }
        for (; j < n; ++j) {
            const float2 v = *reinterpret_cast<const float2*>(mc + j * IS_ATT_LD4);
#pragma unroll
            for (int r = 0; r < 4; ++r) { out[r].x = fmaf(W[r * IS_ATT_NMAX + j], v.x, out[r].x); out[r].y = fmaf(W[r * IS_ATT_NMAX + j], v.y, out[r].y); }
        }
    }
}

// stage rows [i0, i0+4) x 64 channels of a [N,ld] global matrix (+ optional per-graph row `add`, scaled)
// into this warp's [4][64] area; rows >= n are zero
__device__ __forceinline__ void stage4(float* __restrict__ dst, const float* __restrict__ src, int64_t ld, int64_t n0,
                                       int i0, int n, int lane, const float* __restrict__ add = nullptr, float add_scale = 0.0f) {
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        const int idx = t * 32 + lane, r = idx >> 6, c = idx & 63;
        float v = 0.0f;
        if (i0 + r < n) {
            v = src ? __ldg(src + (n0 + i0 + r) * ld + c) : 0.0f;
            if (add) v = fmaf(__ldg(add + c), add_scale, v);
        }
        dst[idx] = v;
    }
}

__global__ void __launch_bounds__(IS_THREADS)
attn_fwd_kernel(const float* __restrict__ QKV, const int64_t* __restrict__ node_off, int H, float scale,
                float* __restrict__ O, float* __restrict__ LSE, float* __restrict__ pooled,
                float* __restrict__ attn /* or null */, const int64_t* __restrict__ attn_off /* [B] or null */) {
    extern __shared__ __align__(16) float smem[];
    const int g = blockIdx.x;
    const int64_t n0 = node_off[g];
    // clamped: an understated max_nodes (caught by GraphBatch validation) must not become a shared-memory overrun
    const int n = min((int)(node_off[g + 1] - n0), IS_ATT_NMAX);
    float* Ks = smem;                               // [NMAX][68]
    float* Vs = Ks + IS_ATT_NMAX * IS_ATT_LD4;      // [NMAX][68]
    float* qs = Vs + IS_ATT_NMAX * IS_ATT_LD4;      // [8 warps][4][64]
    float* prow = qs + 8 * 256;                     // [8 warps][4][NMAX]
    float* red = prow + 8 * 4 * IS_ATT_NMAX;        // [4][64]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int dh = 64 / H;
    for (int idx = tid; idx < n * 16; idx += IS_THREADS) {
        const int r = idx >> 4, c4 = idx & 15;
        *reinterpret_cast<float4*>(Ks + r * IS_ATT_LD4 + 4 * c4) = __ldg(reinterpret_cast<const float4*>(QKV + (n0 + r) * 192 + 64) + c4);
        *reinterpret_cast<float4*>(Vs + r * IS_ATT_LD4 + 4 * c4) = __ldg(reinterpret_cast<const float4*>(QKV + (n0 + r) * 192 + 128) + c4);
    }
    __syncthreads();
    float* q = qs + warp * 256;
    float* pr = prow + warp * 4 * IS_ATT_NMAX;
    for (int i0 = 4 * warp; i0 < n; i0 += 4 * (IS_THREADS / 32)) {
        __syncwarp();
        stage4(q, QKV, 192, n0, i0, n, lane);
        __syncwarp();
        for (int h = 0; h < H; ++h) {
            float acc[4][IS_ATT_JB];
            blocked_dots(acc, q, Ks, h * dh, dh, n, lane);
            __syncwarp();                              // previous head's prow reads are complete
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                float mx = -INFINITY;
#pragma unroll
                for (int jj = 0; jj < IS_ATT_JB; ++jj) {
                    const float a = (jj * 32 + lane < n) ? acc[r][jj] * scale : -INFINITY;
                    acc[r][jj] = a;
                    mx = fmaxf(mx, a);
                }
                mx = warp_max(mx);
                float z = 0.0f;
#pragma unroll
                for (int jj = 0; jj < IS_ATT_JB; ++jj) {
                    const float e = (jj * 32 + lane < n) ? expf(acc[r][jj] - mx) : 0.0f;
                    acc[r][jj] = e;
                    z += e;
                }
                z = warp_sum(z);
                const float rz = 1.0f / z;
                const int i = i0 + r;
#pragma unroll
                for (int jj = 0; jj < IS_ATT_JB; ++jj) {
                    const int j = jj * 32 + lane;
                    if (jj * 32 < n) {
                        const float pj = acc[r][jj] * rz;
                        pr[r * IS_ATT_NMAX + j] = (j < n) ? pj : 0.0f;
                        if (attn && i < n && j < n) attn[attn_off[g] + ((int64_t)h * n + i) * n + j] = pj;
                    }
                }
                if (lane == 0 && i < n) LSE[(n0 + i) * H + h] = mx + logf(z);
            }
            __syncwarp();
            float2 o[4];
            blocked_wsum(o, pr, Vs, h * dh, dh, n, lane);
            if (2 * lane < dh) {
#pragma unroll
                for (int r = 0; r < 4; ++r)
                    if (i0 + r < n) *reinterpret_cast<float2*>(O + (n0 + i0 + r) * 64 + h * dh + 2 * lane) = o[r];
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

// Backward.  gO[i] = gO_full[i] (optional) + g_pooled[g] / n.  Two register-blocked passes with
// recomputed scores: row blocks (gQ) then column blocks (gK, gV), fixed summation order.
// Pooled-only case (gO_full == NULL: every model's training path): all rows of gO equal g0 = g_pooled / n, so
// dP_ij = g0 . V_j =: c_j does not depend on i and gV_j = (sum_i P_ij) g0 -- c is computed once per graph and three of
// the seven n x n x 64 products (both dP passes and the P^T gO product) disappear.  With O == NULL (pooled-only, forward
// run by the pooled-rows-only tensor-core kernel) the row statistics are recomputed here: pass A derives lse_i from the
// score row it holds anyway and D_i = sum_j P_ij c_j, and parks lse_i in the LSE array (then an OUTPUT scratch) for pass B.
__global__ void __launch_bounds__(IS_THREADS)
attn_bwd_kernel(const float* __restrict__ QKV, const float* __restrict__ O /* or null: recompute */, float* __restrict__ LSE,
                const int64_t* __restrict__ node_off, int H, float scale,
                const float* __restrict__ g_pooled /* [B,64] or null */, const float* __restrict__ gO_full /* or null */,
                float* __restrict__ gQKV) {
    extern __shared__ __align__(16) float smem[];
    const int g = blockIdx.x;
    const int64_t n0 = node_off[g];
    // clamped: an understated max_nodes (caught by GraphBatch validation) must not become a shared-memory overrun
    const int n = min((int)(node_off[g + 1] - n0), IS_ATT_NMAX);
    float* S0 = smem;                               // pass A: K   ; pass B: Q      [NMAX][68]
    float* S1 = S0 + IS_ATT_NMAX * IS_ATT_LD4;      // pass A: V   ; pass B: gO     [NMAX][68]
    float* ra = S1 + IS_ATT_NMAX * IS_ATT_LD4;      // [8][4][64] pass A: q rows  ; pass B: k rows
    float* rb = ra + 8 * 256;                       // [8][4][64] pass A: gO rows ; pass B: v rows
    float* wa = rb + 8 * 256;                       // [8][4][NMAX] gs
    float* wb = wa + 8 * 4 * IS_ATT_NMAX;           // [8][4][NMAX] p (pass B)
    float* Dn = wb + 8 * 4 * IS_ATT_NMAX;           // [NMAX][8]   D[i][h] = gO_i . O_i per head
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int dh = 64 / H;
    const float inv_n = 1.0f / (float)max(n, 1);
    const float* gp_row = g_pooled ? g_pooled + (int64_t)g * 64 : nullptr;
    float* A = ra + warp * 256;
    float* Bv = rb + warp * 256;
    float* WA = wa + warp * 4 * IS_ATT_NMAX;
    float* WB = wb + warp * 4 * IS_ATT_NMAX;

    // ---------------- pass A: row blocks ----------------
    for (int idx = tid; idx < n * 16; idx += IS_THREADS) {
        const int r = idx >> 4, c4 = idx & 15;
        *reinterpret_cast<float4*>(S0 + r * IS_ATT_LD4 + 4 * c4) = __ldg(reinterpret_cast<const float4*>(QKV + (n0 + r) * 192 + 64) + c4);
        *reinterpret_cast<float4*>(S1 + r * IS_ATT_LD4 + 4 * c4) = __ldg(reinterpret_cast<const float4*>(QKV + (n0 + r) * 192 + 128) + c4);
    }
    __syncthreads();
    const bool pooled_only = gO_full == nullptr && gp_row != nullptr;
    float* cvec = rb;                               // [8 heads][NMAX] c_j per head (pooled-only: rb is not used for gO rows)
    if (pooled_only) {
        for (int idx = tid; idx < n * H; idx += IS_THREADS) {
            const int j = idx / H, h = idx - j * H;
            float c = 0.0f;
            for (int k = 0; k < dh; ++k) c = fmaf(__ldg(gp_row + h * dh + k) * inv_n, S1[j * IS_ATT_LD4 + h * dh + k], c);
            cvec[h * IS_ATT_NMAX + j] = c;
        }
        __syncthreads();
    }
    for (int i0 = 4 * warp; i0 < n; i0 += 4 * (IS_THREADS / 32)) {
        __syncwarp();
        stage4(A, QKV, 192, n0, i0, n, lane);
        if (!pooled_only) stage4(Bv, gO_full, 64, n0, i0, n, lane, gp_row, inv_n);
        __syncwarp();
        for (int h = 0; h < H; ++h) {
            // D[r] = sum over this head's channels of gO[r][c] * O[r][c]
            float D[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                float dpart = 0.0f;
                if (i0 + r < n && O != nullptr)
                    for (int k = lane; k < dh; k += 32)
                        dpart += (pooled_only ? __ldg(gp_row + h * dh + k) * inv_n : Bv[r * 64 + h * dh + k]) * __ldg(O + (n0 + i0 + r) * 64 + h * dh + k);
                D[r] = warp_sum(dpart);
                if (lane == 0 && i0 + r < n && O != nullptr) Dn[(i0 + r) * 8 + h] = D[r];
            }
            float s[4][IS_ATT_JB], gp[4][IS_ATT_JB];
            blocked_dots(s, A, S0, h * dh, dh, n, lane);
            if (pooled_only) {
#pragma unroll
                for (int jj = 0; jj < IS_ATT_JB; ++jj) {
                    const float c = (jj * 32 < n && jj * 32 + lane < n) ? cvec[h * IS_ATT_NMAX + jj * 32 + lane] : 0.0f;
#pragma unroll
                    for (int r = 0; r < 4; ++r) gp[r][jj] = c;
                }
            } else {
                blocked_dots(gp, Bv, S1, h * dh, dh, n, lane);
            }
            __syncwarp();
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                float lse = 0.0f;
                if (O != nullptr) {
                    lse = (i0 + r < n) ? LSE[(n0 + i0 + r) * H + h] : 0.0f;
                } else {
                    // recompute the row's log-sum-exp and D = sum_j p_j c_j from the scores held in registers
                    float mx = -INFINITY;
#pragma unroll
                    for (int jj = 0; jj < IS_ATT_JB; ++jj)
                        if (jj * 32 + lane < n) mx = fmaxf(mx, s[r][jj] * scale);
                    mx = warp_max(mx);
                    float z = 0.0f, dsum = 0.0f;
#pragma unroll
                    for (int jj = 0; jj < IS_ATT_JB; ++jj)
                        if (jj * 32 + lane < n) {
                            const float e = expf(s[r][jj] * scale - mx);
                            z += e;
                            dsum = fmaf(e, gp[r][jj], dsum);
                        }
                    z = warp_sum(z);
                    dsum = warp_sum(dsum);
                    lse = mx + logf(z);
                    D[r] = dsum / z;
                    if (lane == 0 && i0 + r < n) { LSE[(n0 + i0 + r) * H + h] = lse; Dn[(i0 + r) * 8 + h] = D[r]; }
                }
#pragma unroll
                for (int jj = 0; jj < IS_ATT_JB; ++jj) {
                    const int j = jj * 32 + lane;
                    if (jj * 32 < n) WA[r * IS_ATT_NMAX + j] = (j < n && i0 + r < n) ? expf(s[r][jj] * scale - lse) * (gp[r][jj] - D[r]) : 0.0f;
                }
            }
            __syncwarp();
            float2 o[4];
            blocked_wsum(o, WA, S0, h * dh, dh, n, lane);
            if (2 * lane < dh) {
#pragma unroll
                for (int r = 0; r < 4; ++r)
                    if (i0 + r < n)
                        *reinterpret_cast<float2*>(gQKV + (n0 + i0 + r) * 192 + h * dh + 2 * lane) = make_float2(o[r].x * scale, o[r].y * scale);
            }
        }
    }
    __syncthreads();
    // ---------------- pass B: column blocks ----------------
    for (int idx = tid; idx < n * 64; idx += IS_THREADS) {
        const int r = idx >> 6, c = idx & 63;
        S0[r * IS_ATT_LD4 + c] = __ldg(QKV + (n0 + r) * 192 + c);
        if (!pooled_only) {
            float v = gp_row ? __ldg(gp_row + c) * inv_n : 0.0f;
            if (gO_full) v += __ldg(gO_full + (n0 + r) * 64 + c);
            S1[r * IS_ATT_LD4 + c] = v;
        }
    }
    __syncthreads();
    for (int j0 = 4 * warp; j0 < n; j0 += 4 * (IS_THREADS / 32)) {
        __syncwarp();
        stage4(A, QKV + 64, 192, n0, j0, n, lane);       // K rows of the 4 columns
        if (!pooled_only) stage4(Bv, QKV + 128, 192, n0, j0, n, lane);     // V rows
        __syncwarp();
        for (int h = 0; h < H; ++h) {
            float s[4][IS_ATT_JB], gp[4][IS_ATT_JB];
            blocked_dots(s, A, S0, h * dh, dh, n, lane);       // s[c][i] = k_c . q_i
            if (pooled_only) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float cj = j0 + c < n ? cvec[h * IS_ATT_NMAX + j0 + c] : 0.0f;
#pragma unroll
                    for (int jj = 0; jj < IS_ATT_JB; ++jj) gp[c][jj] = cj;
                }
            } else {
                blocked_dots(gp, Bv, S1, h * dh, dh, n, lane);     // gp[c][i] = v_c . gO_i
            }
            __syncwarp();
            float psum[4] = {0.0f, 0.0f, 0.0f, 0.0f};            // pooled-only: sum_i p_ic (this lane's rows)
#pragma unroll
            for (int jj = 0; jj < IS_ATT_JB; ++jj) {
                const int i = jj * 32 + lane;
                if (jj * 32 < n) {
                    const bool vi = i < n;
                    const float lse = vi ? LSE[(n0 + i) * H + h] : 0.0f;
                    const float Di = vi ? Dn[i * 8 + h] : 0.0f;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const float pij = (vi && j0 + c < n) ? expf(s[c][jj] * scale - lse) : 0.0f;
                        if (pooled_only) psum[c] += pij; else WB[c * IS_ATT_NMAX + i] = pij;
                        WA[c * IS_ATT_NMAX + i] = pij * (gp[c][jj] - Di);
                    }
                }
            }
            __syncwarp();
            float2 gk[4], gv[4];
            blocked_wsum(gk, WA, S0, h * dh, dh, n, lane);
            if (pooled_only) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float w = warp_sum(psum[c]);
                    gv[c] = 2 * lane < dh ? make_float2(w * __ldg(gp_row + h * dh + 2 * lane) * inv_n, w * __ldg(gp_row + h * dh + 2 * lane + 1) * inv_n)
                                          : make_float2(0.0f, 0.0f);
                }
            } else {
                blocked_wsum(gv, WB, S1, h * dh, dh, n, lane);
            }
            if (2 * lane < dh) {
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (j0 + c < n) {
                        *reinterpret_cast<float2*>(gQKV + (n0 + j0 + c) * 192 + 64 + h * dh + 2 * lane) = make_float2(gk[c].x * scale, gk[c].y * scale);
                        *reinterpret_cast<float2*>(gQKV + (n0 + j0 + c) * 192 + 128 + h * dh + 2 * lane) = gv[c];
                    }
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
    // clamped: an understated max_nodes (caught by GraphBatch validation) must not become a shared-memory overrun
    const int n = min((int)(node_off[g + 1] - n0), IS_ATT_NMAX);
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
    size_t smem = sizeof(float) * (2 * IS_ATT_NMAX * IS_ATT_LD4 + 8 * 256 + 8 * 4 * IS_ATT_NMAX + 4 * 64);
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

int is_attn_pool_bwd(const float* QKV, const float* O, float* LSE, const int64_t* node_off, int n_graphs,
                     int n_head, int max_nodes, const float* g_pooled, const float* gO_full, float* gQKV, void* stream) {
    if (n_graphs <= 0 || !(n_head == 1 || n_head == 2 || n_head == 4 || n_head == 8)) return IS_ERR_ARG;
    if (O == nullptr && (gO_full != nullptr || g_pooled == nullptr)) return IS_ERR_ARG;   // statistics are recomputed only in the pooled-only case
    if (max_nodes > IS_ATT_NMAX) return IS_ERR_UNSUPPORTED;
    const float scale = 1.0f / sqrtf((float)(64 / n_head));
    size_t smem = sizeof(float) * (2 * IS_ATT_NMAX * IS_ATT_LD4 + 2 * 8 * 256 + 2 * 8 * 4 * IS_ATT_NMAX + IS_ATT_NMAX * 8);
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    attn_bwd_kernel<<<n_graphs, IS_THREADS, smem, (cudaStream_t)stream>>>(QKV, O, LSE, node_off, n_head, scale, g_pooled, gO_full, gQKV);
    IS_LAUNCH_CHECK();
    return IS_OK;
}

}  // extern "C"
