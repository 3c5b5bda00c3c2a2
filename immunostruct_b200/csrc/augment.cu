// On-device training augmentations of the reference's data pipeline.
//
//   * RandomRotation (data/utils.py:148-155, applied per sample at data/util_dataloader.py:27,38,41):
//       M = randn(3, 3); Q, _ = numpy.linalg.qr(M); coords <- coords @ Q
//     Here M comes from the caller (one 3x3 standard-normal draw per graph, torch's device generator), the
//     Householder QR follows LAPACK's dgeqrf / dorgqr sign convention (beta = -sign(alpha) * norm), so for the same M
//     the rotation equals numpy's to rounding, and every node of the graph is rotated in the same kernel.
//   * mask_single_structure (data/immmunopred_dataloader.py:104-115, pair form :248-266): pick one non-padded residue
//     (any non-zero one-hot column) uniformly, overwrite its 20 one-hot columns with ones, return its residue id.
//     The reference draws with random.choice in a rejection loop; uniform over the valid residues is the same
//     distribution: pick = floor(u * n_valid) with u ~ U[0,1) from the caller.  `want_aa` >= 0 restricts the choice
//     to residues of that type (the wild-type partner of a pair).
//   * mask_structure (:92-102, pair form :234-247): `count` distinct nodes per graph (random.sample) -> zero their
//     one-hot columns unless the row already sums to more than 1 (the SSL-masked residue).  Here: the `count`
//     smallest of the caller's per-node uniform keys.
//   * mask_sequence (:78-90, pair form :216-231): `count` distinct positions among the first `limit` rows of a
//     sequence -> the padding token's one-hot.  Same key-based sampling.
// The CPU contracts (oracle/kernel_contracts.py) consume the same M / u / keys, so parity is bit-exact for the
// integer choices and 1e-6 for the rotation.
#include "common.cuh"

namespace is {

// Householder QR of a 3x3 matrix, LAPACK convention; q (row-major) = the orthogonal factor
__device__ __forceinline__ void qr3_q(const float* __restrict__ m, float (&q)[9]) {
    float a[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) a[i][j] = m[3 * i + j];
    float v[2][3], tau[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        // dlarfg on column k, rows k..2
        float xnorm2 = 0.0f;
        for (int i = k + 1; i < 3; ++i) xnorm2 += a[i][k] * a[i][k];
        const float alpha = a[k][k];
        v[k][0] = v[k][1] = v[k][2] = 0.0f;
        v[k][k] = 1.0f;
        if (xnorm2 == 0.0f) {
            tau[k] = 0.0f;
        } else {
            const float beta = -copysignf(sqrtf(alpha * alpha + xnorm2), alpha);
            tau[k] = (beta - alpha) / beta;
            const float sc = 1.0f / (alpha - beta);
            for (int i = k + 1; i < 3; ++i) v[k][i] = a[i][k] * sc;
            a[k][k] = beta;
            // apply H = I - tau v v^T to the trailing columns
            for (int j = k + 1; j < 3; ++j) {
                float w = 0.0f;
                for (int i = k; i < 3; ++i) w += v[k][i] * a[i][j];
                w *= tau[k];
                for (int i = k; i < 3; ++i) a[i][j] -= w * v[k][i];
            }
        }
    }
    // Q = H0 H1 (H2 = I for a 3x3: dlarfg on a single element gives tau = 0)
#pragma unroll
    for (int i = 0; i < 9; ++i) q[i] = (i % 4 == 0) ? 1.0f : 0.0f;
#pragma unroll
    for (int k = 1; k >= 0; --k) {
        for (int j = 0; j < 3; ++j) {
            float w = 0.0f;
            for (int i = k; i < 3; ++i) w += v[k][i] * q[3 * i + j];
            w *= tau[k];
            for (int i = k; i < 3; ++i) q[3 * i + j] -= w * v[k][i];
        }
    }
}

// x[:, c0:c0+3] <- x[:, c0:c0+3] @ Q_g for every node of graph g; Qout [B, 9] optional
__global__ void __launch_bounds__(128)
rotate_coords_kernel(float* __restrict__ x, int64_t ldx, int c0, const int64_t* __restrict__ node_off,
                     const float* __restrict__ M, float* __restrict__ Qout) {
    const int g = blockIdx.x;
    float q[9];
    qr3_q(M + 9 * (int64_t)g, q);
    if (Qout != nullptr && threadIdx.x < 9) Qout[9 * (int64_t)g + threadIdx.x] = q[threadIdx.x];
    const int64_t r0 = node_off[g], r1 = node_off[g + 1];
    for (int64_t r = r0 + threadIdx.x; r < r1; r += blockDim.x) {
        float* p = x + r * ldx + c0;
        const float a = p[0], b = p[1], c = p[2];
        p[0] = a * q[0] + b * q[3] + c * q[6];
        p[1] = a * q[1] + b * q[4] + c * q[7];
        p[2] = a * q[2] + b * q[5] + c * q[8];
    }
}

// one CTA (one warp) per graph.  aa_out[g] = residue id of the masked node, 0 if the graph has no valid residue
// (the reference prints "unmaskable graph" and returns tensor([0])); node_out[g] = its global row or -1.
__global__ void __launch_bounds__(32)
mask_single_kernel(float* __restrict__ x, int64_t ldx, int n_feat, const int64_t* __restrict__ node_off,
                   const float* __restrict__ u, const int64_t* __restrict__ want_aa, int64_t* __restrict__ aa_out,
                   int64_t* __restrict__ node_out) {
    const int g = blockIdx.x, lane = threadIdx.x;
    const int64_t r0 = node_off[g], r1 = node_off[g + 1];
    const int want = want_aa ? (int)want_aa[g] : -1;
    // first non-zero column of a row (-1: padded / already zeroed); rows that are all ones report column 0
    auto first_nz = [&](int64_t r) {
        for (int c = 0; c < n_feat; ++c)
            if (x[r * ldx + c] != 0.0f) return c;
        return -1;
    };
    int cnt = 0;
    for (int64_t r = r0 + lane; r < r1; r += 32) {
        const int a = first_nz(r);
        cnt += (a >= 0 && (want < 0 || a == want)) ? 1 : 0;
    }
    // exclusive scan over lanes is not needed: candidates are numbered in row order, so walk in 32-row groups
    int total = cnt;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
    int64_t chosen = -1;
    int chosen_aa = 0;
    if (total > 0) {
        int pick = (int)(u[g] * (float)total);
        if (pick >= total) pick = total - 1;
        int seen = 0;
        for (int64_t rb = r0; rb < r1 && chosen < 0; rb += 32) {
            const int64_t r = rb + lane;
            const int a = r < r1 ? first_nz(r) : -1;
            const bool ok = a >= 0 && (want < 0 || a == want);
            const unsigned m = __ballot_sync(0xffffffffu, ok);
            const int here = __popc(m);
            if (pick < seen + here) {
                // the (pick - seen)-th set bit of m
                unsigned mm = m;
                for (int k = 0; k < pick - seen; ++k) mm &= mm - 1;
                const int src_lane = __ffs(mm) - 1;
                chosen = rb + src_lane;
                chosen_aa = __shfl_sync(0xffffffffu, a, src_lane);
            }
            seen += here;
        }
    }
    if (chosen >= 0)
        for (int c = lane; c < n_feat; c += 32) x[chosen * ldx + c] = 1.0f;
    if (lane == 0) {
        aa_out[g] = chosen_aa;
        if (node_out) node_out[g] = chosen;
    }
}

// rows of segment g (seg_off, at most `limit[g]` leading rows when limit != NULL) with the `count` smallest keys
// are overwritten in columns [0, n_cols): fill_col < 0 -> zeros (mask_structure; rows whose sum exceeds 1 are left
// alone), else the one-hot of `fill_col` (mask_sequence).  One CTA per segment; keys are staged in shared memory.
__global__ void __launch_bounds__(256)
mask_rows_kernel(float* __restrict__ data, int64_t ld, int n_cols, const int64_t* __restrict__ seg_off,
                 const int64_t* __restrict__ limit, const float* __restrict__ keys, int count, int fill_col,
                 int max_rows) {
    extern __shared__ float skey[];               // [max_rows]
    __shared__ float red_k[256];
    __shared__ int red_i[256];
    const int g = blockIdx.x, tid = threadIdx.x;
    const int64_t r0 = seg_off[g];
    int n = (int)(seg_off[g + 1] - r0);
    if (limit) n = min(n, (int)limit[g]);
    n = min(n, max_rows);
    for (int i = tid; i < n; i += blockDim.x) skey[i] = keys[r0 + i];
    __syncthreads();
    const int k = min(count, n);
    for (int it = 0; it < k; ++it) {
        float best = INFINITY;
        int bi = 0x7fffffff;
        for (int i = tid; i < n; i += blockDim.x) {
            const float v = skey[i];
            if (v < best || (v == best && i < bi)) { best = v; bi = i; }
        }
        red_k[tid] = best; red_i[tid] = bi;
        __syncthreads();
        for (int s = 128; s > 0; s >>= 1) {
            if (tid < s) {
                const float v = red_k[tid + s];
                const int i2 = red_i[tid + s];
                if (v < red_k[tid] || (v == red_k[tid] && i2 < red_i[tid])) { red_k[tid] = v; red_i[tid] = i2; }
            }
            __syncthreads();
        }
        const int sel = red_i[0];
        __syncthreads();
        if (sel == 0x7fffffff) break;
        if (tid == 0) skey[sel] = INFINITY;
        float* row = data + (r0 + sel) * ld;
        if (fill_col < 0) {
            // mask_structure: skip rows already set to all ones by mask_single_structure (sum > 1)
            float s = 0.0f;
            for (int c = 0; c < n_cols; ++c) s += row[c];
            __syncthreads();
            if (!(s > 1.0f))
                for (int c = tid; c < n_cols; c += blockDim.x) row[c] = 0.0f;
        } else {
            for (int c = tid; c < n_cols; c += blockDim.x) row[c] = (c == fill_col) ? 1.0f : 0.0f;
        }
        __syncthreads();
    }
}

}  // namespace is

using namespace is;

extern "C" {

// RandomRotation per graph: x[:, c0:c0+3] @= Q(M_g); M [n_graphs, 9] standard-normal draws; Qout [n_graphs, 9] or NULL
int is_rotate_coords(float* x, int64_t ldx, int c0, const int64_t* node_off, int n_graphs, const float* M, float* Qout,
                     void* stream) {
    if (n_graphs < 0 || c0 < 0 || ldx < c0 + 3) return IS_ERR_ARG;
    if (n_graphs == 0) return IS_OK;
    rotate_coords_kernel<<<n_graphs, 128, 0, (cudaStream_t)stream>>>(x, ldx, c0, node_off, M, Qout);
    IS_LAUNCH_CHECK();
    return IS_OK;
}

// mask_single_structure: u [n_graphs] in [0,1); want_aa [n_graphs] or NULL; aa_out [n_graphs]; node_out [n_graphs] or NULL
int is_mask_single_residue(float* x, int64_t ldx, int n_feat, const int64_t* node_off, int n_graphs, const float* u,
                           const int64_t* want_aa, int64_t* aa_out, int64_t* node_out, void* stream) {
    if (n_graphs < 0 || n_feat <= 0 || ldx < n_feat) return IS_ERR_ARG;
    if (n_graphs == 0) return IS_OK;
    mask_single_kernel<<<n_graphs, 32, 0, (cudaStream_t)stream>>>(x, ldx, n_feat, node_off, u, want_aa, aa_out, node_out);
    IS_LAUNCH_CHECK();
    return IS_OK;
}

// mask_structure (fill_col < 0) / mask_sequence (fill_col = padding token): see mask_rows_kernel.
// keys: one uniform per row of `data`; max_rows >= the longest segment (<= 12 000 rows of shared-memory keys).
int is_mask_rows(float* data, int64_t ld, int n_cols, const int64_t* seg_off, const int64_t* limit, int n_segments,
                 const float* keys, int count, int fill_col, int max_rows, void* stream) {
    if (n_segments < 0 || n_cols <= 0 || ld < n_cols || count < 0 || fill_col >= n_cols || max_rows <= 0) return IS_ERR_ARG;
    if (max_rows > 12000) return IS_ERR_UNSUPPORTED;
    if (n_segments == 0 || count == 0) return IS_OK;
    mask_rows_kernel<<<n_segments, 256, (size_t)max_rows * sizeof(float), (cudaStream_t)stream>>>(
        data, ld, n_cols, seg_off, limit, keys, count, fill_col, max_rows);
    IS_LAUNCH_CHECK();
    return IS_OK;
}

}  // extern "C"
