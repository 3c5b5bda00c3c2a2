// Pieces shared by the SIMT (egnn.cu) and tensor-core (egnn_tc.cu) EGNN edge kernels.
#pragma once
#include "common.cuh"

namespace is {

// =============================================================================================
// Tile selection shared by the edge kernels: the CTA owns the node range [nb, ne); a tile is the
// longest run of consecutive destination nodes (<= 32) whose in-edges total <= 128.
// Returns false when the range is exhausted.  tile = {n0, n1, p0, n_edges}.
// =============================================================================================
__device__ __forceinline__ void select_tile(int* s_tile, const int* __restrict__ indptr, int n0, int nend,
                                            int* __restrict__ status) {
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        const int pbase = __ldg(indptr + n0);
        const int cand = n0 + lane + 1;
        const bool ok = (cand <= nend) && (__ldg(indptr + (cand <= nend ? cand : nend)) - pbase <= IS_TM);
        const unsigned mask = __ballot_sync(0xffffffffu, ok);
        const int cnt = __popc(mask);          // ok is a prefix predicate (indptr is non-decreasing)
        if (lane == 0) {
            s_tile[0] = n0;
            s_tile[2] = pbase;
            if (cnt == 0) {                    // a single node with more than 128 in-edges: unsupported
                if (status) atomicExch(status, 1);
                s_tile[1] = n0 + 1;
                s_tile[3] = -1;
            } else {
                s_tile[1] = n0 + cnt;
                s_tile[3] = __ldg(indptr + n0 + cnt) - pbase;
            }
        }
    }
}

struct EdgeCommon {
    const int* indptr;
    const int* csr_src;
    const int* csr_dst;
    const int* csr_eid;
    const float* PQ;         // [N,128]
    const float* x;          // coords, row stride ldx
    int64_t ldx;
    const float* edge_attr;  // [E] in original edge order
    const float* W1;         // edge_mlp.0.weight [64, 2F+2]  (only columns 2F, 2F+1 are used here)
    int F;
    const float* W2; const float* b2;   // edge_mlp.2
    const float* W3; const float* b3;   // coord_mlp.0
    const float* w4;                    // coord_mlp.2.weight [1,64]
    int n_nodes;
    int* status;
};


// ---- tile walk: executed by one full warp; all lanes return the same values ------------------------
// Longest run of consecutive destination nodes (<= 32) starting at n0 whose in-edges total <= 128.
// tn0 >= nend means the CTA's node range is exhausted.  Nodes with more than 128 in-edges are
// unsupported: flagged in *status and skipped.  LIMIT = edge rows of the caller's tile (<= 128).
template <int LIMIT = IS_TM>
__device__ __forceinline__ void next_tile(const int* __restrict__ indptr, int n0, int nend, int* __restrict__ status,
                                          int lane, int& tn0, int& tn1, int& tp0, int& tne) {
    while (n0 < nend) {
        const int pbase = __ldg(indptr + n0);
        const int cand = n0 + lane + 1;
        const bool ok = (cand <= nend) && (__ldg(indptr + (cand <= nend ? cand : nend)) - pbase <= LIMIT);
        const int cnt = __popc(__ballot_sync(0xffffffffu, ok));
        if (cnt == 0) {
            if (lane == 0 && status) atomicExch(status, 1);
            ++n0;
            continue;
        }
        tn0 = n0; tn1 = n0 + cnt; tp0 = pbase; tne = __ldg(indptr + n0 + cnt) - pbase;
        return;
    }
    tn0 = nend; tn1 = nend; tp0 = 0; tne = 0;
}


}  // namespace is
