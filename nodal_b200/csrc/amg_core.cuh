// Per-row decisions of the pairwise-aggregation AMG setup, written once for the CUDA kernels
// (amg.cu) and for the host harness the CPU tests compile with g++ (tests/host_check/amg_host.cpp),
// so the matching rule can be checked against its numpy statement without a GPU.
//
// Every decision depends only on the row's own entries and on state written by an earlier
// kernel, never on what other threads of the same launch do: the aggregates are the same for
// any grid size and any scheduling.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define AMG_HD __host__ __device__ __forceinline__
#else
#define AMG_HD static inline
#endif

// Symmetric 32-bit hash of an edge: breaks the ties of equal couplings (every edge of a uniform
// grid) without the directional bias "smallest index first" would give a handshake matching.
AMG_HD uint32_t amg_edge_hash(int32_t i, int32_t j) {
    const uint32_t lo = (uint32_t)(i < j ? i : j), hi = (uint32_t)(i < j ? j : i);
    uint32_t h = (lo * 2654435761u) ^ (hi * 40503u + 0x9e3779b9u);
    h ^= h >> 15;
    h *= 2246822519u;
    h ^= h >> 13;
    return h;
}

// The neighbour row i prefers: the largest coupling -a_ij > 0, then the largest edge hash, then
// the smallest index.  With `match` only neighbours that are still unmatched qualify.
// Row-partitioned setup (dist_amg.cu): rows and columns are LOCAL indices, columns >= nown are
// halo entries owned by another rank and never qualify (aggregates do not cross the partition),
// and `base` (the global index of local row 0) keeps the tie-breaking hash a function of the
// global edge, so the aggregates do not depend on how many ranks there are beyond that rule.
AMG_HD int32_t amg_pick(int32_t i, const int32_t* indptr, const int32_t* indices,
                        const double* data, const int32_t* match,
                        int32_t nown = 0x7fffffff, int32_t base = 0) {
    int32_t best = -1;
    double bw = 0.0;
    uint32_t bh = 0u;
    const int32_t e = indptr[i + 1];
    for (int32_t p = indptr[i]; p < e; ++p) {
        const int32_t j = indices[p];
        const double a = data[p];
        if (j == i || j >= nown || !(a < 0.0)) continue;
        if (match && match[j] >= 0) continue;
        const double w = -a;
        const uint32_t h = amg_edge_hash(i + base, j + base);
        if (best < 0 || w > bw || (w == bw && (h > bh || (h == bh && j < best)))) {
            best = j;
            bw = w;
            bh = h;
        }
    }
    return best;
}

// Representative of row i's aggregate: the smaller index of its pair; a row the matching left
// alone joins the pair of its preferred neighbour if that neighbour is matched (it never chains
// onto another lone row), otherwise it stays a singleton.
AMG_HD int32_t amg_root(int32_t i, const int32_t* indptr, const int32_t* indices,
                        const double* data, const int32_t* match,
                        int32_t nown = 0x7fffffff, int32_t base = 0) {
    const int32_t m = match[i];
    if (m >= 0) return m < i ? m : i;
    const int32_t t = amg_pick(i, indptr, indices, data, nullptr, nown, base);
    if (t >= 0) {
        const int32_t mt = match[t];
        if (mt >= 0) return mt < t ? mt : t;
    }
    return i;
}
