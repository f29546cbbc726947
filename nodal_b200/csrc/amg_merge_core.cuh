// Galerkin product P^T A P for a piecewise-constant P without a global sort (round 1 relabelled
// every entry and ran the assembly's radix sort + segmented sum: 29 of 51 ms of the AMG setup at
// 16.7 M rows).  One thread owns one coarse row: it walks the fine rows of its aggregate
// in increasing row order, every row in CSR order, maps the columns through `agg`, and keeps a
// small list sorted by coarse column in its slice of the output (insertion from the back -- the
// columns arrive nearly sorted).  Duplicates are added in arrival order, which is the order the
// stable sort + in-order segmented sum produces, so the coarse values are bit-identical to the
// current path and to tests/amg_mirror.galerkin.  Exact zeros are dropped at the end (DOK
// semantics of csr.cu).  Written __host__ __device__: tests/test_amg_merge_host.py compiles it
// for the CPU and checks it against the numpy statement; amg_host.cuh runs it on the device.
// Row-partitioned setup: `label` maps a LOCAL column (owned or halo) to its coarse column id and
// columns >= ncol_limit are skipped (the local-local block the second pairwise pass looks at).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define AMG_MERGE_HD __host__ __device__ __forceinline__
#else
#define AMG_MERGE_HD static inline
#endif

// Upper bound of the entries of coarse row I: the fine rows' lengths added up.
AMG_MERGE_HD int32_t amg_merge_bound(int32_t I, const int32_t* pt_ptr, const int32_t* pt_idx,
                                     const int32_t* indptr) {
    int32_t total = 0;
    for (int32_t q = pt_ptr[I]; q < pt_ptr[I + 1]; ++q) {
        const int32_t i = pt_idx[q];
        total += indptr[i + 1] - indptr[i];
    }
    return total;
}

// Builds coarse row I in out_cols / out_vals (room for amg_merge_bound entries); returns the
// number of entries kept.
AMG_MERGE_HD int32_t amg_merge_row(int32_t I, const int32_t* pt_ptr, const int32_t* pt_idx,
                                   const int32_t* indptr, const int32_t* indices, const double* data,
                                   const int32_t* label, int32_t* out_cols, double* out_vals,
                                   int32_t ncol_limit = 0x7fffffff) {
    int32_t len = 0;
    for (int32_t q = pt_ptr[I]; q < pt_ptr[I + 1]; ++q) {
        const int32_t i = pt_idx[q];
        for (int32_t p = indptr[i]; p < indptr[i + 1]; ++p) {
            if (indices[p] >= ncol_limit) continue;
            const int32_t c = label[indices[p]];
            const double v = data[p];
            int32_t pos = len;                       // first position whose column is >= c, from the back
            while (pos > 0 && out_cols[pos - 1] >= c) --pos;
            if (pos < len && out_cols[pos] == c) {
                out_vals[pos] = out_vals[pos] + v;   // arrival order == stable-sort order
                continue;
            }
            for (int32_t k = len; k > pos; --k) {
                out_cols[k] = out_cols[k - 1];
                out_vals[k] = out_vals[k - 1];
            }
            out_cols[pos] = c;
            out_vals[pos] = v;
            ++len;
        }
    }
    int32_t kept = 0;
    for (int32_t k = 0; k < len; ++k) {
        if (out_vals[k] != 0.0) {
            out_cols[kept] = out_cols[k];
            out_vals[kept] = out_vals[k];
            ++kept;
        }
    }
    return kept;
}
