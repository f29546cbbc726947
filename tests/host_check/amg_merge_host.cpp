// TEST-ONLY host driver of the merge-based Galerkin product (nodal_b200/csrc/amg_merge_core.cuh): the schedule a
// CUDA version would run -- bounds, exclusive scan, one "thread" per coarse row, compaction.
#include <stdint.h>

#include <vector>

#include "../../nodal_b200/csrc/amg_merge_core.cuh"

// out_indptr[nc + 1]; out_indices / out_data sized for the upper bound (sum of fine nnz).
extern "C" int64_t amg_galerkin_merge_host(int32_t nc, const int32_t* pt_ptr, const int32_t* pt_idx,
                                           const int32_t* indptr, const int32_t* indices,
                                           const double* data, const int32_t* agg, int32_t* out_indptr,
                                           int32_t* out_indices, double* out_data) {
    std::vector<int64_t> start(nc + 1, 0);
    for (int32_t I = 0; I < nc; ++I) start[I + 1] = start[I] + amg_merge_bound(I, pt_ptr, pt_idx, indptr);
    std::vector<int32_t> cols(start[nc] ? start[nc] : 1);
    std::vector<double> vals(start[nc] ? start[nc] : 1);
    std::vector<int32_t> count(nc, 0);
    for (int32_t I = 0; I < nc; ++I)
        count[I] = amg_merge_row(I, pt_ptr, pt_idx, indptr, indices, data, agg, cols.data() + start[I],
                                 vals.data() + start[I]);
    out_indptr[0] = 0;
    for (int32_t I = 0; I < nc; ++I) {
        out_indptr[I + 1] = out_indptr[I] + count[I];
        for (int32_t k = 0; k < count[I]; ++k) {
            out_indices[out_indptr[I] + k] = cols[start[I] + k];
            out_data[out_indptr[I] + k] = vals[start[I] + k];
        }
    }
    return out_indptr[nc];
}
