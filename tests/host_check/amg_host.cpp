// TEST-ONLY: compiles the __host__ __device__ aggregation rules (nodal_b200/csrc/amg_core.cuh) for
// the CPU and runs one pairwise pass with the kernel schedule of csrc/amg.cu (all proposals, then
// all acceptances, `rounds` times; roots; ids by exclusive scan), so `pytest -m "not gpu"` can
// compare the rules the CUDA kernels execute with their numpy statement (tests/amg_mirror.py).
// Not part of the product library.
#include <stdint.h>

#include <vector>

#include "../../nodal_b200/csrc/amg_core.cuh"

extern "C" uint32_t amg_edge_hash_host(int32_t i, int32_t j) { return amg_edge_hash(i, j); }

extern "C" int32_t amg_aggregate_host(int32_t n, const int32_t* indptr, const int32_t* indices,
                                      const double* data, int32_t rounds, int32_t* match_out,
                                      int32_t* agg_out) {
    std::vector<int32_t> match(n, -1), best(n, -1), root(n, 0);
    std::vector<uint32_t> ids(n, 0);
    for (int32_t r = 0; r < rounds; ++r) {
        for (int32_t i = 0; i < n; ++i)
            best[i] = match[i] >= 0 ? -1 : amg_pick(i, indptr, indices, data, match.data());
        for (int32_t i = 0; i < n; ++i) {
            if (match[i] >= 0) continue;
            const int32_t b = best[i];
            if (b >= 0 && best[b] == i) match[i] = b;
        }
    }
    for (int32_t i = 0; i < n; ++i) root[i] = amg_root(i, indptr, indices, data, match.data());
    uint32_t run = 0;
    for (int32_t i = 0; i < n; ++i) {
        ids[i] = run;
        run += root[i] == i ? 1u : 0u;
    }
    for (int32_t i = 0; i < n; ++i) {
        agg_out[i] = (int32_t)ids[root[i]];
        if (match_out) match_out[i] = match[i];
    }
    return (int32_t)run;
}
