// TEST-ONLY: compiles the __host__ __device__ stamp core (nodal_b200/csrc/stamp_core.cuh) for
// the CPU so that `pytest -m "not gpu"` can check the stamp arithmetic the CUDA kernel runs
// against the golden vectors without a GPU.  Not part of the product library.
#include <stdint.h>

#include "../../nodal_b200/csrc/stamp_core.cuh"

extern "C" int stamp_table_host(int64_t ncomp, const uint8_t* type, const double* value,
                                const int32_t* a, const int32_t* b, const int32_t* c,
                                const int32_t* d, const int32_t* drv, const int32_t* branch,
                                int32_t kcl, int32_t n, int32_t stride, int32_t* rows,
                                int32_t* cols, double* vals) {
    for (int64_t i = 0; i < ncomp; ++i) {
        StampOut o;
        double dv = 1.0;
        if (type[i] == 5 || type[i] == 6) dv = value[drv[i]];
        stamp_component(type[i], value[i], a[i], b[i], c[i], d[i], dv, branch[i], kcl, n, o);
        if (o.count > stride) return -1;
        for (int k = 0; k < stride; ++k) {
            const bool live = k < o.count;
            rows[i * stride + k] = live ? o.row[k] : n;
            cols[i * stride + k] = live ? o.col[k] : 0;
            vals[i * stride + k] = live ? o.val[k] : 0.0;
        }
    }
    return 0;
}
