// Batched small-system solve for parameter sweeps (config C4): every thread stamps its own
// copy of a shared topology from its row of `values` and solves the n x n system with
// partially pivoted Gaussian elimination.  The matrix never touches HBM: it lives in
// shared memory, element-major / thread-minor so that all lanes of a warp hit different
// banks.  Algorithmic traffic is ncomp + n doubles per system.
// Replaces a Python loop of Netlist + Circuit + numpy.linalg.solve (nodal/nodal.py:306-336).
#include <algorithm>

#include "common.cuh"
#include "stamp_core.cuh"

constexpr int BATCH_MAX_N = 24;

__global__ void lu_batched_kernel(int64_t batch, int ncomp, const uint8_t* __restrict__ type,
                                  const int32_t* __restrict__ a, const int32_t* __restrict__ b,
                                  const int32_t* __restrict__ c, const int32_t* __restrict__ d,
                                  const int32_t* __restrict__ drv, const int32_t* __restrict__ branch,
                                  int kcl, int n, const double* __restrict__ values,
                                  double* __restrict__ x, int32_t* __restrict__ info) {
    extern __shared__ double M[];   // [(n) * (n + 1)][blockDim.x]
    const int T = blockDim.x, tid = threadIdx.x;
    const int ld = n + 1;
#define AT(i, j) M[((i) * ld + (j)) * T + tid]
    for (int64_t sys = (int64_t)blockIdx.x * T + tid; sys < batch; sys += (int64_t)gridDim.x * T) {
        for (int e = 0; e < n * ld; ++e) M[e * T + tid] = 0.0;
        const double* val = values + sys * ncomp;
        for (int k = 0; k < ncomp; ++k) {
            const int t = type[k];
            double dv = 1.0;
            if (t == NODAL_T_CCVS || t == NODAL_T_CCCS) dv = val[drv[k]];
            StampOut o;
            stamp_component(t, val[k], a[k], b[k], c[k], d[k], dv, branch[k], kcl, n, o);
            for (int e = 0; e < o.count; ++e) AT(o.row[e], o.col[e]) += o.val[e];   // col == n is the rhs
        }
        int bad = 0;
        for (int k = 0; k < n && !bad; ++k) {
            int p = k;
            double best = fabs(AT(k, k));
            for (int i = k + 1; i < n; ++i) {
                const double v = fabs(AT(i, k));
                if (v > best) { best = v; p = i; }
            }
            if (!(best > 0.0)) { bad = k + 1; break; }
            if (p != k)
                for (int j = k; j <= n; ++j) { const double t0 = AT(k, j); AT(k, j) = AT(p, j); AT(p, j) = t0; }
            const double inv = 1.0 / AT(k, k);
            for (int i = k + 1; i < n; ++i) {
                const double l = AT(i, k) * inv;
                if (l != 0.0)
                    for (int j = k + 1; j <= n; ++j) AT(i, j) = fma(-l, AT(k, j), AT(i, j));
            }
        }
        double* xs = x + sys * n;
        if (bad) {
            for (int i = 0; i < n; ++i) xs[i] = nan("");
        } else {
            for (int i = n - 1; i >= 0; --i) {
                double s = AT(i, n);
                for (int j = i + 1; j < n; ++j) s = fma(-AT(i, j), AT(j, n), s);
                s /= AT(i, i);
                AT(i, n) = s;
                xs[i] = s;
            }
        }
        info[sys] = bad;
    }
#undef AT
}

extern "C" int nodal_lu_batched(nodal_ctx* ctx, int64_t batch, int32_t ncomp, const uint8_t* type,
                                const int32_t* a, const int32_t* b, const int32_t* c,
                                const int32_t* d, const int32_t* drv, const int32_t* branch,
                                int32_t kcl, int32_t n, const double* values, double* x,
                                int32_t* info, void* stream) {
    if (!ctx || batch < 0 || ncomp < 0 || n < 0) return NODAL_BAD_ARG;
    if (n > BATCH_MAX_N) {
        nodal_set_error("nodal_lu_batched: n=%d > %d unknowns per system is not supported", n, BATCH_MAX_N);
        return NODAL_BAD_ARG;
    }
    if (batch == 0 || n == 0) return NODAL_OK;
    CUDA_TRY(cudaSetDevice(ctx->device));
    const size_t per_thread = sizeof(double) * (size_t)n * (n + 1);
    int threads = (int)std::min<size_t>(128, (200 * 1024) / per_thread);
    threads = std::max(32, threads / 32 * 32);
    const size_t smem = per_thread * threads;
    CUDA_TRY(cudaFuncSetAttribute(lu_batched_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t want = (batch + threads - 1) / threads;
    const int grid = (int)std::min<int64_t>(want, (int64_t)ctx->num_sms * 16);
    lu_batched_kernel<<<grid, threads, smem, (cudaStream_t)stream>>>(batch, ncomp, type, a, b, c, d, drv,
                                                                     branch, kcl, n, values, x, info);
    KERNEL_CHECK();
    return NODAL_OK;
}
