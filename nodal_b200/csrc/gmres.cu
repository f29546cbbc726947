// Restarted GMRES(m), FP64, right-preconditioned with the (zero-safe) diagonal.
// Replaces scipy.sparse.linalg.spsolve (nodal/nodal.py:325) for netlists whose voltage /
// controlled sources make G non-symmetric and put zeros on the diagonal of the branch rows
// (nodal/models.py:35-78: write_E / write_VCVS never touch G[r, r]).
//
// Arnoldi with classical Gram-Schmidt applied twice (CGS2): per inner step
//   t = D^-1 v_j ; w = A t                                      (scale + CSR SpMV)
//   h = V^T w ; w -= V h          twice                         (multi-dot + multi-axpy)
//   h_{j+1,j} = ||w|| ; v_{j+1} = w / h_{j+1,j}
// Dot products are two-phase and deterministic: per-block partials, then every block of the
// consumer kernel re-reduces them in a fixed order.  The (m+1) x m Hessenberg least-squares
// problem is tiny and stays on the device: a one-thread kernel applies the Givens rotations after
// every inner step and freezes the state once the residual estimate meets the tolerance, the
// host only polls a control word a few steps behind (round 1: a D2H copy and a stream
// synchronisation per inner step).
#include <math.h>

#include <algorithm>
#include <vector>

#include "sparse.cuh"

constexpr int GM_T = 256;
constexpr int GM_MAXM = 128;

__global__ void __launch_bounds__(GM_T)
gm_diag_kernel(int32_t n, const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices,
               const double* __restrict__ data, double* __restrict__ dinv) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
        double dg = 0.0;
        for (int32_t j = indptr[r]; j < indptr[r + 1]; ++j)
            if (indices[j] == r) dg += data[j];
        dinv[r] = dg != 0.0 ? 1.0 / dg : 1.0;
    }
}

// out = a .* b
__global__ void __launch_bounds__(GM_T)
gm_mul_kernel(int32_t n, const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = a[i] * b[i];
}

// r = b - q ; partial[blk] = sum r^2 (and partial_b[blk] = sum b^2)
__global__ void __launch_bounds__(GM_T)
gm_residual_kernel(int32_t n, const double* __restrict__ b, const double* __restrict__ q,
                   double* __restrict__ r, double* __restrict__ part_rr, double* __restrict__ part_bb) {
    __shared__ double sm[40];
    double rr = 0.0, bb = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double v = b[i] - q[i];
        r[i] = v;
        rr = fma(v, v, rr);
        bb = fma(b[i], b[i], bb);
    }
    rr = block_sum(rr, sm);
    bb = block_sum(bb, sm);
    if (threadIdx.x == 0) { part_rr[blockIdx.x] = rr; part_bb[blockIdx.x] = bb; }
}

// v0 = r / sqrt(sum(part_rr)) ; also publishes the two sums for the host
__global__ void __launch_bounds__(GM_T)
gm_normalize_kernel(int32_t n, const double* __restrict__ w, const double* __restrict__ part, int np,
                    const double* __restrict__ part2, double* __restrict__ v, double* __restrict__ out_h) {
    __shared__ double sm[40];
    const double ss = reduce_partials(part, np, sm);
    const double s2 = part2 ? reduce_partials(part2, np, sm) : 0.0;
    if (blockIdx.x == 0 && threadIdx.x == 0) { out_h[0] = ss; out_h[1] = s2; }
    const double inv = ss > 0.0 ? 1.0 / sqrt(ss) : 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        v[i] = w[i] * inv;
}

// partial[i][blk] = sum over the block's chunk of V_i . w, for i < nv (nv <= GM_MAXM + 1);
// with nv == 0 computes w . w into partial[0][blk].
__global__ void __launch_bounds__(GM_T)
gm_multidot_kernel(int32_t n, const double* __restrict__ V, int nv, const double* __restrict__ w,
                   double* __restrict__ partial, int np) {
    __shared__ double sm[40];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (nv == 0) {
        double s = 0.0;
        for (int64_t i = i0; i < n; i += stride) s = fma(w[i], w[i], s);
        s = block_sum(s, sm);
        if (threadIdx.x == 0) partial[blockIdx.x] = s;
        return;
    }
    for (int k0 = 0; k0 < nv; k0 += 4) {
        double s[4] = {0, 0, 0, 0};
        const int kn = min(4, nv - k0);
        for (int64_t i = i0; i < n; i += stride) {
            const double wi = w[i];
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (k < kn) s[k] = fma(V[(size_t)(k0 + k) * n + i], wi, s[k]);
        }
        for (int k = 0; k < kn; ++k) {
            const double t = block_sum(s[k], sm);
            if (threadIdx.x == 0) partial[(size_t)(k0 + k) * np + blockIdx.x] = t;
        }
    }
}

// h_i = sum(partial[i][*]) ; w -= sum_i h_i V_i ; block 0 stores h (accumulating if accumulate)
__global__ void __launch_bounds__(GM_T)
gm_multiaxpy_kernel(int32_t n, const double* __restrict__ V, int nv, double* __restrict__ w,
                    const double* __restrict__ partial, int np, double* __restrict__ h_out, int accumulate) {
    __shared__ double sm[40];
    __shared__ double h[GM_MAXM + 1];
    for (int k = 0; k < nv; ++k) {
        const double t = reduce_partials(partial + (size_t)k * np, np, sm);
        if (threadIdx.x == 0) h[k] = t;
    }
    __syncthreads();
    if (blockIdx.x == 0)
        for (int k = threadIdx.x; k < nv; k += blockDim.x) h_out[k] = accumulate ? h_out[k] + h[k] : h[k];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double wi = w[i];
        for (int k = 0; k < nv; ++k) wi = fma(-h[k], V[(size_t)k * n + i], wi);
        w[i] = wi;
    }
}

// Hessenberg state of one restart cycle, kept on the device.
struct GmCtl {
    int jdone;       // inner steps whose column has been rotated into H
    int stop;        // 0 running, 1 converged / lucky breakdown, 2 breakdown (NaN or zero column)
    double resid;    // |g[jdone]|: residual norm estimate
};

__global__ void gm_cycle_init_kernel(GmCtl* ctl, double* __restrict__ g, int m, const double* __restrict__ sums) {
    if (threadIdx.x != 0) return;
    ctl->jdone = 0;
    ctl->stop = 0;
    const double beta = sqrt(sums[0]);
    ctl->resid = beta;
    for (int i = 0; i <= m; ++i) g[i] = 0.0;
    g[0] = beta;
}

// Column j of H from the step's dot products (hcol[0..j], squared norm of the new vector at
// hcol[GM_MAXM + 2]): previous rotations, new rotation, g update -- what LAPACK-style GMRES does
// on the host, one thread.
__global__ void gm_givens_kernel(int j, int m, const double* __restrict__ hcol_in, double* __restrict__ H,
                                 double* __restrict__ cs, double* __restrict__ sn, double* __restrict__ g,
                                 GmCtl* ctl, double tol_abs, double tiny) {
    if (threadIdx.x != 0 || ctl->stop) return;
    double* hcol = H + (size_t)j * (m + 1);
    for (int i = 0; i <= j; ++i) hcol[i] = hcol_in[i];
    const double hn = sqrt(hcol_in[GM_MAXM + 2]);
    hcol[j + 1] = hn;
    for (int i = 0; i < j; ++i) {
        const double a = cs[i] * hcol[i] + sn[i] * hcol[i + 1];
        hcol[i + 1] = -sn[i] * hcol[i] + cs[i] * hcol[i + 1];
        hcol[i] = a;
    }
    const double denom = hypot(hcol[j], hcol[j + 1]);
    if (denom == 0.0 || !(denom == denom)) { ctl->stop = 2; return; }
    cs[j] = hcol[j] / denom;
    sn[j] = hcol[j + 1] / denom;
    hcol[j] = denom;
    hcol[j + 1] = 0.0;
    g[j + 1] = -sn[j] * g[j];
    g[j] = cs[j] * g[j];
    ctl->resid = fabs(g[j + 1]);
    ctl->jdone = j + 1;
    if (ctl->resid <= tol_abs || hn <= tiny) ctl->stop = 1;
}

// y = H(0:j,0:j)^-1 g(0:j) for j = jdone, zeros beyond (so x += D^-1 V y can run over every vector
// the host launched)
__global__ void gm_solve_y_kernel(int m, int launched, const double* __restrict__ H, const double* __restrict__ g,
                                  const GmCtl* __restrict__ ctl, double* __restrict__ y) {
    if (threadIdx.x != 0) return;
    const int j = ctl->jdone;
    for (int i = j - 1; i >= 0; --i) {
        double s = g[i];
        for (int k = i + 1; k < j; ++k) s -= H[(size_t)k * (m + 1) + i] * y[k];
        y[i] = s / H[(size_t)i * (m + 1) + i];
    }
    for (int i = j; i < launched; ++i) y[i] = 0.0;
}

// x += D^-1 (V y)
__global__ void __launch_bounds__(GM_T)
gm_update_x_kernel(int32_t n, const double* __restrict__ V, int nv, const double* __restrict__ y,
                   const double* __restrict__ dinv, double* __restrict__ x) {
    __shared__ double ys[GM_MAXM + 1];
    for (int k = threadIdx.x; k < nv; k += blockDim.x) ys[k] = y[k];
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double s = 0.0;
        for (int k = 0; k < nv; ++k) s = fma(ys[k], V[(size_t)k * n + i], s);
        x[i] = fma(dinv[i], s, x[i]);
    }
}

extern "C" int nodal_gmres(nodal_ctx* ctx, int32_t n, int64_t nnz, const int32_t* indptr,
                           const int32_t* indices, const double* data, const double* rhs, double* x,
                           double rtol, int32_t restart, int32_t maxit, int32_t* iters_h,
                           double* relres_h, void* stream) {
    NvtxRange nvtx_range("nodal_gmres");
    if (!ctx || n < 0 || !iters_h || !relres_h) return NODAL_BAD_ARG;
    *iters_h = 0;
    *relres_h = 0.0;
    if (n == 0) return NODAL_OK;
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(cudaSetDevice(ctx->device));
    const int m = std::max(1, std::min({(int)restart, GM_MAXM, (int)n}));
    const int np = (int)std::min<int64_t>((int64_t)ctx->num_sms * 4, ((int64_t)n + GM_T - 1) / GM_T);
    const size_t vec = align_up(sizeof(double) * (size_t)n, 256);
    const size_t need = (size_t)(m + 1) * vec + 4 * vec + align_up(sizeof(double) * (size_t)(m + 2) * np, 256) * 2 +
                        sizeof(double) * 4 * (GM_MAXM + 8) + (1 << 16);
    NODAL_TRY(ctx_reserve(ctx, need));
    double* V = carve<double>(ctx, (size_t)(m + 1) * n);
    double* w = carve<double>(ctx, n);
    double* t = carve<double>(ctx, n);
    double* dinv = carve<double>(ctx, n);
    double* partial = carve<double>(ctx, (size_t)(m + 2) * np);
    double* part2 = carve<double>(ctx, (size_t)np);
    double* hdev = carve<double>(ctx, GM_MAXM + 8);     // h column of the current step
    double* ydev = carve<double>(ctx, GM_MAXM + 8);
    double* sums = carve<double>(ctx, 8);
    if (!V || !w || !t || !dinv || !partial || !part2 || !hdev || !ydev || !sums) return NODAL_CUDA_ERROR;
    double* hhost = reinterpret_cast<double*>(ctx->pinned);   // 4096 B pinned: >= 2*m+8 doubles

    gm_diag_kernel<<<np, GM_T, 0, st>>>(n, indptr, indices, data, dinv);
    KERNEL_CHECK();

    // Hessenberg state on the device (pool: the arena above is full of vectors)
    double* Hd = static_cast<double*>(ctx_pool_alloc(ctx, sizeof(double) * ((size_t)(m + 1) * m + 3 * (size_t)(m + 2)) + 256));
    if (!Hd) return NODAL_CUDA_ERROR;
    double* csd = Hd + (size_t)(m + 1) * m;
    double* snd = csd + (m + 2);
    double* gd = snd + (m + 2);
    GmCtl* ctl = reinterpret_cast<GmCtl*>(gd + (m + 2));
    GmCtl* ctl_h = reinterpret_cast<GmCtl*>(static_cast<char*>(ctx->pinned) + 1024);   // two polling slots
    cudaEvent_t ev_poll[2];
    CUDA_TRY(cudaEventCreateWithFlags(&ev_poll[0], cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&ev_poll[1], cudaEventDisableTiming));
    constexpr int POLL = 4;      // inner steps between two looks at the control word

    double bnorm = -1.0, resid = 0.0;
    int total = 0, status = NODAL_NOT_CONVERGED;
    const int max_cycles = std::max(1, (maxit + m - 1) / m) + 1;
    auto body = [&]() -> int {
        for (int cycle = 0; cycle < max_cycles; ++cycle) {
            // true residual r = b - A x  -> v_0
            NODAL_TRY(csr_spmv_launch(ctx, n, nnz, indptr, indices, data, x, t, st));
            gm_residual_kernel<<<np, GM_T, 0, st>>>(n, rhs, t, w, partial, part2);
            KERNEL_CHECK();
            gm_normalize_kernel<<<np, GM_T, 0, st>>>(n, w, partial, np, part2, V, sums);
            KERNEL_CHECK();
            CUDA_TRY(cudaMemcpyAsync(hhost, sums, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaStreamSynchronize(st));
            const double beta = sqrt(hhost[0]);
            if (bnorm < 0.0) bnorm = sqrt(hhost[1]);
            resid = beta;
            if (!(beta == beta)) { status = NODAL_BREAKDOWN; break; }
            if (bnorm == 0.0) {   // b = 0 -> x = 0
                CUDA_TRY(cudaMemsetAsync(x, 0, sizeof(double) * (size_t)n, st));
                resid = 0.0;
                status = NODAL_OK;
                break;
            }
            if (beta <= rtol * bnorm) { status = NODAL_OK; break; }
            if (total >= maxit) break;
            gm_cycle_init_kernel<<<1, 32, 0, st>>>(ctl, gd, m, sums);
            KERNEL_CHECK();
            const int budget = std::min(m, maxit - total);
            int launched = 0, polls = 0;
            bool stopped = false;
            for (int j = 0; j < budget && !stopped; ++j) {
                const double* vj = V + (size_t)j * n;
                double* vn = V + (size_t)(j + 1) * n;
                gm_mul_kernel<<<np, GM_T, 0, st>>>(n, vj, dinv, t);
                KERNEL_CHECK();
                NODAL_TRY(csr_spmv_launch(ctx, n, nnz, indptr, indices, data, t, w, st));
                for (int pass = 0; pass < 2; ++pass) {
                    gm_multidot_kernel<<<np, GM_T, 0, st>>>(n, V, j + 1, w, partial, np);
                    KERNEL_CHECK();
                    gm_multiaxpy_kernel<<<np, GM_T, 0, st>>>(n, V, j + 1, w, partial, np, hdev, pass);
                    KERNEL_CHECK();
                }
                gm_multidot_kernel<<<np, GM_T, 0, st>>>(n, V, 0, w, partial, np);
                KERNEL_CHECK();
                gm_normalize_kernel<<<np, GM_T, 0, st>>>(n, w, partial, np, nullptr, vn, hdev + GM_MAXM + 2);
                KERNEL_CHECK();
                gm_givens_kernel<<<1, 32, 0, st>>>(j, m, hdev, Hd, csd, snd, gd, ctl, rtol * bnorm, 1e-300 * bnorm);
                KERNEL_CHECK();
                launched = j + 1;
                if (launched % POLL == 0) {
                    // look at the control word one poll behind: the stream never drains
                    CUDA_TRY(cudaMemcpyAsync(&ctl_h[polls & 1], ctl, sizeof(GmCtl), cudaMemcpyDeviceToHost, st));
                    CUDA_TRY(cudaEventRecord(ev_poll[polls & 1], st));
                    if (polls >= 1) {
                        CUDA_TRY(cudaEventSynchronize(ev_poll[(polls - 1) & 1]));
                        if (ctl_h[(polls - 1) & 1].stop) stopped = true;
                    }
                    ++polls;
                }
            }
            // y = H^-1 g over the steps that counted ; x += D^-1 V y
            gm_solve_y_kernel<<<1, 32, 0, st>>>(m, launched, Hd, gd, ctl, ydev);
            KERNEL_CHECK();
            gm_update_x_kernel<<<np, GM_T, 0, st>>>(n, V, launched, ydev, dinv, x);
            KERNEL_CHECK();
            CUDA_TRY(cudaMemcpyAsync(&ctl_h[0], ctl, sizeof(GmCtl), cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaStreamSynchronize(st));
            total += ctl_h[0].jdone;
            resid = ctl_h[0].resid;
            if (ctl_h[0].stop == 2) { status = NODAL_BREAKDOWN; break; }
            if (ctl_h[0].jdone == 0) { status = NODAL_BREAKDOWN; break; }
        }
        return NODAL_OK;
    };
    const int brc = body();
    cudaEventDestroy(ev_poll[0]);
    cudaEventDestroy(ev_poll[1]);
    ctx_pool_free(ctx, Hd);
    NODAL_TRY(brc);
    *iters_h = total;
    *relres_h = bnorm > 0.0 ? resid / bnorm : 0.0;
    if (status == NODAL_NOT_CONVERGED) {
        // the loop may have ended on the iteration budget right after an update: re-check
        NODAL_TRY(csr_spmv_launch(ctx, n, nnz, indptr, indices, data, x, t, st));
        gm_residual_kernel<<<np, GM_T, 0, st>>>(n, rhs, t, w, partial, part2);
        KERNEL_CHECK();
        gm_normalize_kernel<<<np, GM_T, 0, st>>>(n, w, partial, np, part2, V, sums);
        KERNEL_CHECK();
        CUDA_TRY(cudaMemcpyAsync(hhost, sums, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        *relres_h = bnorm > 0.0 ? sqrt(hhost[0]) / bnorm : 0.0;
        if (*relres_h <= rtol) status = NODAL_OK;
    }
    return status;
}
