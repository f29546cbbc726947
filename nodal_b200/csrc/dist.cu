// Row-partitioned Jacobi-PCG across GPUs (one process per GPU), NCCL over NVLink/NVSwitch.
//
// Rank k owns the contiguous rows [bounds[k], bounds[k+1]) of G as a local CSR with GLOBAL
// column indices.  Setup (all on device except the tiny per-peer bookkeeping):
//   - the off-rank columns referenced by local rows are compacted, radix-sorted and made
//     unique -> the halo list; columns are renumbered to the local layout
//     [ owned x (nloc) | halo x (nhalo) ];
//   - ranks tell every owner which entries they need (ncclAllGather of the count matrix,
//     grouped ncclSend/ncclRecv of the index lists).
// Iteration = the three single-GPU kernels (pcg_kernels.cuh) plus
//   - one gather of the entries peers need + grouped ncclSend/ncclRecv straight into the
//     halo tail of p (before the SpMV);
//   - two small ncclAllReduce (p.q ; r.z and r.r) of per-rank sums.
// The scalar results are bitwise identical on all ranks, so every rank takes the same
// convergence decision with no extra traffic.  libnccl.so.2 is dlopen'ed: the single-GPU
// library has no NCCL dependency.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <vector>

#include "pcg_kernels.cuh"

namespace {
struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;

int load_nccl() {
    if (g_nccl.handle) return NODAL_OK;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
        nodal_set_error("cannot dlopen libnccl.so.2: %s", dlerror());
        return NODAL_CUDA_ERROR;
    }
#define SYM(field, name)                                                      \
    g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(h, name)); \
    if (!g_nccl.field) {                                                      \
        nodal_set_error("libnccl: missing symbol %s", name);                  \
        return NODAL_CUDA_ERROR;                                              \
    }
    SYM(GetUniqueId, "ncclGetUniqueId")
    SYM(CommInitRank, "ncclCommInitRank")
    SYM(CommDestroy, "ncclCommDestroy")
    SYM(AllReduce, "ncclAllReduce")
    SYM(AllGather, "ncclAllGather")
    SYM(Send, "ncclSend")
    SYM(Recv, "ncclRecv")
    SYM(GroupStart, "ncclGroupStart")
    SYM(GroupEnd, "ncclGroupEnd")
    SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    g_nccl.handle = h;
    return NODAL_OK;
}
}  // namespace

#define NCCL_TRY(expr)                                                                       \
    do {                                                                                     \
        ncclResult_t _r = (expr);                                                            \
        if (_r != ncclSuccess) {                                                             \
            nodal_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,                    \
                            g_nccl.GetErrorString(_r));                                      \
            return NODAL_CUDA_ERROR;                                                         \
        }                                                                                    \
    } while (0)

struct nodal_dist {
    ncclComm_t comm = nullptr;
    int rank = 0, nranks = 1, device = 0;
};

extern "C" int nodal_dist_unique_id(uint8_t* id_h) {
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
    NODAL_TRY(load_nccl());
    ncclUniqueId id;
    NCCL_TRY(g_nccl.GetUniqueId(&id));
    memcpy(id_h, &id, sizeof(id));
    return NODAL_OK;
}

extern "C" int nodal_dist_create(nodal_ctx* ctx, const uint8_t* id_h, int32_t rank, int32_t nranks,
                                 nodal_dist** out) {
    if (!ctx || !id_h || !out || rank < 0 || rank >= nranks) return NODAL_BAD_ARG;
    NODAL_TRY(load_nccl());
    CUDA_TRY(cudaSetDevice(ctx->device));
    ncclUniqueId id;
    memcpy(&id, id_h, sizeof(id));
    nodal_dist* d = new nodal_dist();
    d->rank = rank;
    d->nranks = nranks;
    d->device = ctx->device;
    ncclResult_t r = g_nccl.CommInitRank(&d->comm, nranks, id, rank);
    if (r != ncclSuccess) {
        nodal_set_error("ncclCommInitRank -> %s", g_nccl.GetErrorString(r));
        delete d;
        return NODAL_CUDA_ERROR;
    }
    *out = d;
    return NODAL_OK;
}

extern "C" int nodal_dist_destroy(nodal_dist* d) {
    if (!d) return NODAL_OK;
    cudaSetDevice(d->device);
    if (d->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(d->comm);
    delete d;
    return NODAL_OK;
}

// ---------------------------------------------------------------- setup kernels
constexpr int DT = 256;

__global__ void __launch_bounds__(DT)
dist_flag_external_kernel(int64_t nnz, const int32_t* __restrict__ cols, int32_t rb, int32_t re,
                          u32* __restrict__ flag) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += (int64_t)gridDim.x * blockDim.x) {
        const int32_t c = cols[i];
        flag[i] = (c < rb || c >= re) ? 1u : 0u;
    }
}

__global__ void __launch_bounds__(DT)
dist_compact_external_kernel(int64_t nnz, const int32_t* __restrict__ cols, int32_t rb, int32_t re,
                             const u32* __restrict__ pos, u64* __restrict__ keys) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += (int64_t)gridDim.x * blockDim.x) {
        const int32_t c = cols[i];
        if (c < rb || c >= re) keys[pos[i]] = (u64)(u32)c;
    }
}

__global__ void __launch_bounds__(DT)
dist_unique_flag_kernel(int64_t m, const u64* __restrict__ keys, u32* __restrict__ flag) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x)
        flag[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u;
}

__global__ void __launch_bounds__(DT)
dist_unique_compact_kernel(int64_t m, const u64* __restrict__ keys, const u32* __restrict__ pos,
                           int32_t* __restrict__ halo) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x)
        if (i == 0 || keys[i] != keys[i - 1]) halo[pos[i]] = (int32_t)keys[i];
}

// global column -> local layout [owned | halo]
__global__ void __launch_bounds__(DT)
dist_remap_kernel(int64_t nnz, const int32_t* __restrict__ cols, int32_t rb, int32_t re,
                  const int32_t* __restrict__ halo, int32_t nhalo, int32_t* __restrict__ out) {
    const int32_t nloc = re - rb;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += (int64_t)gridDim.x * blockDim.x) {
        const int32_t c = cols[i];
        if (c >= rb && c < re) { out[i] = c - rb; continue; }
        int lo = 0, hi = nhalo - 1;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (halo[mid] < c) lo = mid + 1; else hi = mid;
        }
        out[i] = nloc + lo;
    }
}

__global__ void __launch_bounds__(DT)
dist_rebase_kernel(int64_t m, int32_t* __restrict__ idx, int32_t rb) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x)
        idx[i] -= rb;
}

__global__ void __launch_bounds__(DT)
dist_gather_kernel(int64_t m, const int32_t* __restrict__ idx, const double* __restrict__ src,
                   double* __restrict__ dst) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x)
        dst[i] = src[idx[i]];
}

// out[k] = deterministic sum of partial array k (up to 3 arrays)
__global__ void __launch_bounds__(PCG_THREADS)
dist_reduce_kernel(const double* __restrict__ a, const double* __restrict__ b, const double* __restrict__ c,
                   int count, double* __restrict__ out) {
    __shared__ double sm[40];
    const double sa = reduce_partials(a, count, sm);
    const double sb = b ? reduce_partials(b, count, sm) : 0.0;
    const double sc = c ? reduce_partials(c, count, sm) : 0.0;
    if (threadIdx.x == 0) {
        out[0] = sa;
        if (b) out[1] = sb;
        if (c) out[2] = sc;
    }
}

static int grid_of(nodal_ctx* ctx, int64_t work) {
    int64_t b = (work + DT - 1) / DT;
    const int64_t cap = (int64_t)ctx->num_sms * 8;
    return (int)std::max<int64_t>(1, std::min(b, cap));
}

extern "C" int nodal_dist_pcg(nodal_ctx* ctx, nodal_dist* d, int32_t n_global, const int32_t* bounds_h,
                              int64_t nnz, const int32_t* indptr, const int32_t* indices,
                              const double* data, const double* rhs_local, double* x_local,
                              double rtol, int32_t maxit, int32_t* iters_h, double* relres_h,
                              double* stats_h, void* stream) {
    if (!ctx || !d || !bounds_h || !iters_h || !relres_h) return NODAL_BAD_ARG;
    *iters_h = 0;
    *relres_h = 0.0;
    if (stats_h) memset(stats_h, 0, 16 * sizeof(double));
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(cudaSetDevice(ctx->device));
    const int R = d->nranks, me = d->rank;
    const int32_t rb = bounds_h[me], re = bounds_h[me + 1];
    const int32_t nloc = re - rb;
    if (nloc <= 0 || bounds_h[0] != 0 || bounds_h[R] != n_global) {
        nodal_set_error("nodal_dist_pcg: every rank must own at least one row and bounds must span [0, n)");
        return NODAL_BAD_ARG;
    }
    cudaEvent_t ev0, ev1, ev2, ev_poll[2];
    CUDA_TRY(cudaEventCreate(&ev0));
    CUDA_TRY(cudaEventCreate(&ev1));
    CUDA_TRY(cudaEventCreate(&ev2));
    CUDA_TRY(cudaEventCreateWithFlags(&ev_poll[0], cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&ev_poll[1], cudaEventDisableTiming));
    nodal_sell* sell = nullptr;
    std::vector<void*> owned;   // cudaMalloc'ed buffers that outlive the arena resets
    auto dmalloc = [&](size_t bytes) -> void* {
        void* p = nullptr;
        if (cudaMalloc(&p, std::max<size_t>(bytes, 256)) != cudaSuccess) return nullptr;
        owned.push_back(p);
        return p;
    };
    PcgDev host{};
    int restarts = 0;
    float ms_setup = 0.f, ms_solve = 0.f;
    int64_t halo_total = 0, send_total = 0;

    auto run = [&]() -> int {
        CUDA_TRY(cudaEventRecord(ev0, st));
        // ---------------- halo discovery (device) ----------------
        const size_t need = align_up((size_t)nnz * 4, 256) * 2 + align_up((size_t)nnz * 8, 256) * 4 +
                            radix_sort_scratch_bytes(std::max<int64_t>(nnz, 1)) +
                            2 * scan_scratch_bytes(std::max<int64_t>(nnz, 1)) + (1 << 16);
        NODAL_TRY(ctx_reserve(ctx, need));
        u32* flag = carve<u32>(ctx, (size_t)std::max<int64_t>(nnz, 1));
        u32* tot = carve<u32>(ctx, 16);
        if (!flag || !tot) return NODAL_CUDA_ERROR;
        u32* host_tot = reinterpret_cast<u32*>(ctx->pinned);
        const int gnz = grid_of(ctx, nnz);
        int64_t next = 0, nhalo = 0;
        int32_t* halo = nullptr;
        if (nnz > 0) {
            dist_flag_external_kernel<<<gnz, DT, 0, st>>>(nnz, indices, rb, re, flag);
            KERNEL_CHECK();
            const size_t mark = ctx->arena_used;
            NODAL_TRY(scan_exclusive_u32(ctx, flag, flag, nnz, tot, st));
            ctx->arena_used = mark;
            CUDA_TRY(cudaMemcpyAsync(host_tot, tot, sizeof(u32), cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaStreamSynchronize(st));
            next = host_tot[0];
        }
        if (next > 0) {
            u64* keys = carve<u64>(ctx, (size_t)next);
            u64* vals = carve<u64>(ctx, (size_t)next);
            u64* keys_alt = carve<u64>(ctx, (size_t)next);
            u64* vals_alt = carve<u64>(ctx, (size_t)next);
            u32* uflag = carve<u32>(ctx, (size_t)next);
            if (!keys || !vals || !keys_alt || !vals_alt || !uflag) return NODAL_CUDA_ERROR;
            dist_compact_external_kernel<<<gnz, DT, 0, st>>>(nnz, indices, rb, re, flag, keys);
            KERNEL_CHECK();
            int bits = 1;
            while (((int64_t)n_global >> bits) != 0) ++bits;
            bool in_alt = false;
            const size_t mark = ctx->arena_used;
            NODAL_TRY(radix_sort_pairs(ctx, keys, vals, keys_alt, vals_alt, next, bits, &in_alt, st));
            ctx->arena_used = mark;
            const u64* sk = in_alt ? keys_alt : keys;
            dist_unique_flag_kernel<<<grid_of(ctx, next), DT, 0, st>>>(next, sk, uflag);
            KERNEL_CHECK();
            NODAL_TRY(scan_exclusive_u32(ctx, uflag, uflag, next, tot, st));
            ctx->arena_used = mark;
            CUDA_TRY(cudaMemcpyAsync(host_tot, tot, sizeof(u32), cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaStreamSynchronize(st));
            nhalo = host_tot[0];
            halo = static_cast<int32_t*>(dmalloc(sizeof(int32_t) * (size_t)nhalo));
            if (!halo) return NODAL_CUDA_ERROR;
            dist_unique_compact_kernel<<<grid_of(ctx, next), DT, 0, st>>>(next, sk, uflag, halo);
            KERNEL_CHECK();
        }
        halo_total = nhalo;
        int32_t* lcols = static_cast<int32_t*>(dmalloc(sizeof(int32_t) * (size_t)std::max<int64_t>(nnz, 1)));
        if (!lcols) return NODAL_CUDA_ERROR;
        if (nnz > 0) {
            dist_remap_kernel<<<gnz, DT, 0, st>>>(nnz, indices, rb, re, halo, (int32_t)nhalo, lcols);
            KERNEL_CHECK();
        }
        // ---------------- who needs what (host bookkeeping of <= R counts) ----------------
        std::vector<int32_t> halo_h((size_t)nhalo);
        if (nhalo) CUDA_TRY(cudaMemcpyAsync(halo_h.data(), halo, sizeof(int32_t) * (size_t)nhalo, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        std::vector<int32_t> need_from(R, 0), need_off(R + 1, 0);
        {
            int o = 0;
            for (int64_t i = 0; i < nhalo; ++i) {
                while (halo_h[i] >= bounds_h[o + 1]) ++o;
                need_from[o]++;
            }
            for (int o2 = 0; o2 < R; ++o2) need_off[o2 + 1] = need_off[o2] + need_from[o2];
        }
        int32_t* cnt_dev = static_cast<int32_t*>(dmalloc(sizeof(int32_t) * (size_t)R * (R + 1)));
        if (!cnt_dev) return NODAL_CUDA_ERROR;
        CUDA_TRY(cudaMemcpyAsync(cnt_dev + (size_t)R * R, need_from.data(), sizeof(int32_t) * R, cudaMemcpyHostToDevice, st));
        NCCL_TRY(g_nccl.AllGather(cnt_dev + (size_t)R * R, cnt_dev, R, ncclInt32, d->comm, st));
        std::vector<int32_t> cnt((size_t)R * R);
        CUDA_TRY(cudaMemcpyAsync(cnt.data(), cnt_dev, sizeof(int32_t) * (size_t)R * R, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        // cnt[s*R + o] = entries rank s needs from owner o
        std::vector<int32_t> send_cnt(R, 0), send_off(R + 1, 0);
        for (int s = 0; s < R; ++s) send_cnt[s] = cnt[(size_t)s * R + me];
        for (int s = 0; s < R; ++s) send_off[s + 1] = send_off[s] + send_cnt[s];
        send_total = send_off[R];
        int32_t* send_idx = static_cast<int32_t*>(dmalloc(sizeof(int32_t) * (size_t)std::max<int64_t>(send_total, 1)));
        double* send_buf = static_cast<double*>(dmalloc(sizeof(double) * (size_t)std::max<int64_t>(send_total, 1)));
        if (!send_idx || !send_buf) return NODAL_CUDA_ERROR;
        NCCL_TRY(g_nccl.GroupStart());
        for (int o = 0; o < R; ++o) {
            if (o == me) continue;
            if (need_from[o]) NCCL_TRY(g_nccl.Send(halo + need_off[o], need_from[o], ncclInt32, o, d->comm, st));
            if (send_cnt[o]) NCCL_TRY(g_nccl.Recv(send_idx + send_off[o], send_cnt[o], ncclInt32, o, d->comm, st));
        }
        NCCL_TRY(g_nccl.GroupEnd());
        if (send_total) {
            dist_rebase_kernel<<<grid_of(ctx, send_total), DT, 0, st>>>(send_total, send_idx, rb);
            KERNEL_CHECK();
        }
        // ---------------- local operator in the solver-private layout ----------------
        NODAL_TRY(sell_from_csr(ctx, nloc, nnz, indptr, lcols, data, &sell, st));   // resets the arena
        Mat A;
        A.n = nloc; A.nnz = nnz; A.indptr = indptr; A.indices = lcols; A.data = data;
        const double mean = (double)nnz / nloc;
        A.tpr = mean <= 2.5 ? 2 : mean <= 6.0 ? 4 : mean <= 12.0 ? 8 : mean <= 24.0 ? 16 : 32;
        const bool use_sell = (double)sell->padded <= 1.5 * (double)nnz + 1024.0;
        if (use_sell) A.sell = sell;
        const int g2 = (int)std::min<int64_t>((int64_t)ctx->num_sms * 8,
                                              std::max<int64_t>(1, ((nloc >> 1) + PCG_THREADS - 1) / PCG_THREADS));
        if (A.sell) {
            const int64_t want = ((int64_t)sell->nslices * 32 + PCG_THREADS - 1) / PCG_THREADS;
            A.g1 = (int)std::min<int64_t>((int64_t)ctx->num_sms * 4, want);
        } else {
            const int64_t want = ((int64_t)nloc * A.tpr + PCG_THREADS - 1) / PCG_THREADS;
            A.g1 = (int)std::min<int64_t>((int64_t)ctx->num_sms * 8, want);
        }
        const int gmax = std::max(A.g1, g2);
        const size_t vloc = align_up(sizeof(double) * (size_t)nloc, 256);
        const size_t vext = align_up(sizeof(double) * (size_t)(nloc + nhalo + 2), 256);
        NODAL_TRY(ctx_reserve(ctx, 3 * vloc + 2 * vext + 8 * align_up(sizeof(double) * gmax, 256) + 8192));
        double* r = carve<double>(ctx, nloc);
        double* q = carve<double>(ctx, nloc);
        double* p = carve<double>(ctx, (size_t)nloc + nhalo + 2);      // [owned | halo]
        double* xe = carve<double>(ctx, (size_t)nloc + nhalo + 2);     // x in the same layout (residual checks)
        double* part_pq = carve<double>(ctx, gmax);
        double* part_rz = carve<double>(ctx, gmax);
        double* part_rr = carve<double>(ctx, gmax);
        double* part_bb = carve<double>(ctx, gmax);
        double* S = carve<double>(ctx, 16);    // S[par*3 + {pq, rz, rr}], S[8..10] start sums
        PcgDev* dev = carve<PcgDev>(ctx, 1);
        double* dinv_own = A.sell ? nullptr : carve<double>(ctx, nloc);
        if (!r || !q || !p || !xe || !part_bb || !S || !dev) return NODAL_CUDA_ERROR;
        const double* dinv = A.sell ? sell->dinv : dinv_own;
        if (dinv_own) {
            csr_dinv_kernel<<<g2, PCG_THREADS, 0, st>>>(nloc, indptr, lcols, data, dinv_own);
            KERNEL_CHECK();
        }
        CUDA_TRY(cudaMemsetAsync(dev, 0, sizeof(PcgDev), st));
        CUDA_TRY(cudaMemsetAsync(S, 0, sizeof(double) * 16, st));

        auto exchange = [&](double* v) -> int {   // fills v[nloc .. nloc+nhalo) from the owners
            if (R == 1) return NODAL_OK;
            if (send_total) {
                dist_gather_kernel<<<grid_of(ctx, send_total), DT, 0, st>>>(send_total, send_idx, v, send_buf);
                KERNEL_CHECK();
            }
            NCCL_TRY(g_nccl.GroupStart());
            for (int o = 0; o < R; ++o) {
                if (o == me) continue;
                if (send_cnt[o]) NCCL_TRY(g_nccl.Send(send_buf + send_off[o], send_cnt[o], ncclFloat64, o, d->comm, st));
                if (need_from[o]) NCCL_TRY(g_nccl.Recv(v + nloc + need_off[o], need_from[o], ncclFloat64, o, d->comm, st));
            }
            NCCL_TRY(g_nccl.GroupEnd());
            return NODAL_OK;
        };
        auto spmv_plain = [&](const double* in, double* out) -> int {
            if (A.sell) return nodal_sell_spmv(ctx, A.sell, in, out, st);
            return csr_spmv_launch(ctx, nloc, nnz, indptr, lcols, data, in, out, st);
        };
        auto start = [&](int first) -> int {
            CUDA_TRY(cudaMemcpyAsync(xe, x_local, sizeof(double) * (size_t)nloc, cudaMemcpyDeviceToDevice, st));
            NODAL_TRY(exchange(xe));
            NODAL_TRY(spmv_plain(xe, q));
            pcg_start_kernel<<<g2, PCG_THREADS, 0, st>>>(nloc, rhs_local, q, dinv, r, p, part_rz, part_rr, part_bb);
            KERNEL_CHECK();
            dist_reduce_kernel<<<1, PCG_THREADS, 0, st>>>(part_rz, part_rr, part_bb, g2, S + 8);
            KERNEL_CHECK();
            NCCL_TRY(g_nccl.AllReduce(S + 8, S + 8, 3, ncclFloat64, ncclSum, d->comm, st));
            // rz of the "previous" iteration lives in the odd-parity slot
            CUDA_TRY(cudaMemcpyAsync(S + 3 + 1, S + 8, sizeof(double), cudaMemcpyDeviceToDevice, st));
            pcg_scalars_kernel<<<1, PCG_THREADS, 0, st>>>(dev, S + 9, S + 10, 1, rtol, maxit, first);
            KERNEL_CHECK();
            return NODAL_OK;
        };
        auto iteration = [&](int par) -> int {
            double* Sp = S + par * 3;
            double* Sq = S + (par ^ 1) * 3;
            NODAL_TRY(exchange(p));
            NODAL_TRY(launch_k1(A, dev, p, q, part_pq, st));
            dist_reduce_kernel<<<1, PCG_THREADS, 0, st>>>(part_pq, nullptr, nullptr, A.g1, Sp);
            KERNEL_CHECK();
            NCCL_TRY(g_nccl.AllReduce(Sp, Sp, 1, ncclFloat64, ncclSum, d->comm, st));
            pcg_update_kernel<<<g2, PCG_THREADS, 0, st>>>(dev, nloc, Sp, 1, Sq + 1, 1, x_local, p, r, q, dinv,
                                                         part_rz, part_rr);
            KERNEL_CHECK();
            dist_reduce_kernel<<<1, PCG_THREADS, 0, st>>>(part_rz, part_rr, nullptr, g2, Sp + 1);
            KERNEL_CHECK();
            NCCL_TRY(g_nccl.AllReduce(Sp + 1, Sp + 1, 2, ncclFloat64, ncclSum, d->comm, st));
            pcg_direction_kernel<<<g2, PCG_THREADS, 0, st>>>(dev, nloc, Sq + 1, Sp + 1, Sp + 2, 1, p, r, dinv);
            KERNEL_CHECK();
            return NODAL_OK;
        };

        NODAL_TRY(start(1));
        CUDA_TRY(cudaEventRecord(ev1, st));
        PcgDev* poll = reinterpret_cast<PcgDev*>(ctx->pinned);
        double last_true_rr = -1.0;
        const int CH = 32;
        for (;;) {
            const int64_t max_chunks = (int64_t)maxit / CH + 3;
            int64_t k = 0;
            for (;; ++k) {
                for (int i = 0; i < CH; ++i) NODAL_TRY(iteration(i & 1));
                CUDA_TRY(cudaMemcpyAsync(&poll[k & 1], dev, sizeof(PcgDev), cudaMemcpyDeviceToHost, st));
                CUDA_TRY(cudaEventRecord(ev_poll[k & 1], st));
                if (k >= 1) {
                    CUDA_TRY(cudaEventSynchronize(ev_poll[(k - 1) & 1]));
                    if (poll[(k - 1) & 1].done) break;
                }
                if (k > max_chunks) break;
            }
            CUDA_TRY(cudaStreamSynchronize(st));
            host = poll[k & 1];
            if (!host.done) host.status = NODAL_NOT_CONVERGED;
            if (host.status == NODAL_BREAKDOWN) break;
            NODAL_TRY(start(0));
            CUDA_TRY(cudaMemcpyAsync(&poll[0], dev, sizeof(PcgDev), cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaStreamSynchronize(st));
            const int recurrence_status = host.status;
            host = poll[0];
            if (host.rr <= host.tol2) { host.status = NODAL_OK; break; }
            if (recurrence_status == NODAL_NOT_CONVERGED || host.iters >= host.maxit) {
                host.status = NODAL_NOT_CONVERGED;
                break;
            }
            if (restarts >= 8 || (last_true_rr >= 0.0 && host.rr > 0.25 * last_true_rr)) {
                host.status = NODAL_NOT_CONVERGED;
                break;
            }
            last_true_rr = host.rr;
            ++restarts;
        }
        CUDA_TRY(cudaEventRecord(ev2, st));
        CUDA_TRY(cudaEventSynchronize(ev2));
        CUDA_TRY(cudaEventElapsedTime(&ms_setup, ev0, ev1));
        CUDA_TRY(cudaEventElapsedTime(&ms_solve, ev1, ev2));
        *iters_h = host.iters;
        *relres_h = host.bb > 0.0 ? sqrt(host.rr / host.bb) : 0.0;
        if (stats_h) {
            stats_h[0] = host.iters;
            stats_h[1] = *relres_h;
            stats_h[2] = restarts;
            stats_h[3] = ms_solve;
            stats_h[4] = ms_setup;
            stats_h[5] = A.sell ? 1.0 : 0.0;
            stats_h[6] = A.sell ? (double)sell->padded : (double)nnz;
            stats_h[7] = A.g1;
            stats_h[12] = (double)halo_total;
            stats_h[13] = (double)send_total;
        }
        return host.status;
    };
    const int rc = run();
    cudaStreamSynchronize(st);
    if (sell) sell_free(sell);
    for (void* p : owned) cudaFree(p);
    cudaEventDestroy(ev0); cudaEventDestroy(ev1); cudaEventDestroy(ev2);
    cudaEventDestroy(ev_poll[0]); cudaEventDestroy(ev_poll[1]);
    return rc;
}
