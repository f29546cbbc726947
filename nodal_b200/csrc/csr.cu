// CSR build from keyed COO triples: radix sort + in-order segmented reduce.
// Replaces scipy's dok_matrix accumulation and G.tocsr() (nodal/nodal.py:349-351,396-397).
//
// Bit-exactness contract: duplicates of one (row, col) are summed left to right in
// emission (= component) order, which is the order the reference's `G[i, j] += g`
// statements run in; entries whose sum is exactly zero are removed, as scipy's DOK
// deletes a key when a falsy value is stored (scipy/sparse/_dok.py).
#include "common.cuh"

constexpr int CB_THREADS = 256;

__global__ void __launch_bounds__(CB_THREADS)
mark_heads_kernel(const u64* __restrict__ keys, int64_t count, int32_t n, int colbits,
                  u32* __restrict__ head) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count;
         i += (int64_t)gridDim.x * blockDim.x) {
        const u64 k = keys[i];
        const bool live = (int64_t)(k >> colbits) < (int64_t)n;
        head[i] = (live && (i == 0 || keys[i - 1] != k)) ? 1u : 0u;
    }
}

// One thread per segment head: sequential in-order sum of the run.
__global__ void __launch_bounds__(CB_THREADS)
segment_sum_kernel(const u64* __restrict__ keys, const double* __restrict__ vals, int64_t count,
                   int32_t n, int colbits, const u32* __restrict__ head_scan,
                   u64* __restrict__ ukey, double* __restrict__ uval, u32* __restrict__ keep,
                   double* __restrict__ rhs) {
    const u64 colmask = ((u64)1 << colbits) - 1;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count;
         i += (int64_t)gridDim.x * blockDim.x) {
        const u64 k = keys[i];
        if ((int64_t)(k >> colbits) >= (int64_t)n) continue;
        if (i != 0 && keys[i - 1] == k) continue;
        double s = vals[i];
        for (int64_t j = i + 1; j < count && keys[j] == k; ++j) s = s + vals[j];
        const u32 seg = head_scan[i];
        const int32_t col = (int32_t)(k & colmask);
        ukey[seg] = k;
        uval[seg] = s;
        if (col == n) {
            rhs[(int32_t)(k >> colbits)] = s;
            keep[seg] = 0u;
        } else {
            keep[seg] = (s != 0.0) ? 1u : 0u;
        }
    }
}

// First-touch variant (the column order scipy's DOK -> CSR conversion leaves, nodal/nodal.py:396-397
// before spsolve sorts the indices in place): the sort's payload is the emission index, values are
// gathered through it, and every unique entry remembers the emission at which its key was
// (re-)inserted -- a DOK key whose running sum hits exact zero is deleted and a later `+=`
// re-inserts it at the END of the dict order.
__global__ void __launch_bounds__(CB_THREADS)
iota_u64_kernel(int64_t count, u64* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = (u64)i;
}

__global__ void __launch_bounds__(CB_THREADS)
segment_sum_first_touch_kernel(const u64* __restrict__ keys, const u64* __restrict__ idx, const double* __restrict__ vals,
                               int64_t count, int32_t n, int colbits, const u32* __restrict__ head_scan,
                               u64* __restrict__ ukey, double* __restrict__ uval, u32* __restrict__ keep,
                               u32* __restrict__ ufirst, double* __restrict__ rhs) {
    const u64 colmask = ((u64)1 << colbits) - 1;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count;
         i += (int64_t)gridDim.x * blockDim.x) {
        const u64 k = keys[i];
        if ((int64_t)(k >> colbits) >= (int64_t)n) continue;
        if (i != 0 && keys[i - 1] == k) continue;
        double s = 0.0;
        bool absent = true;
        u32 first = 0;
        for (int64_t j = i; j < count && keys[j] == k; ++j) {
            const double t = s + vals[idx[j]];
            if (t != 0.0) {
                if (absent) { first = (u32)idx[j]; absent = false; }
            } else {
                absent = true;
            }
            s = absent ? 0.0 : t;
        }
        const u32 seg = head_scan[i];
        const int32_t col = (int32_t)(k & colmask);
        ukey[seg] = k;
        uval[seg] = s;
        ufirst[seg] = first;
        if (col == n) {
            rhs[(int32_t)(k >> colbits)] = s;
            keep[seg] = 0u;
        } else {
            keep[seg] = absent ? 0u : 1u;
        }
    }
}

// kept entries in sorted order -> key (row, first-touch emission) and their position
__global__ void __launch_bounds__(CB_THREADS)
first_touch_keys_kernel(const u64* __restrict__ ukey, const u32* __restrict__ ufirst, const u32* __restrict__ keep_scan,
                        const u32* __restrict__ keep, int64_t useg, int colbits, u64* __restrict__ keys2,
                        u64* __restrict__ pos2) {
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < useg; s += (int64_t)gridDim.x * blockDim.x) {
        if (!keep[s]) continue;
        const u32 p = keep_scan[s];
        keys2[p] = ((ukey[s] >> colbits) << 32) | (u64)ufirst[s];
        pos2[p] = (u64)p;
    }
}

__global__ void __launch_bounds__(CB_THREADS)
permute_entries_kernel(int64_t nnz, const u64* __restrict__ pos, const int32_t* __restrict__ cols_in,
                       const double* __restrict__ vals_in, int32_t* __restrict__ cols_out, double* __restrict__ vals_out) {
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nnz; q += (int64_t)gridDim.x * blockDim.x) {
        const u64 p = pos[q];
        cols_out[q] = cols_in[p];
        vals_out[q] = vals_in[p];
    }
}

__global__ void __launch_bounds__(CB_THREADS)
compact_kernel(const u64* __restrict__ ukey, const double* __restrict__ uval,
               const u32* __restrict__ keep_scan, const u32* __restrict__ keep_flag_src,
               int64_t useg, int colbits, int32_t* __restrict__ indices, double* __restrict__ data,
               int32_t* __restrict__ krow) {
    const u64 colmask = ((u64)1 << colbits) - 1;
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < useg;
         s += (int64_t)gridDim.x * blockDim.x) {
        if (!keep_flag_src[s]) continue;
        const u32 p = keep_scan[s];
        const u64 k = ukey[s];
        indices[p] = (int32_t)(k & colmask);
        data[p] = uval[s];
        krow[p] = (int32_t)(k >> colbits);
    }
}

// indptr[q] = position of the first kept entry with row >= q.  Short runs of empty rows are
// filled by the thread that sees the row change; long runs (a rank of the multi-GPU path owns
// 1/R of the rows of a global-shaped matrix) go to a gap list that a second kernel fills in
// parallel.
struct RowGap { int32_t first, last, value; int32_t pad; };
constexpr int ROWPTR_INLINE = 32;
constexpr int ROWPTR_MAX_GAPS = 4096;

__device__ __forceinline__ void fill_rows(int32_t first, int32_t last, int32_t value,
                                          int32_t* __restrict__ indptr, RowGap* gaps, unsigned int* ngaps) {
    if (last - first < ROWPTR_INLINE) {
        for (int32_t q = first; q <= last; ++q) indptr[q] = value;
        return;
    }
    const unsigned int slot = atomicAdd(ngaps, 1u);
    if (slot < ROWPTR_MAX_GAPS) { gaps[slot].first = first; gaps[slot].last = last; gaps[slot].value = value; }
    else for (int32_t q = first; q <= last; ++q) indptr[q] = value;   // list full: do it here
}

__global__ void __launch_bounds__(CB_THREADS)
row_ptr_kernel(const int32_t* __restrict__ krow, int64_t nnz, int32_t n,
               int32_t* __restrict__ indptr, RowGap* gaps, unsigned int* ngaps) {
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < nnz;
         p += (int64_t)gridDim.x * blockDim.x) {
        const int32_t r = krow[p];
        const int32_t rp = p ? krow[p - 1] : -1;
        if (r != rp) fill_rows(rp + 1, r, (int32_t)p, indptr, gaps, ngaps);
        if (p == nnz - 1) fill_rows(r + 1, n, (int32_t)nnz, indptr, gaps, ngaps);
    }
}

__global__ void __launch_bounds__(CB_THREADS)
row_ptr_gaps_kernel(int32_t* __restrict__ indptr, const RowGap* __restrict__ gaps,
                    const unsigned int* __restrict__ ngaps) {
    const unsigned int count = min(*ngaps, (unsigned int)ROWPTR_MAX_GAPS);
    for (unsigned int g = 0; g < count; ++g) {
        const RowGap gp = gaps[g];
        for (int64_t q = (int64_t)gp.first + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q <= gp.last;
             q += (int64_t)gridDim.x * blockDim.x)
            indptr[q] = gp.value;
    }
}

static int grid_for(nodal_ctx* ctx, int64_t work, int threads) {
    int64_t b = (work + threads - 1) / threads;
    int64_t cap = (int64_t)ctx->num_sms * 16;
    if (b < 1) b = 1;
    return (int)(b < cap ? b : cap);
}

// state kept between nodal_csr_build and nodal_csr_fetch (pointers into the arena)
struct PendingCsr {
    nodal_ctx* ctx = nullptr;
    const u64* ukey = nullptr;
    const double* uval = nullptr;
    const u32* keep = nullptr;
    const u32* keep_scan = nullptr;
    int32_t* krow = nullptr;
    const u32* ufirst = nullptr;      // first-touch order only
    int order = 0;
    int64_t useg = 0, nnz = 0;
    int32_t n = 0;
    int colbits = 0;
    uint64_t generation = 0;
};
static thread_local PendingCsr g_pending;

static int csr_build_impl(nodal_ctx* ctx, int32_t n, int64_t nslots, int32_t colbits, uint64_t* keys_, double* vals,
                          double* rhs, int64_t* nnz_h, void* stream, int order);

extern "C" int nodal_csr_build(nodal_ctx* ctx, int32_t n, int64_t nslots, int32_t colbits,
                               uint64_t* keys_, double* vals, double* rhs, int64_t* nnz_h,
                               void* stream) {
    return csr_build_impl(ctx, n, nslots, colbits, keys_, vals, rhs, nnz_h, stream, 0);
}

extern "C" int nodal_csr_build_ordered(nodal_ctx* ctx, int32_t n, int64_t nslots, int32_t colbits,
                                       uint64_t* keys_, double* vals, int32_t order, double* rhs,
                                       int64_t* nnz_h, void* stream) {
    if (order != 0 && order != 1) return NODAL_BAD_ARG;
    return csr_build_impl(ctx, n, nslots, colbits, keys_, vals, rhs, nnz_h, stream, order);
}

static int csr_build_impl(nodal_ctx* ctx, int32_t n, int64_t nslots, int32_t colbits, uint64_t* keys_, double* vals,
                          double* rhs, int64_t* nnz_h, void* stream, int order) {
    NvtxRange nvtx_range("nodal_csr_build");
    if (!ctx || n < 0 || nslots < 0 || !nnz_h) return NODAL_BAD_ARG;
    if (nslots >= ((int64_t)1 << 31)) {
        nodal_set_error("nodal_csr_build: more than 2^31 triples are not supported");
        return NODAL_BAD_ARG;
    }
    cudaStream_t st = (cudaStream_t)stream;
    u64* keys = reinterpret_cast<u64*>(keys_);
    g_pending = PendingCsr();
    *nnz_h = 0;
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(cudaMemsetAsync(rhs, 0, sizeof(double) * (size_t)n, st));
    if (nslots == 0 || n == 0) {
        g_pending.ctx = ctx;
        g_pending.n = n;
        NODAL_TRY(ctx_reserve(ctx, 4096));
        g_pending.generation = ctx->generation;
        return NODAL_OK;
    }
    int rowbits = 1;
    while (((int64_t)n >> rowbits) != 0) ++rowbits;  // rows go up to n (invalid marker)
    const int bits = colbits + rowbits;

    const size_t slots = (size_t)nslots;
    size_t need = 2 * align_up(slots * 8, 256) + 3 * align_up(slots * 4, 256) +
                  radix_sort_scratch_bytes(nslots) + 2 * scan_scratch_bytes(nslots) + (1 << 16);
    if (order == 1) need += 2 * align_up(slots * 8, 256) + align_up(slots * 4, 256);
    NODAL_TRY(ctx_reserve(ctx, need));
    u64* keys_alt = carve<u64>(ctx, slots);
    u64* vals_alt = carve<u64>(ctx, slots);
    // first-touch order: the sort carries emission indices, the values stay where they are
    u64* idx = order == 1 ? carve<u64>(ctx, slots) : nullptr;
    double* usum = order == 1 ? carve<double>(ctx, slots) : nullptr;
    u32* ufirst = order == 1 ? carve<u32>(ctx, slots) : nullptr;
    if (order == 1 && (!idx || !usum || !ufirst)) return NODAL_CUDA_ERROR;
    u32* head = carve<u32>(ctx, slots);      // head flags -> scan; later reused as krow
    u32* keep = carve<u32>(ctx, slots);      // keep flags per unique key
    u32* keep_scan = carve<u32>(ctx, slots);
    u32* totals = carve<u32>(ctx, 64);
    if (!keys_alt || !vals_alt || !head || !keep || !keep_scan || !totals) return NODAL_CUDA_ERROR;

    bool in_alt = false;
    const size_t mark = ctx->arena_used;
    const int grid = grid_for(ctx, nslots, CB_THREADS);
    if (order == 1) {
        iota_u64_kernel<<<grid, CB_THREADS, 0, st>>>(nslots, idx);
        KERNEL_CHECK();
    }
    NODAL_TRY(radix_sort_pairs(ctx, keys, order == 1 ? idx : reinterpret_cast<u64*>(vals), keys_alt, vals_alt, nslots,
                               bits, &in_alt, st));
    ctx->arena_used = mark;
    const u64* sk = in_alt ? keys_alt : keys;
    const double* sv = in_alt ? reinterpret_cast<double*>(vals_alt) : vals;
    const u64* sidx = in_alt ? vals_alt : idx;
    u64* ukey = in_alt ? keys : keys_alt;  // the other pair is free now
    double* uval = order == 1 ? usum : (in_alt ? vals : reinterpret_cast<double*>(vals_alt));

    mark_heads_kernel<<<grid, CB_THREADS, 0, st>>>(sk, nslots, n, colbits, head);
    KERNEL_CHECK();
    NODAL_TRY(scan_exclusive_u32(ctx, head, head, nslots, totals + 0, st));
    ctx->arena_used = mark;
    if (order == 1)
        segment_sum_first_touch_kernel<<<grid, CB_THREADS, 0, st>>>(sk, sidx, vals, nslots, n, colbits, head, ukey, uval,
                                                                    keep, ufirst, rhs);
    else
        segment_sum_kernel<<<grid, CB_THREADS, 0, st>>>(sk, sv, nslots, n, colbits, head, ukey, uval, keep, rhs);
    KERNEL_CHECK();
    u32* host_tot = reinterpret_cast<u32*>(ctx->pinned);
    CUDA_TRY(cudaMemcpyAsync(host_tot, totals, sizeof(u32), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    NODAL_TRY(radix_sort_check(ctx));
    const int64_t useg = host_tot[0];
    NODAL_TRY(scan_exclusive_u32(ctx, keep, keep_scan, useg, totals + 1, st));
    CUDA_TRY(cudaMemcpyAsync(host_tot + 1, totals + 1, sizeof(u32), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    const int64_t nnz = host_tot[1];

    g_pending.ctx = ctx;
    g_pending.ukey = ukey;
    g_pending.uval = uval;
    g_pending.keep = keep;
    g_pending.keep_scan = keep_scan;
    g_pending.krow = reinterpret_cast<int32_t*>(head);
    g_pending.ufirst = ufirst;
    g_pending.order = order;
    g_pending.useg = useg;
    g_pending.nnz = nnz;
    g_pending.n = n;
    g_pending.colbits = colbits;
    g_pending.generation = ctx->generation;
    *nnz_h = nnz;
    return NODAL_OK;
}

extern "C" int nodal_csr_fetch(nodal_ctx* ctx, int32_t n, int64_t nnz, int32_t* indptr,
                               int32_t* indices, double* data, void* stream) {
    PendingCsr& p = g_pending;
    if (!ctx || p.ctx != ctx || p.n != n || p.nnz != nnz ||
        ctx->generation != p.generation) {
        nodal_set_error("nodal_csr_fetch: no matching nodal_csr_build result is pending on this ctx");
        return NODAL_BAD_ARG;
    }
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(cudaSetDevice(ctx->device));
    if (nnz == 0) {
        CUDA_TRY(cudaMemsetAsync(indptr, 0, sizeof(int32_t) * ((size_t)n + 1), st));
        g_pending = PendingCsr();
        return NODAL_OK;
    }
    compact_kernel<<<grid_for(ctx, p.useg, CB_THREADS), CB_THREADS, 0, st>>>(
        p.ukey, p.uval, p.keep_scan, p.keep, p.useg, p.colbits, indices, data, p.krow);
    KERNEL_CHECK();
    RowGap* gaps = reinterpret_cast<RowGap*>(ctx_pool_alloc(ctx, sizeof(RowGap) * ROWPTR_MAX_GAPS + 256));
    if (!gaps) return NODAL_CUDA_ERROR;
    unsigned int* ngaps = reinterpret_cast<unsigned int*>(gaps + ROWPTR_MAX_GAPS);
    CUDA_TRY(cudaMemsetAsync(ngaps, 0, sizeof(unsigned int), st));
    row_ptr_kernel<<<grid_for(ctx, nnz, CB_THREADS), CB_THREADS, 0, st>>>(p.krow, nnz, n, indptr, gaps, ngaps);
    KERNEL_CHECK();
    row_ptr_gaps_kernel<<<ctx->num_sms * 4, CB_THREADS, 0, st>>>(indptr, gaps, ngaps);
    KERNEL_CHECK();
    ctx_pool_free(ctx, gaps);   // stream-ordered reuse only
    if (p.order == 1 && nnz > 1) {
        // entries of every row by first-touch emission: stable sort of (row, emission) over the kept
        // entries, then gather (pool buffers: the arena still holds the pending result)
        int rowbits = 1;
        while (((int64_t)n >> rowbits) != 0) ++rowbits;
        const size_t cnt = (size_t)nnz;
        u64* k2 = static_cast<u64*>(ctx_pool_alloc(ctx, cnt * 8));
        u64* p2 = static_cast<u64*>(ctx_pool_alloc(ctx, cnt * 8));
        u64* k2a = static_cast<u64*>(ctx_pool_alloc(ctx, cnt * 8));
        u64* p2a = static_cast<u64*>(ctx_pool_alloc(ctx, cnt * 8));
        int32_t* ctmp = static_cast<int32_t*>(ctx_pool_alloc(ctx, cnt * 4));
        double* vtmp = static_cast<double*>(ctx_pool_alloc(ctx, cnt * 8));
        char* sort_ws = static_cast<char*>(ctx_pool_alloc(ctx, radix_sort_scratch_bytes(nnz) + 4096));
        int rc = (k2 && p2 && k2a && p2a && ctmp && vtmp && sort_ws) ? NODAL_OK : NODAL_CUDA_ERROR;
        if (rc == NODAL_OK) {
            first_touch_keys_kernel<<<grid_for(ctx, p.useg, CB_THREADS), CB_THREADS, 0, st>>>(
                p.ukey, p.ufirst, p.keep_scan, p.keep, p.useg, p.colbits, k2, p2);
            ++g_nodal_launches;
            // the sort carves its scratch from the arena: lend it a private one for this call
            char* arena = ctx->arena; const size_t bytes = ctx->arena_bytes, used = ctx->arena_used;
            ctx->arena = sort_ws; ctx->arena_bytes = radix_sort_scratch_bytes(nnz) + 4096; ctx->arena_used = 0;
            bool in_alt = false;
            rc = radix_sort_pairs(ctx, k2, p2, k2a, p2a, nnz, 32 + rowbits, &in_alt, st);
            ctx->arena = arena; ctx->arena_bytes = bytes; ctx->arena_used = used;
            if (rc == NODAL_OK) {
                cudaMemcpyAsync(ctmp, indices, cnt * 4, cudaMemcpyDeviceToDevice, st);
                cudaMemcpyAsync(vtmp, data, cnt * 8, cudaMemcpyDeviceToDevice, st);
                permute_entries_kernel<<<grid_for(ctx, nnz, CB_THREADS), CB_THREADS, 0, st>>>(
                    nnz, in_alt ? p2a : p2, ctmp, vtmp, indices, data);
                ++g_nodal_launches;
                if (cudaGetLastError() != cudaSuccess) rc = NODAL_CUDA_ERROR;
            }
        }
        for (void* q : {(void*)k2, (void*)p2, (void*)k2a, (void*)p2a, (void*)ctmp, (void*)vtmp, (void*)sort_ws})
            ctx_pool_free(ctx, q);
        if (rc != NODAL_OK) { g_pending = PendingCsr(); return rc; }
    }
    g_pending = PendingCsr();
    return NODAL_OK;
}

// ---------------------------------------------------------------- dense conversions
__global__ void __launch_bounds__(CB_THREADS)
csr_to_dense_kernel(int32_t n, const int32_t* __restrict__ indptr,
                    const int32_t* __restrict__ indices, const double* __restrict__ data,
                    double* __restrict__ G) {
    // one warp per row
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < n; r += nwarps) {
        const int32_t e = indptr[r + 1];
        for (int32_t j = indptr[r] + lane; j < e; j += 32)
            G[(size_t)r * n + indices[j]] = data[j];
    }
}

extern "C" int nodal_csr_to_dense(nodal_ctx* ctx, int32_t n, const int32_t* indptr,
                                  const int32_t* indices, const double* data, double* G,
                                  void* stream) {
    if (!ctx || n < 0) return NODAL_BAD_ARG;
    if (n == 0) return NODAL_OK;
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(cudaMemsetAsync(G, 0, sizeof(double) * (size_t)n * n, st));
    csr_to_dense_kernel<<<grid_for(ctx, (int64_t)n * 32, CB_THREADS), CB_THREADS, 0, st>>>(
        n, indptr, indices, data, G);
    KERNEL_CHECK();
    return NODAL_OK;
}

// Warp-aggregated atomic scatter-add: lanes of a warp that hit the same matrix entry are
// combined (lane order) and issue one atomicAdd.
__global__ void __launch_bounds__(CB_THREADS)
coo_to_dense_atomic_kernel(int32_t n, int64_t nslots, int colbits, const u64* __restrict__ keys,
                           const double* __restrict__ vals, double* __restrict__ G,
                           double* __restrict__ rhs) {
    const u64 colmask = ((u64)1 << colbits) - 1;
    const int lane = threadIdx.x & 31;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t rounds = (nslots + stride - 1) / stride;
    for (int64_t it = 0; it < rounds; ++it) {  // warp-uniform trip count
        const int64_t i = it * stride + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        u64 k = ~0ull;
        double v = 0.0;
        bool live = false;
        if (i < nslots) {
            k = keys[i];
            v = vals[i];
            live = (int64_t)(k >> colbits) < (int64_t)n;
        }
        if (!live) k = ~0ull - lane;  // unique per lane: never aggregated
        const u32 peers = __match_any_sync(0xffffffffu, k);
        const bool dup = __any_sync(0xffffffffu, live && peers != (1u << lane));
        double s = v;
        if (dup) {
            s = 0.0;
#pragma unroll
            for (int l = 0; l < 32; ++l) {
                const double o = __shfl_sync(0xffffffffu, v, l);
                if ((peers >> l) & 1u) s += o;
            }
        }
        const bool leader = (peers & ((1u << lane) - 1)) == 0u;
        if (live && leader) {
            const int32_t row = (int32_t)(k >> colbits), col = (int32_t)(k & colmask);
            if (col == n) atomicAdd(&rhs[row], s);
            else atomicAdd(&G[(size_t)row * n + col], s);
        }
    }
}

extern "C" int nodal_coo_to_dense_atomic(nodal_ctx* ctx, int32_t n, int64_t nslots, int32_t colbits,
                                         const uint64_t* keys, const double* vals, double* G,
                                         double* rhs, void* stream) {
    if (!ctx || n < 0 || nslots < 0) return NODAL_BAD_ARG;
    if (n == 0) return NODAL_OK;
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(cudaMemsetAsync(G, 0, sizeof(double) * (size_t)n * n, st));
    CUDA_TRY(cudaMemsetAsync(rhs, 0, sizeof(double) * (size_t)n, st));
    if (nslots == 0) return NODAL_OK;
    coo_to_dense_atomic_kernel<<<grid_for(ctx, nslots, CB_THREADS), CB_THREADS, 0, st>>>(
        n, nslots, colbits, reinterpret_cast<const u64*>(keys), vals, G, rhs);
    KERNEL_CHECK();
    return NODAL_OK;
}
