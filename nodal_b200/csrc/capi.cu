// Context, error reporting and scratch arena of libnodal_b200.so.
#include <stdarg.h>

#include "common.cuh"

static thread_local char g_err[1024] = "";
unsigned long long g_nodal_launches = 0;

void nodal_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" int nodal_abi_version(void) { return NODAL_ABI_VERSION; }
extern "C" const char* nodal_last_error(void) { return g_err; }
extern "C" uint64_t nodal_launch_count(void) { return g_nodal_launches; }

extern "C" int nodal_ctx_create(int device, nodal_ctx** out) {
    if (!out) return NODAL_BAD_ARG;
    int count = 0;
    CUDA_TRY(cudaGetDeviceCount(&count));
    if (device < 0 || device >= count) {
        nodal_set_error("nodal_ctx_create: device %d out of range (%d visible)", device, count);
        return NODAL_BAD_ARG;
    }
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        nodal_set_error("nodal_b200 needs an sm_100a device, found sm_%d%d", prop.major, prop.minor);
        return NODAL_CUDA_ERROR;
    }
    nodal_ctx* ctx = new nodal_ctx();
    ctx->device = device;
    ctx->num_sms = prop.multiProcessorCount;
    CUDA_TRY(cudaMallocHost(&ctx->pinned, 4096));
    memset(ctx->pinned, 0, 4096);
    *out = ctx;
    return NODAL_OK;
}

extern "C" int nodal_ctx_destroy(nodal_ctx* ctx) {
    if (!ctx) return NODAL_OK;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    if (ctx->arena) cudaFree(ctx->arena);
    for (auto& b : ctx->pool)
        if (b.ptr) cudaFree(b.ptr);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    delete ctx;
    return NODAL_OK;
}

extern "C" int64_t nodal_ctx_workspace_bytes(nodal_ctx* ctx) {
    if (!ctx) return 0;
    size_t total = ctx->arena_bytes;
    for (auto& b : ctx->pool) total += b.bytes;
    return (int64_t)total;
}

void* ctx_pool_alloc(nodal_ctx* ctx, size_t bytes) {
    bytes = align_up(bytes < 256 ? 256 : bytes, 256);
    int best = -1;
    for (size_t i = 0; i < ctx->pool.size(); ++i) {
        auto& b = ctx->pool[i];
        if (!b.used && b.bytes >= bytes && b.bytes <= 2 * bytes + (1 << 20) &&
            (best < 0 || b.bytes < ctx->pool[best].bytes))
            best = (int)i;
    }
    if (best >= 0) {
        ctx->pool[best].used = true;
        return ctx->pool[best].ptr;
    }
    void* p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) {
        (void)cudaGetLastError();
        // drop the cache and retry once
        for (auto& b : ctx->pool)
            if (!b.used && b.ptr) { cudaFree(b.ptr); b.ptr = nullptr; b.bytes = 0; }
        if (cudaMalloc(&p, bytes) != cudaSuccess) {
            (void)cudaGetLastError();
            return nullptr;
        }
    }
    ctx->pool.push_back({p, bytes, true});
    return p;
}

void ctx_pool_free(nodal_ctx* ctx, void* ptr) {
    if (!ptr) return;
    for (auto& b : ctx->pool)
        if (b.ptr == ptr) { b.used = false; return; }
    cudaFree(ptr);
}

int ctx_reserve(nodal_ctx* ctx, size_t bytes) {
    CUDA_TRY(cudaSetDevice(ctx->device));
    bytes = align_up(bytes + 4096, 1 << 20);
    if (bytes > ctx->arena_bytes) {
        CUDA_TRY(cudaDeviceSynchronize());
        if (ctx->arena) CUDA_TRY(cudaFree(ctx->arena));
        ctx->arena = nullptr;
        ctx->arena_bytes = 0;
        CUDA_TRY(cudaMalloc(&ctx->arena, bytes));
        ctx->arena_bytes = bytes;
    }
    ctx->arena_used = 0;
    ctx->generation++;
    return NODAL_OK;
}

void* ctx_carve(nodal_ctx* ctx, size_t bytes) {
    size_t off = align_up(ctx->arena_used, 256);
    if (off + bytes > ctx->arena_bytes) {
        nodal_set_error("internal: scratch arena overflow (%zu + %zu > %zu)", off, bytes,
                        ctx->arena_bytes);
        return nullptr;
    }
    ctx->arena_used = off + bytes;
    return ctx->arena + off;
}
