// Context management of the C ABI (include/haslr_b200.h).
#include <cstdlib>
#include <string>

#include "common.cuh"

extern "C" int hgpu_abi_version(void) { return 4; }   // 3: device-resident stages (hgpu_hits_group, *_dev), hgpu_stage_stats / hgpu_set_timing; 4: hgpu_host_staging

extern "C" const char* hgpu_strerror(int code) {
    switch (code) {
        case HGPU_OK: return "ok";
        case HGPU_E_INVALID: return "invalid argument";
        case HGPU_E_CUDA: return "CUDA runtime error";
        case HGPU_E_NOMEM: return "out of memory";
        case HGPU_E_NOSPACE: return "output capacity too small";
        case HGPU_E_UNSUPPORTED: return "unsupported request";
        case HGPU_E_INTERNAL: return "internal inconsistency";
        default: return "unknown error";
    }
}

extern "C" int hgpu_create(int device, hgpu_t** out) {
    if (!out) return HGPU_E_INVALID;
    *out = nullptr;
    // the POA scheduler runs up to 6 size classes + helpers side by side: give them hardware work queues of their own (only
    // effective if the process has not initialised CUDA yet; poa.cu stays within the default 8 otherwise)
    setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) return HGPU_E_CUDA;   // no silent CPU path: the product needs a GPU
    if (device < 0) { if (cudaGetDevice(&device) != cudaSuccess) return HGPU_E_CUDA; }
    if (device >= n) return HGPU_E_INVALID;
    if (cudaSetDevice(device) != cudaSuccess) return HGPU_E_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return HGPU_E_CUDA;
    if (prop.major < 10) return HGPU_E_UNSUPPORTED;       // built for sm_100a only
    hgpu_ctx* c = new hgpu_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->smem_optin = prop.sharedMemPerBlockOptin;
    *out = c;
    return HGPU_OK;
}

extern "C" void hgpu_destroy(hgpu_t* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    poa_state_destroy(ctx->poa);
    k12_state_destroy(ctx->k12);
    coord_state_destroy(ctx->coords);
    paf_state_destroy(ctx->paf);
    for (void* p : ctx->staging) if (p) cudaFreeHost(p);
    delete ctx;
}

extern "C" int hgpu_set_stream(hgpu_t* ctx, void* cuda_stream) {
    if (!ctx) return HGPU_E_INVALID;
    ctx->stream = (cudaStream_t)cuda_stream;
    return HGPU_OK;
}

extern "C" int hgpu_host_staging(hgpu_t* ctx, uint32_t which, uint64_t bytes, void** out) {
    if (!ctx || !out || which > 1) return HGPU_E_INVALID;
    *out = nullptr;
    if (ctx->staging_bytes[which] < bytes || !ctx->staging[which]) {
        HGPU_CUDA(ctx, cudaSetDevice(ctx->device));
        if (ctx->staging[which]) { HGPU_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); cudaFreeHost(ctx->staging[which]); ctx->staging[which] = nullptr; ctx->staging_bytes[which] = 0; }
        const uint64_t want = bytes + bytes / 8 + 4096;      // head room: the next batch of a similar size reuses the buffer
        void* p = nullptr;
        if (cudaHostAlloc(&p, want, cudaHostAllocDefault) != cudaSuccess) {
            cudaGetLastError();                               // not sticky, but the next launch check must not find it: callers fall back to pageable memory
            ctx->last_error = "hgpu_host_staging: cannot page-lock " + std::to_string(want) + " bytes";
            return HGPU_E_NOMEM;
        }
        ctx->staging[which] = p; ctx->staging_bytes[which] = want;
    }
    *out = ctx->staging[which];
    return HGPU_OK;
}

extern "C" const char* hgpu_last_error(const hgpu_t* ctx) { return ctx ? ctx->last_error.c_str() : "null context"; }

extern "C" uint64_t hgpu_launch_count(const hgpu_t* ctx) { return ctx ? ctx->launches : 0; }

extern "C" int hgpu_get_stage_stats(const hgpu_t* ctx, hgpu_stage_stats* out) {
    if (!ctx || !out) return HGPU_E_INVALID;
    *out = ctx->stage;
    return HGPU_OK;
}

extern "C" int hgpu_set_timing(hgpu_t* ctx, int enabled) {
    if (!ctx) return HGPU_E_INVALID;
    ctx->timing = enabled != 0;
    return hgpu_poa_set_timing(ctx, enabled);
}
