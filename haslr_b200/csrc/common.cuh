// Context object and small host-side helpers shared by the three kernel families.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/haslr_b200.h"

struct PoaState;  // poa.cu
struct K12State;  // k12.cu
struct CoordState;  // coords.cu
struct PafState;  // paf.cu

// Tables a context keeps resident on the device between the stages of the path (K0 -> K1 -> K2 -> K4), so that nothing
// but results the host really consumes travels back: the hit table (made by hgpu_paf_tokenize or hgpu_hits_upload, grouped
// by read with hgpu_hits_group) and the compact long reads (made by hgpu_compact_lr[_dev]).
struct ResidentHits {
    bool valid = false, grouped = false;
    uint32_t n_hits = 0, n_reads = 0;
    uint64_t n_ops = 0;
    const uint32_t *q_id = nullptr, *q_start = nullptr, *q_end = nullptr, *t_id = nullptr, *t_len = nullptr, *t_start = nullptr, *t_end = nullptr,
                   *n_match = nullptr, *n_block = nullptr, *cg_off = nullptr, *cg_ops = nullptr, *read_off = nullptr;
    const uint8_t *is_rev = nullptr, *mapq = nullptr;
};
struct ResidentCompact {
    bool valid = false;
    uint32_t n_elems = 0, n_reads = 0;
    const hgpu_cl_elem* elems = nullptr;
    const uint32_t* read_off = nullptr;
};
// CUDA-event pair around the kernels of one stage (recorded only when timing is on)
struct StageEvents {
    cudaEvent_t a = nullptr, b = nullptr;
    ~StageEvents() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }
};

struct hgpu_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    int sm_count = 148;
    size_t smem_optin = 0;
    std::string last_error;
    uint64_t launches = 0;
    bool timing = false;
    hgpu_stage_stats stage{};
    StageEvents ev_k0, ev_k1, ev_k2, ev_k4;
    ResidentHits hits;
    ResidentCompact compact;
    PoaState* poa = nullptr;
    K12State* k12 = nullptr;
    CoordState* coords = nullptr;
    PafState* paf = nullptr;
    void* staging[2] = {nullptr, nullptr};     // page-locked host buffers lent to the caller (hgpu_host_staging)
    uint64_t staging_bytes[2] = {0, 0};
};

#define HGPU_CUDA(ctx, expr)                                                                         \
    do {                                                                                             \
        cudaError_t _e = (expr);                                                                     \
        if (_e != cudaSuccess) {                                                                     \
            char _b[512];                                                                            \
            snprintf(_b, sizeof _b, "%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            (ctx)->last_error = _b;                                                                  \
            return _e == cudaErrorMemoryAllocation ? HGPU_E_NOMEM : HGPU_E_CUDA;                     \
        }                                                                                            \
    } while (0)

#define HGPU_FAIL(ctx, code, ...)                          \
    do {                                                   \
        char _b[512];                                      \
        snprintf(_b, sizeof _b, __VA_ARGS__);              \
        (ctx)->last_error = _b;                            \
        return (code);                                     \
    } while (0)

// RAII device buffer (freed with the owning scope; never shared across contexts)
template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
    cudaError_t alloc(size_t count) {
        release();
        if (count == 0) count = 1;
        cudaError_t e = cudaMalloc((void**)&p, count * sizeof(T));
        if (e == cudaSuccess) n = count;
        return e;
    }
    cudaError_t ensure(size_t count) { return count <= n ? cudaSuccess : alloc(count); }
};

// copies that count themselves (hgpu_stage_stats::h2d_bytes / d2h_bytes)
#define HGPU_H2D(ctx, dst, src, bytes) do { const size_t _n = (bytes); if (_n) { HGPU_CUDA(ctx, cudaMemcpyAsync((dst), (src), _n, cudaMemcpyHostToDevice, (ctx)->stream)); (ctx)->stage.h2d_bytes += _n; } } while (0)
#define HGPU_D2H(ctx, dst, src, bytes) do { const size_t _n = (bytes); if (_n) { HGPU_CUDA(ctx, cudaMemcpyAsync((dst), (src), _n, cudaMemcpyDeviceToHost, (ctx)->stream)); (ctx)->stage.d2h_bytes += _n; } } while (0)

static inline void stage_begin(hgpu_ctx* c, StageEvents& e) {
    if (!c->timing) return;
    if (!e.a) { cudaEventCreate(&e.a); cudaEventCreate(&e.b); }
    cudaEventRecord(e.a, c->stream);
}
static inline void stage_end(hgpu_ctx* c, StageEvents& e) { if (c->timing && e.b) cudaEventRecord(e.b, c->stream); }
// after the stream has been synchronised
static inline float stage_ms(hgpu_ctx* c, StageEvents& e) {
    float ms = 0;
    if (c->timing && e.a && e.b && cudaEventSynchronize(e.b) == cudaSuccess) cudaEventElapsedTime(&ms, e.a, e.b);
    return ms;
}

void poa_state_destroy(PoaState* s);
void k12_state_destroy(K12State* s);
void coord_state_destroy(CoordState* s);
void paf_state_destroy(PafState* s);
