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

struct hgpu_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    int sm_count = 148;
    size_t smem_optin = 0;
    std::string last_error;
    uint64_t launches = 0;
    PoaState* poa = nullptr;
    K12State* k12 = nullptr;
    CoordState* coords = nullptr;
    PafState* paf = nullptr;
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

void poa_state_destroy(PoaState* s);
void k12_state_destroy(K12State* s);
void coord_state_destroy(CoordState* s);
void paf_state_destroy(PafState* s);
