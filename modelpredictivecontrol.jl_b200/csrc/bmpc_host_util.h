// Host-side helpers shared by the translation units of libbmpc.so.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/bmpc.h"

namespace bmpc_host {

std::string& last_error();  // thread-local message returned by bmpc_last_error()

inline int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    last_error() = buf;
    return code;
}

#define CK(call)                                                                                           \
    do {                                                                                                   \
        cudaError_t e_ = (call);                                                                           \
        if (e_ != cudaSuccess)                                                                             \
            return bmpc_host::fail(BMPC_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_),  \
                                   __FILE__, __LINE__);                                                    \
    } while (0)

inline int even(int v) { return (v + 1) & ~1; }

// The dynamic shared-memory limit of a kernel is a per-FUNCTION attribute shared by every handle of the process:
// only ever raise it (a handle with a smaller footprint must not shrink the limit under another handle's launch).
inline cudaError_t raise_dyn_smem(const void* func, int bytes) {
    cudaFuncAttributes at{};
    cudaError_t e = cudaFuncGetAttributes(&at, func);
    if (e != cudaSuccess) return e;
    if (bytes <= at.maxDynamicSharedSizeBytes) return cudaSuccess;
    return cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}

template <class T>
struct DevBuf {  // owning device allocation: freed by the destructor (handles are destroyed with the device current)
    T* p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    cudaError_t alloc(size_t count) {
        if (count <= n && p) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
        if (count == 0) return cudaSuccess;
        cudaError_t e = cudaMalloc(&p, count * sizeof(T));
        if (e == cudaSuccess) n = count;
        return e;
    }
    cudaError_t upload(const T* h, size_t count, cudaStream_t s) {
        cudaError_t e = alloc(count);
        if (e != cudaSuccess || count == 0) return e;
        return cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, s);
    }
    cudaError_t upload(const std::vector<T>& v, cudaStream_t s) { return upload(v.data(), v.size(), s); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
};

}  // namespace bmpc_host
