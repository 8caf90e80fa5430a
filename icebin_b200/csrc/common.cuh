// common.cuh -- error handling, stream-ordered device buffers, launch accounting.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/icebin_b200.h"

namespace ib {

struct Error : std::runtime_error {
    int code;
    Error(int c, std::string const &m) : std::runtime_error(m), code(c) {}
};

[[noreturn]] inline void fail(int code, const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    throw Error(code, buf);
}

#define IB_CUDA(expr)                                                                                   \
    do {                                                                                                \
        cudaError_t e_ = (expr);                                                                        \
        if (e_ != cudaSuccess)                                                                          \
            ::ib::fail(IB_ERR_CUDA, "CUDA error %s at %s:%d: %s", cudaGetErrorName(e_), __FILE__, __LINE__, \
                       cudaGetErrorString(e_));                                                         \
    } while (0)

extern std::atomic<int64_t> g_launches;   // kernels launched by this library (ib_launch_count)
#define IB_LAUNCHED()                 \
    do {                              \
        ::ib::g_launches.fetch_add(1, std::memory_order_relaxed); \
        IB_CUDA(cudaGetLastError());  \
    } while (0)

cudaStream_t default_stream();   // the library's own non-blocking stream (created on first use, per device)
int sm_count();
int device_index();              // current device, < IB_MAX_DEVICES
constexpr int IB_MAX_DEVICES = 64;
// performance knobs (ib_tune): never change results
struct Tune {
    int stream_bps = 3;          // CTAs per SM of spmm_stream_kernel
    int stream_spare = 24;       // CTA slots spmm_stream_kernel leaves free: a grid that fills EVERY slot (444 CTAs) runs 62 us, one that
                                 // leaves 8 / 24 / 48 free 59.0 / 58.0 / 57.1 us (tools/stream_sweep.py --spare); also the room a concurrent
                                 // small kernel needs (the multi-GPU exchange)
    int tma_bps = 4;             // CTAs per SM of spmm_tma_kernel
    int gc_nq = 0;               // spmm_gc_kernel: 64-field groups per lane (0: by nvar)
    int use_stream = 1;          // one-pass kernel for wide matrices
    int use_gc = 1;              // distinct-column group kernel for short-row matrices
    int use_group = 1;           // tensor-core group kernel for tall matrices with ~16 nonzeros per row that share columns
    int xchg_blocks = 96;        // 128-thread blocks of the two exchange kernels (<= the CTA slots a concurrent SpMM leaves free, or they wait for its CTAs to exit)
    int hntr_generic = 0;        // Hntr: always the generic kernels (skip the int16 / constant-weight fast path)
};
extern Tune g_tune;

inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Stream-ordered device allocation (cudaMallocAsync; the default pool keeps freed blocks so steady-state
// builds do not hit the driver).  Move-only.
template <class T>
struct DBuf {
    T *p = nullptr;
    size_t n = 0;
    cudaStream_t st = nullptr;
    DBuf() {}
    DBuf(size_t n_, cudaStream_t s) { alloc(n_, s); }
    DBuf(DBuf const &) = delete;
    DBuf &operator=(DBuf const &) = delete;
    DBuf(DBuf &&o) noexcept : p(o.p), n(o.n), st(o.st) { o.p = nullptr; o.n = 0; }
    DBuf &operator=(DBuf &&o) noexcept
    {
        if (this != &o) { release(); p = o.p; n = o.n; st = o.st; o.p = nullptr; o.n = 0; }
        return *this;
    }
    ~DBuf() { release(); }
    void alloc(size_t n_, cudaStream_t s)
    {
        release();
        n = n_; st = s;
        if (n) IB_CUDA(cudaMallocAsync((void **)&p, n * sizeof(T) , s));
    }
    void release()
    {
        if (p) cudaFreeAsync(p, st);
        p = nullptr; n = 0;
    }
    void zero() { if (n) IB_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), st)); }
    void fill_byte(int b) { if (n) IB_CUDA(cudaMemsetAsync(p, b, n * sizeof(T), st)); }
    void upload(const T *h, size_t cnt) { if (cnt) IB_CUDA(cudaMemcpyAsync(p, h, cnt * sizeof(T), cudaMemcpyHostToDevice, st)); }
    void download(T *h, size_t cnt) const { if (cnt) IB_CUDA(cudaMemcpyAsync(h, p, cnt * sizeof(T), cudaMemcpyDeviceToHost, st)); }
    std::vector<T> to_host() const
    {
        std::vector<T> v(n);
        download(v.data(), n);
        IB_CUDA(cudaStreamSynchronize(st));
        return v;
    }
};

void set_last_error(std::string const &msg);   // what ib_last_error() returns on this thread
void ensure_device();   // throws IB_ERR_CUDA when no CUDA device is present (no CPU fallback)

}   // namespace ib
