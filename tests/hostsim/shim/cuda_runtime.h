/* tests/hostsim/shim/cuda_runtime.h -- TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * Just enough of the CUDA runtime's surface, implemented on the host, for tests/hostsim/ to compile the product's
 * thread-independent slot path (libav_b200/csrc/slots.cu + h264dsp.cuh) as plain C++ with g++: "device memory" is host
 * memory, copies are memcpy, a launch runs the kernel body once per (block, thread) index in sequence.  That is only
 * valid for kernels whose threads do not communicate (no shared memory, shuffles, barriers or atomics) -- exactly the
 * per-call slot kernel; every other kernel of the product is checked on the GPU only.
 */
#ifndef HOSTSIM_CUDA_RUNTIME_H
#define HOSTSIM_CUDA_RUNTIME_H
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __launch_bounds__(...)

typedef int cudaError_t;
enum { cudaSuccess = 0 };
typedef struct hostsim_stream *cudaStream_t;
typedef struct hostsim_event *cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };

struct uint3 { unsigned x, y, z; };
struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
extern uint3 threadIdx, blockIdx;
extern dim3 blockDim, gridDim;

static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline const char *cudaGetErrorString(cudaError_t) { return "hostsim"; }

static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned sel)
{
    const uint64_t v = ((uint64_t)b << 32) | a;
    unsigned r = 0;
    for (int i = 0; i < 4; i++) {
        const unsigned s = (sel >> (4 * i)) & 0xf;
        unsigned byte = (unsigned)(v >> (8 * (s & 7))) & 0xff;
        if (s & 8) byte = (byte & 0x80) ? 0xff : 0;
        r |= byte << (8 * i);
    }
    return r;
}
static inline int atomicMin(int *p, int v) { const int old = *p; if (v < old) *p = v; return old; }      /* threads run one after the other here */
static inline int __mulhi(int a, int b) { return (int)(((int64_t)a * b) >> 32); }
static inline size_t __cvta_generic_to_shared(const void *p) { return (size_t)p; }

namespace hostsim {
template <class... P> struct Launch {
    void (*kernel)(P...);
    dim3 grid, block;
    void operator()(P... args) const
    {
        gridDim = grid; blockDim = block;
        for (unsigned bz = 0; bz < grid.z; bz++) for (unsigned by = 0; by < grid.y; by++) for (unsigned bx = 0; bx < grid.x; bx++)
            for (unsigned tz = 0; tz < block.z; tz++) for (unsigned ty = 0; ty < block.y; ty++) for (unsigned tx = 0; tx < block.x; tx++) {
                blockIdx = { bx, by, bz }; threadIdx = { tx, ty, tz };
                kernel(args...);
            }
    }
};
template <class... P> Launch<P...> launch(void (*k)(P...), dim3 g, dim3 b) { return Launch<P...>{ k, g, b }; }
}
#define AVB_LAUNCH(kernel, grid, block, smem, stream) hostsim::launch(kernel, grid, block)

#endif
