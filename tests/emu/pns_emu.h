// TEST BUILD ONLY.  Host-side shims that let g++ compile pednstream_b200/csrc/pns_kernels.cu as
// ordinary C++ (-DPNS_HOST_EMULATION): every "thread" of a launch runs sequentially.  The
// resulting library (tests/_emu/libpns_emu.so, built by tests/emu/build_emu.py) exists so the
// kernel logic can be unit-tested against the oracle on machines without a GPU; the product
// package never loads it (pednstream_b200/_native.py only opens lib/libpns_b200.so).
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __grid_constant__
#define __noinline__
struct int4 { int x, y, z, w; };
struct double2 { double x, y; };
struct float2 { float x, y; };

struct emu_dim3 { unsigned x, y, z; };
static thread_local emu_dim3 blockIdx, threadIdx, blockDim, gridDim;

typedef void* cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0 };
static inline const char* cudaGetErrorString(cudaError_t) { return "emulation"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }

template <class T> static inline T __ldg(const T* p) { return *p; }
template <class T> static inline T __ldcg(const T* p) { return *p; }
static inline int atomicOr(int32_t* p, int v) { int o = *p; *p = o | v; return o; }
static inline int __float2int_rn(float x) { return (int)lrintf(x); }   // default rounding: nearest even
using std::max;
using std::min;

#define PNS_LAUNCH(kern, nblk, nthr, stream, ...)                                   \
    do {                                                                            \
        gridDim.x = (unsigned)(nblk); blockDim.x = (unsigned)(nthr);                \
        for (unsigned _b = 0; _b < (unsigned)(nblk); ++_b)                          \
            for (unsigned _t = 0; _t < (unsigned)(nthr); ++_t) {                    \
                blockIdx.x = _b; threadIdx.x = _t;                                  \
                kern(__VA_ARGS__);                                                  \
            }                                                                       \
    } while (0)
