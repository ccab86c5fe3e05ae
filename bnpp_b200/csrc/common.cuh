// Shared device helpers and the context object of the bn-pp B200 factor-algebra library.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

#include "../../include/bnpp_b200.h"

namespace bnpp {

constexpr int kMaxK = BNPP_MAX_OPERANDS;
constexpr int kMaxR = BNPP_MAX_AXES;
constexpr int kBlock = 256;
constexpr int kMaxPartials = 4096;

// Exact unsigned 32-bit division by an invariant divisor d >= 2 (Granlund-Montgomery,
// round-up form): q = (t + ((n - t) >> 1)) >> sh with t = umulhi(n, m).  Powers of two
// get m = 1 => t = 0 and the expression degenerates to a shift, so one code path
// serves the (dominant) binary-variable case and mixed cardinalities alike.
struct FastDiv {
    uint32_t d, m, sh;
};

inline FastDiv make_fastdiv(uint32_t d)
{
    FastDiv f;
    f.d = d;
    uint32_t l = 0;
    while ((1ull << l) < d) ++l;            // l = ceil(log2 d), d >= 2 => l >= 1
    f.m = (uint32_t)(((1ull << 32) * ((1ull << l) - d)) / d + 1);
    f.sh = l - 1;
    return f;
}

__device__ __forceinline__ uint32_t fastdiv(uint32_t n, const FastDiv &f)
{
    uint32_t t = __umulhi(n, f.m);
    return (t + ((n - t) >> 1)) >> f.sh;
}

struct __align__(32) double4_t {
    double x, y, z, w;
};

// read-only path loads; the 256-bit form is sm_100+ (LDG.E.256)
__device__ __forceinline__ double ld1(const double *p) { return __ldg(p); }
__device__ __forceinline__ double2 ld2(const double *p) { return __ldg(reinterpret_cast<const double2 *>(p)); }
__device__ __forceinline__ double4_t ld4(const double *p)
{
    double4_t r;
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w)
                 : "l"(p));
    return r;
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

// Deterministic grid-wide sum of one double per thread: block tree -> partials[block]
// -> the last block to arrive (ticket) adds the partials in a fixed order and writes
// *z.  The summation order depends only on (gridDim, blockDim), never on timing.
__device__ __forceinline__ void grid_sum_to(double v, double *partials, unsigned int *ticket, double *z)
{
    __shared__ double s_w[kBlock / 32];
    __shared__ bool s_last;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = warp_sum(v);
    if (lane == 0) s_w[w] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double b = 0;
#pragma unroll
        for (int i = 0; i < kBlock / 32; ++i) b += s_w[i];
        partials[blockIdx.x] = b;
        __threadfence();
        unsigned int t = atomicAdd(ticket, 1u);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        double a = 0;
        for (unsigned i = threadIdx.x; i < gridDim.x; i += blockDim.x) a += __ldcg(partials + i);
        a = warp_sum(a);
        __syncthreads();
        if (lane == 0) s_w[w] = a;
        __syncthreads();
        if (threadIdx.x == 0) {
            double b = 0;
#pragma unroll
            for (int i = 0; i < kBlock / 32; ++i) b += s_w[i];
            if (z) *z = b;
            *ticket = 0;
        }
    }
}

}  // namespace bnpp

// The opaque context handed out by the C ABI.
struct bnpp_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaMemPool_t pool = nullptr;
    int sm_count = 148;
    double *partials = nullptr;        // kMaxPartials doubles
    unsigned int *ticket = nullptr;    // last-block ticket, zero between launches
    unsigned int *status = nullptr;    // BNPP_STATUS_* bits
    uint64_t launches = 0;
    std::string last_error;
    std::string last_kernel;
    void *last_desc = nullptr;         // bnpp::LaunchDesc of the most recent contraction (name formatted on demand)
    // pinned staging ring for the small tables a plan uploads while it resolves its launches (offset tables of the
    // multi-valued kernels, task programs): a copy from pageable memory would synchronise the stream every time
    unsigned char *stage = nullptr;
    size_t stage_bytes = 0, stage_off = 0;
    uint32_t last_grid = 0, last_block = 0;
};

namespace bnpp {
// host -> device copy of a small table, stream-ordered and asynchronous (through the context's pinned ring)
int stage_upload(bnpp_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes);
void *stage_reserve(bnpp_ctx *ctx, size_t bytes);
int fail(bnpp_ctx *ctx, int code, const std::string &msg);
int cuda_fail(bnpp_ctx *ctx, cudaError_t e, const char *what);
#define BNPP_CUDA(ctx, expr)                                                   \
    do {                                                                       \
        cudaError_t _e = (expr);                                               \
        if (_e != cudaSuccess) return ::bnpp::cuda_fail((ctx), _e, #expr);     \
    } while (0)
}  // namespace bnpp
