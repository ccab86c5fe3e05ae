// K9 -- variable elimination of a whole plan in one launch, intermediates in shared memory.
//
// Replaces, for plans whose steps are all small, the per-bucket launches of ve.cu / batched.cu
// (reference: the bucket loop of BN::variable_elimination, code/model.cpp:409-439, run once per
// evidence set by BN::partition, code/model.cpp:275-294).  A group of G lanes owns one evidence
// set and walks the step list of fused.hpp:
//   * the arena of a set (its live intermediates, first-fit over a depth-first step order that
//     keeps it small) sits in shared memory, the sets of a warp interleaved entry by entry so
//     that lanes reading consecutive entries of their sets hit consecutive banks;
//   * a resident CPT is read in place through the base offset its observed axes select
//     (Factor::conditioning, code/factor.cpp:214-242, reduced to one multiply-add per axis);
//   * operand offsets per output entry come from the step's table (same numbers for every set);
//   * nothing but the result of the last step is written to HBM.
// Per entry the arithmetic is that of contract.cu / batched.cu -- operands multiplied in bucket
// order with __dmul_rn, values of the eliminated variable added in index order with __dadd_rn
// -- so a result is bit-identical to the launch-per-bucket paths.
#include <cstdlib>
#include <map>

#include "common.cuh"
#include "fused.hpp"

namespace bnpp {

__device__ __forceinline__ uint4 ldg4(const uint32_t *p) { return __ldg(reinterpret_cast<const uint4 *>(p)); }

template <int G>
__device__ __forceinline__ void group_sync()
{
    if (G <= 32) __syncwarp();
    else __syncthreads();
}

// the output entries lane, lane + G, ... of one step; returns their sum
template <int K, int CX, int G>
__device__ __forceinline__ double fused_entries(const double *(&src)[K], const uint32_t (&mul)[K],
                                                const uint32_t (&sxm)[K], const uint32_t *__restrict__ tab, uint32_t n_out,
                                                uint32_t cx, uint32_t lane, double *dst, uint64_t dmul, bool store)
{
    double zacc = 0.0;
    for (uint32_t o = lane; o < n_out; o += G) {
        const double *ptr[K];
#pragma unroll
        for (int q = 0; q < K; ++q) ptr[q] = src[q] + (size_t)(__ldg(tab + (size_t)q * n_out + o) * mul[q]);
        double acc;
        if (CX == 1) {
            double t[K];
#pragma unroll
            for (int q = 0; q < K; ++q) t[q] = *ptr[q];
            acc = t[0];
#pragma unroll
            for (int q = 1; q < K; ++q) acc = __dmul_rn(acc, t[q]);
        } else if (CX == 2) {
            double t0[K], t1[K];
#pragma unroll
            for (int q = 0; q < K; ++q) {
                t0[q] = *ptr[q];
                t1[q] = *(ptr[q] + sxm[q]);
            }
            double v0 = t0[0], v1 = t1[0];
#pragma unroll
            for (int q = 1; q < K; ++q) {
                v0 = __dmul_rn(v0, t0[q]);
                v1 = __dmul_rn(v1, t1[q]);
            }
            acc = __dadd_rn(v0, v1);
        } else {
            acc = 0.0;
            for (uint32_t x = 0; x < cx; ++x) {
                double v = *(ptr[0] + (size_t)x * sxm[0]);
#pragma unroll
                for (int q = 1; q < K; ++q) v = __dmul_rn(v, *(ptr[q] + (size_t)x * sxm[q]));
                acc = (x == 0) ? v : __dadd_rn(acc, v);
            }
        }
        if (store) dst[(uint64_t)o * dmul] = acc;
        zacc = __dadd_rn(zacc, acc);
    }
    return zacc;
}

// operand records of one step, then its entries; returns the program counter after the step
template <int K, int G>
__device__ __forceinline__ uint32_t fused_step(const FusedLaunch &p, uint32_t pc, double *abase, uint32_t amul,
                                               const uint8_t *ev, uint32_t n_out, uint32_t cx, uint32_t tab_off,
                                               uint32_t lane, double *dst, uint64_t dmul, bool store, double &zacc)
{
    const double *src[K];
    uint32_t mul[K], sxm[K];
#pragma unroll
    for (int q = 0; q < K; ++q) {
        const uint4 r = ldg4(p.prog + pc);
        pc += kFusedOperandWords;
        const uint32_t kind = r.x & 0xffu, nobs = r.x >> 8;
        if (kind == 0) {
            src[q] = abase + (size_t)r.y * amul;
            mul[q] = amul;
        } else {
            uint32_t e = 0;
            for (uint32_t j = 0; j < nobs; j += 2) {
                const uint4 ob = ldg4(p.prog + pc);
                pc += 4;
                e += ob.x * ev[ob.y];
                if (j + 1 < nobs) e += ob.z * ev[ob.w];
            }
            src[q] = reinterpret_cast<const double *>(((uint64_t)r.w << 32) | (uint64_t)r.y) + e;
            mul[q] = 1;
        }
        sxm[q] = r.z * mul[q];
    }
    const uint32_t *tab = p.offtab + tab_off;
    if (cx == 2) zacc = fused_entries<K, 2, G>(src, mul, sxm, tab, n_out, cx, lane, dst, dmul, store);
    else if (cx == 1) zacc = fused_entries<K, 1, G>(src, mul, sxm, tab, n_out, cx, lane, dst, dmul, store);
    else zacc = fused_entries<K, 0, G>(src, mul, sxm, tab, n_out, cx, lane, dst, dmul, store);
    return pc;
}

template <int G>
__global__ void __launch_bounds__(kFusedThreads) ve_fused(const __grid_constant__ FusedLaunch p)
{
    extern __shared__ double arena_smem[];
    constexpr int SPW = G >= 32 ? 1 : 32 / G;       // sets per warp, interleaved in the warp's slice of the arena
    constexpr int SPC = kFusedThreads / G;          // sets per CTA
    __shared__ double s_red[kFusedThreads / 32];
    const uint32_t lane = threadIdx.x % G;
    const uint32_t set_in_cta = threadIdx.x / G;
    double *abase;
    if (G < 32) abase = arena_smem + (size_t)(threadIdx.x / 32) * SPW * p.arena + (set_in_cta % SPW);
    else abase = arena_smem + (size_t)set_in_cta * p.arena;
    constexpr uint32_t amul = SPW;
    const uint32_t n_groups = (p.nb + SPC - 1) / SPC;
    for (uint32_t grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
        uint32_t b = grp * SPC + set_in_cta;
        const bool live = b < p.nb;      // lanes of a missing set repeat the last one (they take part in the syncs), stores off
        if (!live) b = p.nb - 1;
        const uint8_t *ev = p.ev ? p.ev + (uint64_t)b * p.n_obs : p.ev_inline;
        uint32_t pc = 0;
        for (uint32_t s = 0; s < p.n_steps; ++s) {
            const uint4 h0 = ldg4(p.prog + pc), h1 = ldg4(p.prog + pc + 4);
            pc += kFusedHeaderWords;
            const uint32_t n_out = h0.x, cx = h0.y, k = h0.z & 0xffu, flags = h0.z >> 8, out_off = h0.w, tab_off = h1.x;
            double *dst;
            uint64_t dmul;
            bool store = true;
            if (flags & kFusedToResult) {
                dst = p.result + ((uint64_t)out_off * p.nb + b);
                dmul = p.nb;
                store = live;
            } else {
                dst = abase + (size_t)out_off * amul;
                dmul = amul;
            }
            double zacc = 0.0;
            switch (k) {
            case 1: pc = fused_step<1, G>(p, pc, abase, amul, ev, n_out, cx, tab_off, lane, dst, dmul, store, zacc); break;
            case 2: pc = fused_step<2, G>(p, pc, abase, amul, ev, n_out, cx, tab_off, lane, dst, dmul, store, zacc); break;
            case 3: pc = fused_step<3, G>(p, pc, abase, amul, ev, n_out, cx, tab_off, lane, dst, dmul, store, zacc); break;
            case 4: pc = fused_step<4, G>(p, pc, abase, amul, ev, n_out, cx, tab_off, lane, dst, dmul, store, zacc); break;
            case 5: pc = fused_step<5, G>(p, pc, abase, amul, ev, n_out, cx, tab_off, lane, dst, dmul, store, zacc); break;
            default: pc = fused_step<6, G>(p, pc, abase, amul, ev, n_out, cx, tab_off, lane, dst, dmul, store, zacc); break;
            }
            if ((flags & kFusedWantZ) && p.z) {      // uniform over the grid; only single queries ask for it
                double v = zacc;
                if (G <= 32) {
#pragma unroll
                    for (int o = (G < 32 ? G : 32) / 2; o > 0; o >>= 1) v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
                } else {
                    v = warp_sum(v);
                    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
                    __syncthreads();
                    v = 0.0;
#pragma unroll
                    for (int i = 0; i < kFusedThreads / 32; ++i) v = __dadd_rn(v, s_red[i]);
                }
                if (lane == 0 && live) p.z[b] = v;
            }
            group_sync<G>();     // the step's output is complete, and its operands are dead, before the next step
        }
    }
}

bool fused_valid_g(int G) { return G == 8 || G == 16 || G == 32 || G == 128; }

size_t fused_smem_bytes(int G, uint32_t arena)
{
    return (size_t)(kFusedThreads / G) * (arena ? arena : 1) * sizeof(double);
}

typedef void (*fused_fn)(const FusedLaunch);

int fused_launch(bnpp_ctx *ctx, int G, const FusedLaunch &p)
{
    fused_fn fn = nullptr;
    switch (G) {
    case 8: fn = ve_fused<8>; break;
    case 16: fn = ve_fused<16>; break;
    case 32: fn = ve_fused<32>; break;
    case 128: fn = ve_fused<128>; break;
    default: return fail(ctx, BNPP_EINVAL, "fused VE: lanes per set must be 8, 16, 32 or 128");
    }
    const size_t smem = fused_smem_bytes(G, p.arena);
    static std::map<const void *, size_t> granted;      // dynamic shared memory opted into, per variant
    size_t &have = granted[reinterpret_cast<const void *>(fn)];
    if (smem > have) {
        BNPP_CUDA(ctx, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        have = smem;
    }
    int per_sm = 0;
    BNPP_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, kFusedThreads, smem));
    if (per_sm < 1) return fail(ctx, BNPP_ETOOBIG, "fused VE: the arena of one CTA does not fit in shared memory");
    // leave part of the SM's 256 KB to L1: the program, the offset tables and the CPTs are read through it
    int cap = 6;
    if (const char *e = getenv("BNPP_FUSED_CTAS_PER_SM")) cap = atoi(e) > 0 ? atoi(e) : cap;
    if (per_sm > cap) per_sm = cap;
    const uint32_t spc = kFusedThreads / G;
    uint64_t blocks = ((uint64_t)p.nb + spc - 1) / spc;
    const uint64_t resident = (uint64_t)ctx->sm_count * per_sm;
    if (blocks > resident) blocks = resident;
    if (blocks < 1) blocks = 1;
    void *args[1] = {const_cast<FusedLaunch *>(&p)};
    BNPP_CUDA(ctx, cudaLaunchKernel(reinterpret_cast<const void *>(fn), dim3((unsigned)blocks), dim3(kFusedThreads), args, smem,
                                    ctx->stream));
    ctx->launches++;
    ctx->last_desc = nullptr;
    ctx->last_kernel = G == 8 ? "ve_fused<G=8>" : (G == 16 ? "ve_fused<G=16>" : (G == 32 ? "ve_fused<G=32>" : "ve_fused<G=128>"));
    ctx->last_grid = (uint32_t)blocks;
    ctx->last_block = kFusedThreads;
    return BNPP_OK;
}

}  // namespace bnpp
