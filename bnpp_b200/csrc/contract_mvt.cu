// Multi-valued elimination, TMA-staged:  out[o] = sum_{x < cx} prod_k F_k[pi_k(o, x)],  cx > 2.
//
// The hot variant for the canonical layout of ve.cu (and any layout in which the operand entries one TILE of the
// output needs lie in a compact range of the operand).  Replaces `prod *= *pf` over a bucket, then
// `prod.sum_out(var)` -- code/model.cpp:414-418, code/factor.cpp:117-147, 182-212.
//
//   tile    = T consecutive output entries = [g digits of one "split" axis] x [all axes inside it]
//   operand = for every tile ONE contiguous range of `range_q` doubles (base depends on the tile, what an output
//             entry reads inside the range does not): the range is brought into shared memory by ONE bulk copy
//             (cp.async.bulk global -> shared, the TMA engine, completion counted on an mbarrier) -- no thread, no
//             register and no L1 wavefront is spent on moving operands, and S stages of tiles are in flight per CTA.
//   compute = thread j owns output entry j of the tile: per value x it reads one double per operand from the
//             staged ranges (row offset of entry j from a small table, + x * stride), multiplies in the
//             reference's order and adds in the reference's order (0 + p(x=0) + p(x=1) + ...), then stores out[j]
//             (consecutive threads, consecutive addresses).  Per union entry: K shared loads, K-1 DMUL, 1 DADD.
//
// A bulk copy needs 16-byte aligned addresses and sizes; operand bases are only 8-byte aligned (odd strides), so
// the staged copy keeps the PARITY of the global index (entry i of the range sits at shared index i + (base & 1)),
// the bulk copy moves the 16-byte aligned interior and the (at most one) head and tail doubles are moved by the
// producer thread itself before it arrives on the barrier.  Nothing outside the operand's own range is read.
//
// Producers are threads 0..K-1 (one operand each: tile base by multiply-high mixed-radix decomposition, head/tail,
// bulk copy) and thread K (output base); a stage is re-armed right after the barrier that ends its tile.
// Every entry is bit-identical to code/factor.cpp (__dmul_rn / __dadd_rn, no FMA).  HBM-bound: algorithmic bytes
// 8 * (sum #F_k + #out), SURVEY 8d.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "contract.hpp"
#include "contract_mv.hpp"

extern "C" int bnpp_alloc(bnpp_ctx *ctx, uint64_t n_doubles, double **dptr);
extern "C" int bnpp_free(bnpp_ctx *ctx, double *dptr);

namespace bnpp {

namespace {

__device__ __forceinline__ uint32_t saddr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void bar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile("{\n.reg .pred P1;\nMVT_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra MVT_DONE;\nbra MVT_WAIT;\nMVT_DONE:\n}"
                 ::"r"(bar), "r"(parity)
                 : "memory");
}

// what thread `who` contributes to the stage of tile t: operand who's range (who < K) or the output base (who == K)
template <int K>
__device__ __forceinline__ void mvt_produce(const ParamsMVT &p, uint32_t t, uint32_t who, double *stage, uint32_t *meta,
                                            uint32_t bar)
{
    uint32_t outer = t, c = 0;
    if (p.n_split > 1) {
        outer = fastdiv(t, p.dsplit);
        c = t - outer * p.n_split;
    }
    if (who == (uint32_t)K) {
        // the outer axes are walked in the PLAN's order (axes a large operand lacks fastest: its tile stays in L2), so
        // the tile's place in the output comes from per-axis output strides
        uint32_t ob = c * p.g * p.inner, rem = outer;
#pragma unroll 1
        for (int a = (int)p.R - 1; a > 0; --a) {
            const uint32_t q = fastdiv(rem, p.div[a]);
            ob += (rem - q * p.div[a].d) * p.so[a];
            rem = q;
        }
        if (p.R > 0) ob += rem * p.so[0];
        meta[K] = ob;
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
        return;
    }
    uint32_t g0 = c * p.g * p.s_split[who], rem = outer;
#pragma unroll 1
    for (int a = (int)p.R - 1; a > 0; --a) {
        const uint32_t q = fastdiv(rem, p.div[a]);
        g0 += (rem - q * p.div[a].d) * p.s[who][a];
        rem = q;
    }
    if (p.R > 0) g0 += rem * p.s[who][0];
    const double *src = p.h.in[who];
    double *dst = stage + p.soff[who];
    const uint32_t shift = g0 & 1u, range = p.range[who];
    const uint32_t a = g0 + shift, b = (g0 + range) & ~1u;        // the 16-byte aligned interior [a, b)
    meta[who] = shift;
    if (shift) dst[1] = src[g0];
    if ((g0 + range) & 1u) dst[range - 1 + shift] = src[g0 + range - 1];
    const uint32_t bytes = b > a ? (b - a) * 8u : 0u;
    if (bytes) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(saddr(dst + 2u * shift)), "l"(src + a), "r"(bytes), "r"(bar)
                     : "memory");
    } else {
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
    }
}

// A row of 2 * P doubles per thread at a pitch of 2 * P doubles puts the 8 lanes of a 128-bit shared load on 8 / P
// distinct 16-byte columns (P = 2: 2-way, P = 4: 4-way bank conflicts).  Reading the pairs of a row in an order ROTATED
// by r -- r differs between the lanes that would collide -- removes the conflicts.
template <int K, int P>
__device__ __forceinline__ double row_sum_rot(const double *const (&row)[K], uint32_t r)
{
    // lane-dependent ADDRESSES, one instruction stream: a switch over the rotation would run its cases one after the
    // other with a quarter of the lanes each, and give back the wavefronts the rotation saves
    double2 pr[P];      // pr[i]: the products of pair (i + r) mod P
#pragma unroll
    for (int i = 0; i < P; ++i) {
        const uint32_t pi = (i + r) & (P - 1);
        double2 a = *reinterpret_cast<const double2 *>(row[0] + 2 * pi);
#pragma unroll
        for (int k = 1; k < K; ++k) {
            const double2 b = *reinterpret_cast<const double2 *>(row[k] + 2 * pi);
            a.x = __dmul_rn(a.x, b.x);
            a.y = __dmul_rn(a.y, b.y);
        }
        pr[i] = a;
    }
    // back into index order with selects (a barrel rotation by r), so the sum runs over x = 0, 1, ... like the reference's
#pragma unroll
    for (int b = 1; b < P; b <<= 1) {
        const bool on = (r & b) != 0;
        double2 t[P];
#pragma unroll
        for (int s = 0; s < P; ++s) {
            const double2 moved = pr[(s + P - b) % P];
            t[s].x = on ? moved.x : pr[s].x;
            t[s].y = on ? moved.y : pr[s].y;
        }
#pragma unroll
        for (int s = 0; s < P; ++s) pr[s] = t[s];
    }
    double acc = 0.0;
#pragma unroll
    for (int s = 0; s < P; ++s) {
        acc = __dadd_rn(acc, pr[s].x);
        acc = __dadd_rn(acc, pr[s].y);
    }
    return acc;
}

// MODE 0: any stride of the eliminated variable.  MODE 1: every operand has it at stride 1 (the canonical layout):
// immediate offsets.  MODE 2: stride 1, an even cardinality and every row on a 16-byte boundary: two values per
// shared load (LDS.128) -- rows of an even number of doubles put the lanes of a warp on few banks, and the wider
// load halves the wavefronts that costs.
template <int K, int MODE>
__global__ void __launch_bounds__(kBlock) contract_mvt(const __grid_constant__ ParamsMVT p)
{
    extern __shared__ __align__(128) unsigned char mvt_smem[];
    __shared__ __align__(8) uint64_t s_bar[kMvtMaxStages];
    __shared__ uint32_t s_meta[kMvtMaxStages][kMaxK + 1];
    uint32_t *rowtab = reinterpret_cast<uint32_t *>(mvt_smem);                      // [T][K]
    double *stages = reinterpret_cast<double *>(mvt_smem + p.rowtab_bytes);        // S x stage_doubles
    const uint32_t tid = threadIdx.x, S = p.stages, T = p.T, cx = p.h.cx;

    if (tid == 0) {
        for (uint32_t s = 0; s < S; ++s)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(saddr(&s_bar[s])), "r"(K + 1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (uint32_t i = tid; i < T * K; i += kBlock) rowtab[i] = p.rowtab[i];
    __syncthreads();
    // prologue: the first S tiles of this CTA
    if (tid <= (uint32_t)K)
        for (uint32_t s = 0; s < S; ++s) {
            const uint64_t t = (uint64_t)blockIdx.x + (uint64_t)s * gridDim.x;
            if (t < p.n_tiles) mvt_produce<K>(p, (uint32_t)t, tid, stages + (size_t)s * p.stage_doubles, s_meta[s], saddr(&s_bar[s]));
        }

    double zacc = 0.0;
    uint32_t s = 0, parity = 0;
    for (uint64_t t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
        bar_wait(saddr(&s_bar[s]), parity);
        const double *stage = stages + (size_t)s * p.stage_doubles;
        double *dst = p.h.out + s_meta[s][K];
        for (uint32_t j = tid; j < T; j += kBlock) {
            const double *row[K];
#pragma unroll
            for (int k = 0; k < K; ++k) row[k] = stage + p.soff[k] + (rowtab[j * K + k] + s_meta[s][k]);
            double acc = 0.0;
            if (MODE == 2 && cx == 8) {
                acc = row_sum_rot<K, 4>(row, (j >> 1) & 3u);
            } else if (MODE == 2 && cx == 4) {
                acc = row_sum_rot<K, 2>(row, (j >> 2) & 1u);
            } else if (MODE == 2) {
#pragma unroll 2
                for (uint32_t x = 0; x < cx; x += 2) {
                    double2 a = *reinterpret_cast<const double2 *>(row[0] + x);
#pragma unroll
                    for (int k = 1; k < K; ++k) {
                        const double2 b = *reinterpret_cast<const double2 *>(row[k] + x);
                        a.x = __dmul_rn(a.x, b.x);
                        a.y = __dmul_rn(a.y, b.y);
                    }
                    acc = __dadd_rn(acc, a.x);
                    acc = __dadd_rn(acc, a.y);
                }
            } else if (MODE == 1) {
#pragma unroll 4
                for (uint32_t x = 0; x < cx; ++x) {
                    double a = row[0][x];
#pragma unroll
                    for (int k = 1; k < K; ++k) a = __dmul_rn(a, row[k][x]);
                    acc = __dadd_rn(acc, a);
                }
            } else {
#pragma unroll 2
                for (uint32_t x = 0; x < cx; ++x) {
                    double a = row[0][x * p.h.sx[0]];
#pragma unroll
                    for (int k = 1; k < K; ++k) a = __dmul_rn(a, row[k][x * p.h.sx[k]]);
                    acc = __dadd_rn(acc, a);
                }
            }
            dst[j] = acc;
            zacc = __dadd_rn(zacc, acc);
        }
        __syncthreads();          // every thread is done with this stage: re-arm it for the tile S rounds ahead
        const uint64_t tn = t + (uint64_t)S * gridDim.x;
        if (tid <= (uint32_t)K && tn < p.n_tiles)
            mvt_produce<K>(p, (uint32_t)tn, tid, stages + (size_t)s * p.stage_doubles, s_meta[s], saddr(&s_bar[s]));
        if (++s == S) {
            s = 0;
            parity ^= 1u;
        }
    }
    if (p.h.z) grid_sum_to(zacc, p.h.partials, p.h.ticket, p.h.z);
}

typedef void (*mvt_fn)(const ParamsMVT);

mvt_fn pick_mvt(int k, int mode)
{
    switch (k * 4 + mode) {
    case 4: return contract_mvt<1, 0>;
    case 5: return contract_mvt<1, 1>;
    case 6: return contract_mvt<1, 2>;
    case 8: return contract_mvt<2, 0>;
    case 9: return contract_mvt<2, 1>;
    case 10: return contract_mvt<2, 2>;
    case 12: return contract_mvt<3, 0>;
    case 13: return contract_mvt<3, 1>;
    case 14: return contract_mvt<3, 2>;
    case 16: return contract_mvt<4, 0>;
    case 17: return contract_mvt<4, 1>;
    case 18: return contract_mvt<4, 2>;
    case 20: return contract_mvt<5, 0>;
    case 21: return contract_mvt<5, 1>;
    case 22: return contract_mvt<5, 2>;
    case 24: return contract_mvt<6, 0>;
    case 25: return contract_mvt<6, 1>;
    case 26: return contract_mvt<6, 2>;
    default: return nullptr;
    }
}

int mvt_resident(bnpp_ctx *ctx, mvt_fn fn, unsigned smem)
{
    static std::mutex mu;
    static std::map<std::pair<int, const void *>, unsigned> granted;
    std::lock_guard<std::mutex> lock(mu);
    const auto key = std::make_pair(ctx->device, reinterpret_cast<const void *>(fn));
    smem = ((smem + 1023u) >> 10) << 10;        // opt-in and occupancy are asked per KB of shared memory (cached)
    auto it = granted.find(key);
    if (it == granted.end() || it->second < smem) {
        if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
            cudaGetLastError();
            return 0;
        }
        granted[key] = smem;
    }
    static std::map<std::pair<std::pair<int, const void *>, unsigned>, int> occupancy;       // (device, fn, KB of shared memory)
    const auto okey = std::make_pair(key, smem >> 10);
    auto oc = occupancy.find(okey);
    if (oc != occupancy.end()) return oc->second;
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, kBlock, smem) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    occupancy[okey] = per_sm;
    return per_sm;
}

}  // namespace

int plan_mvt(bnpp_ctx *ctx, LaunchDesc *d, int k, uint32_t cx, const uint64_t *sx, const uint64_t *op_bytes,
             const std::vector<MVAxis> &axes, uint64_t n_out, const ParamsHead &h)
{
    if (k < 1 || k > kMaxK || cx < 2 || cx > 64) return 1;
    const int n = (int)axes.size();
    for (int q = 0; q < k; ++q)
        if ((reinterpret_cast<uintptr_t>(h.in[q]) & 15u) != 0) return 1;       // the bulk copy wants 16-byte aligned tables

    // range of operand q touched by a tile [g digits of axis `split`] x [axes inside]
    auto range_of = [&](int q, int split, uint32_t g) {
        uint64_t top = (uint64_t)(cx - 1) * sx[q];
        for (int a = split + 1; a < n; ++a) top += (uint64_t)(axes[a].ext - 1) * axes[a].s[q];
        if (split >= 0) top += (uint64_t)(g - 1) * axes[split].s[q];
        return top + 1;
    };
    const uint64_t smem_budget = 110u << 10;        // two CTAs per SM
    const uint64_t e_max = 8192;
    // candidates: every suffix of axes as the inner block, every divisor g of the next axis; keep the best tile
    double best = -1.0;
    int best_split = -2;
    uint32_t best_g = 1;
    uint64_t inner = 1;
    for (int split = n - 1; split >= -1; --split) {
        // `inner` = product of the axes inside `split`
        const uint32_t ext = split >= 0 ? axes[split].ext : 1;
        for (uint32_t g = 1; g <= ext; ++g) {
            if (ext % g) continue;
            if (split >= 0 && g == ext && split > 0) continue;      // the same tile as (split - 1, g = 1) up to the next axis
            const uint64_t T = inner * g, E = T * cx;
            if (E > e_max || T >= (1u << 16)) break;
            uint64_t stage = 0, useful = 0;
            bool ok = true;
            for (int q = 0; q < k && ok; ++q) {
                const uint64_t r = range_of(q, split, g);
                uint64_t distinct = sx[q] ? cx : 1;
                for (int a = std::max(split, 0); a < n; ++a)
                    if (axes[a].s[q]) distinct *= (a == split ? g : axes[a].ext);
                ok = r <= 2 * distinct + 16 && r < (1u << 24);
                stage += (r + 3) & ~(uint64_t)1;
                useful += distinct;
            }
            static const uint64_t min_stages = [] {
                const char *e = getenv("BNPP_MVT_MIN_STAGES");
                return e ? (uint64_t)atoi(e) : 2ull;
            }();
            if (!ok || min_stages * stage * 8 + T * k * 4 > smem_budget) continue;
            const double lanes = (double)T / (double)(((T + kBlock - 1) / kBlock) * kBlock);
            const double amort = (double)E / (double)(E + 256);
            const double score = lanes * amort;
            if (score > best) {
                best = score;
                best_split = split;
                best_g = g;
            }
        }
        if (split >= 0) {
            inner *= axes[split].ext;
            if (inner * cx > e_max) break;
        }
    }
    if (best_split == -2) return 1;
    const int split = best_split;
    const uint32_t g = best_g;
    inner = 1;
    for (int a = split + 1; a < n; ++a) inner *= axes[a].ext;
    const uint32_t ext_split = split >= 0 ? axes[split].ext : 1;
    const uint32_t T = (uint32_t)(inner * g);
    if (T < 64 && n_out > 4096) return 1;       // too few lanes per tile: the gather kernel

    // outer axes (outside the split axis).  Above the tile the iteration order is the plan's: an axis that a large
    // operand LACKS is walked fastest, so that operand's tile range is re-read from L2 instead of HBM (the output is
    // addressed through per-axis strides, its layout is untouched).  Neighbours contiguous in every operand and in
    // the output merge.
    struct Outer { uint64_t ext; uint64_t s[kMaxK]; uint64_t so; uint64_t miss; };
    std::vector<Outer> raw;
    {
        uint64_t so = (uint64_t)inner * ext_split;
        for (int a = split - 1; a >= 0; --a) {
            Outer o;
            o.ext = axes[a].ext;
            o.so = so;
            o.miss = 0;
            for (int q = 0; q < kMaxK; ++q) {
                o.s[q] = q < k ? axes[a].s[q] : 0;
                if (q < k && axes[a].s[q] == 0 && op_bytes[q] > (32ull << 20)) o.miss += op_bytes[q];
            }
            so *= axes[a].ext;
            raw.insert(raw.begin(), o);
        }
        bool any = false;
        for (const Outer &o : raw) any = any || o.miss != 0;
        if (any) std::stable_sort(raw.begin(), raw.end(), [](const Outer &x, const Outer &y) { return x.miss < y.miss; });
    }
    std::vector<Outer> outer;
    for (const Outer &o : raw) {
        if (!outer.empty()) {
            Outer &up = outer.back();
            bool ok = up.ext * o.ext < (1ull << 32) && up.so == o.so * o.ext;
            for (int q = 0; q < k && ok; ++q) ok = (up.s[q] == o.s[q] * o.ext);
            if (ok) {
                up.ext *= o.ext;
                up.so = o.so;
                for (int q = 0; q < k; ++q) up.s[q] = o.s[q];
                continue;
            }
        }
        outer.push_back(o);
    }
    if ((int)outer.size() > kMaxR) return 1;

    ParamsMVT &p = d->mvtp;
    memset(&p, 0, sizeof p);
    p.h = h;
    p.h.n_items = n_out;
    p.h.cx = cx;
    p.T = T;
    p.inner = (uint32_t)inner;
    p.g = g;
    p.ext_split = ext_split;
    p.n_split = ext_split / g;
    p.R = (uint32_t)outer.size();
    uint64_t n_outer = 1;
    for (size_t a = 0; a < outer.size(); ++a) {
        p.div[a] = make_fastdiv((uint32_t)outer[a].ext);
        n_outer *= outer[a].ext;
        if (outer[a].so >= (1ull << 32)) return 1;
        p.so[a] = (uint32_t)outer[a].so;
        for (int q = 0; q < k; ++q) {
            if (outer[a].s[q] >= (1ull << 32)) return 1;
            p.s[q][a] = (uint32_t)outer[a].s[q];
        }
    }
    if (n_outer * p.n_split >= (1ull << 32)) return 1;
    p.n_tiles = (uint32_t)(n_outer * p.n_split);
    if (p.n_split > 1) p.dsplit = make_fastdiv(p.n_split);
    bool unit = true;
    uint64_t stage_doubles = 0;
    for (int q = 0; q < k; ++q) {
        const uint64_t s = split >= 0 ? axes[split].s[q] : 0;
        if (s >= (1ull << 32) || sx[q] >= (1ull << 32)) return 1;
        p.s_split[q] = (uint32_t)s;
        p.h.sx[q] = (uint32_t)sx[q];
        unit = unit && sx[q] == 1;
        p.range[q] = (uint32_t)range_of(q, split, g);
        p.soff[q] = (uint32_t)stage_doubles;
        stage_doubles += ((uint64_t)p.range[q] + 3) & ~(uint64_t)1;     // + parity shift, rounded to 16 bytes
    }
    p.stage_doubles = (uint32_t)stage_doubles;
    p.rowtab_bytes = (uint32_t)((((size_t)T * k * 4) + 127) & ~(size_t)127);
    uint32_t stages = (uint32_t)((smem_budget - p.rowtab_bytes) / (stage_doubles * 8));
    stages = std::min<uint32_t>(stages, kMvtMaxStages);
    if (stages < 2) return 1;
    p.stages = stages;

    // row offsets of the tile's output entries, per operand
    std::vector<uint32_t> tab((size_t)T * k + 4, 0);
    {
        const int first = std::max(split, 0);
        std::vector<uint32_t> digit(n, 0);
        uint64_t loc[kMaxK] = {0};
        for (uint32_t o = 0; o < T; ++o) {
            for (int q = 0; q < k; ++q) tab[(size_t)o * k + q] = (uint32_t)loc[q];
            for (int a = n - 1; a >= first; --a) {
                if (++digit[a] < axes[a].ext || a == first) {
                    for (int q = 0; q < k; ++q) loc[q] += axes[a].s[q];
                    break;
                }
                digit[a] = 0;
                for (int q = 0; q < k; ++q) loc[q] -= (uint64_t)(axes[a].ext - 1) * axes[a].s[q];
            }
        }
    }
    // two values per shared load: stride 1, even cardinality, and every row of every tile on a 16-byte boundary
    bool pair = unit && cx % 2 == 0;
    for (size_t i = 0; i < (size_t)T * k && pair; ++i) pair = tab[i] % 2 == 0;
    for (int q = 0; q < k && pair; ++q) {
        pair = p.s_split[q] % 2 == 0;
        for (uint32_t a = 0; a < p.R && pair; ++a) pair = p.s[q][a] % 2 == 0;
    }
    const int mode = pair ? 2 : (unit ? 1 : 0);
    mvt_fn fn = pick_mvt(k, mode);
    if (!fn) return 1;
    const unsigned smem = p.rowtab_bytes + stages * p.stage_doubles * 8u;
    const int per_sm = mvt_resident(ctx, fn, smem);
    if (per_sm < 1) return 1;
    double *store = nullptr;
    int rc = bnpp_alloc(ctx, tab.size() / 2 + 2, &store);
    if (rc != BNPP_OK) return rc;
    rc = stage_upload(ctx, store, tab.data(), tab.size() * 4);
    if (rc != BNPP_OK) {
        bnpp_free(ctx, store);
        return rc;
    }
    d->mv_tab = reinterpret_cast<uint32_t *>(store);
    p.rowtab = d->mv_tab;
    d->mv = false;
    d->mvt = true;
    d->p2 = false;
    d->staged = false;
    d->smem = smem;
    d->fn = reinterpret_cast<const void *>(fn);
    d->grid = (unsigned)std::min<uint64_t>(p.n_tiles, (uint64_t)ctx->sm_count * per_sm);
    d->k = k;
    d->variant = mode == 2 ? "mvt/pair" : (mode == 1 ? "mvt/unit" : "mvt/strided");
    d->C = (int)cx;
    d->V = 1;
    d->U = (int)stages;
    d->div = false;
    d->generic = false;
    d->R = p.R;
    return BNPP_OK;
}

}  // namespace bnpp
