// K7 -- factor-graph sum-product (loopy BP) with every message of a phase updated in
// one launch.  Replaces FactorGraph (reference code/graph.cpp:256-403).
//
// The reference sweeps edge by edge; within a phase each update reads only the other
// direction's messages (code/graph.cpp:340-359, 367-388), so a flood over all edges
// is the same schedule and the converging sweep index is preserved.
// Messages live in two flat device arrays indexed by edge e = foff[f] + slot; the
// host never touches them between sweeps except for the one max-error word.
#include <cooperative_groups.h>

#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace cg = cooperative_groups;

struct bnpp_fg {
    bnpp_ctx *ctx = nullptr;
    int nvars = 0, nfac = 0, nedges = 0;
    uint32_t nmsg = 0;              // total message entries
    uint32_t *card = nullptr;       // [nvars]
    int32_t *foff = nullptr;        // [nfac+1]
    uint32_t *evar = nullptr;       // [nedges] variable of edge
    int32_t *efac = nullptr;        // [nedges] factor of edge
    uint32_t *moff = nullptr;       // [nedges+1] message offsets
    int32_t *voff = nullptr;        // [nvars+1] CSR variable -> edges
    int32_t *vedges = nullptr;
    uint64_t *toff = nullptr;       // [nfac] table offsets
    uint32_t *estride = nullptr;    // [nedges] stride of the edge's axis inside its factor table
    uint32_t *fsize = nullptr;      // [nfac]
    double *ftab = nullptr;
    double *f2v = nullptr, *v2f = nullptr, *tmp = nullptr;
    unsigned long long *maxerr = nullptr;       // device: bit pattern of the sweep's max error
    unsigned long long *maxerr_host = nullptr;  // pinned
    uint32_t *mvoff = nullptr;      // [nvars+1] marginal output offsets
    double *marg = nullptr;
    uint32_t nmarg = 0;
    std::vector<uint32_t> h_card;
    // FactorGraph::update as ONE cooperative launch: three rotating max-error words and the sweep count
    unsigned long long *err3 = nullptr;         // device [3] + sweeps word [1]
    uint32_t *sweeps_host = nullptr;            // pinned
    int coop_blocks = 0;                        // 0 = not available (falls back to a launch per phase)
    // factor -> variable updates: a thread per edge where the factor is small (fsize / card <= kSmallSub entries to sum
    // per message entry: every Ising / pairwise factor), a warp per edge otherwise
    int n_small = 0, n_big = 0;
    int32_t *small_edges = nullptr, *big_edges = nullptr;
    // pre-resolved reads of every message update (built once, host side): what a sweep costs is the LATENCY of a chain of
    // dependent loads, so each update gets everything it needs from one record -- no CSR walk, no div/mod per term
    //   variable -> factor : a_hdr[e] = (first word in a_nb, count | card << 16); a_nb = message offsets of the variable's OTHER edges
    //   factor -> variable (small factors): b_hdr[e] = (r | sub << 8 | w << 16, first word in b_terms, table offset lo, hi);
    //       b_terms = for i < r, ts < sub: table index, then the message entry of each other slot (w - 1 words)
    uint2 *a_hdr = nullptr;
    uint32_t *a_nb = nullptr;
    uint4 *b_hdr = nullptr;
    uint32_t *b_terms = nullptr;
};

namespace bnpp {

__global__ void fg_init_kernel(int nedges, const uint32_t *evar, const uint32_t *card, const uint32_t *moff,
                               double *f2v, double *v2f)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nedges) return;
    const uint32_t r = card[evar[e]];
    const double u = 1.0 / r;   // code/graph.cpp:271-272
    for (uint32_t i = 0; i < r; ++i) {
        f2v[moff[e] + i] = u;
        v2f[moff[e] + i] = u;
    }
}

__device__ __forceinline__ void note_error(unsigned long long *maxerr, double err)
{
    // reference: `if (err > maxerror)` starting from 0.0 -- NaN and negatives never count, +inf does
    if (err > 0.0) atomicMax(maxerr, (unsigned long long)__double_as_longlong(err));
}

// one atomic per warp instead of one per message: the max-error word is a single L2 address, and 15 680 atomics on
// it per phase (a 40x40 Ising grid) cost more than the phase itself
__device__ __forceinline__ void note_error_warp(unsigned long long *maxerr, double worst)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double other = __shfl_xor_sync(0xffffffffu, worst, o);
        if (other > worst) worst = other;
    }
    if ((threadIdx.x & 31) == 0) note_error(maxerr, worst);
}

// variable -> factor, code/graph.cpp:334-362: m_{v->f} = normalize(prod_{g in N(v)\f} m_{g->v})
__device__ __forceinline__ double var_to_fac_edge(int e, const uint32_t *__restrict__ evar, const uint32_t *__restrict__ card,
                                                const uint32_t *__restrict__ moff, const int32_t *__restrict__ voff,
                                                const int32_t *__restrict__ vedges, const double *f2v, double *v2f)
{
    const uint32_t v = evar[e], r = card[v];
    const int b = voff[v], n = voff[v + 1];
    double z = 0.0;
    for (uint32_t i = 0; i < r; ++i) {
        double p = 1.0;
        for (int q = b; q < n; ++q) {
            const int e2 = vedges[q];
            if (e2 != e) p *= __ldcg(f2v + moff[e2] + i);
        }
        z += p;
    }
    double worst = 0.0;
    for (uint32_t i = 0; i < r; ++i) {
        double p = 1.0;
        for (int q = b; q < n; ++q) {
            const int e2 = vedges[q];
            if (e2 != e) p *= __ldcg(f2v + moff[e2] + i);
        }
        const double nv = p / z;
        const double ov = __ldcg(v2f + moff[e] + i);
        const double err = fabs(ov - nv) / ov;
        if (err > worst) worst = err;
        v2f[moff[e] + i] = nv;
    }
    return worst;
}

constexpr int kMaxNb = 8;       // neighbours / terms kept in registers by the record-driven updates

// variable -> factor from the edge's record: all neighbour messages are requested before the first multiply
__device__ __forceinline__ double var_to_fac_rec(int e, const uint2 *__restrict__ a_hdr, const uint32_t *__restrict__ a_nb,
                                                 const uint32_t *__restrict__ moff, const double *f2v, double *v2f)
{
    const uint2 h = __ldg(a_hdr + e);
    const uint32_t n = h.y & 0xffffu, r = h.y >> 16, mo = __ldg(moff + e);
    if (r == 2 && n <= (uint32_t)kMaxNb) {
        uint32_t nb[kMaxNb];
#pragma unroll
        for (int q = 0; q < kMaxNb; ++q) nb[q] = q < (int)n ? __ldg(a_nb + h.x + q) : 0u;
        double m0[kMaxNb], m1[kMaxNb];
#pragma unroll
        for (int q = 0; q < kMaxNb; ++q)
            if (q < (int)n) {
                m0[q] = __ldcg(f2v + nb[q]);
                m1[q] = __ldcg(f2v + nb[q] + 1);
            }
        const double o0 = __ldcg(v2f + mo), o1 = __ldcg(v2f + mo + 1);
        double p0 = 1.0, p1 = 1.0;
#pragma unroll
        for (int q = 0; q < kMaxNb; ++q)
            if (q < (int)n) {
                p0 *= m0[q];
                p1 *= m1[q];
            }
        double z = 0.0;
        z += p0;
        z += p1;
        const double n0 = p0 / z, n1 = p1 / z;
        double worst = 0.0;
        const double e0 = fabs(o0 - n0) / o0, e1 = fabs(o1 - n1) / o1;
        if (e0 > worst) worst = e0;
        if (e1 > worst) worst = e1;
        v2f[mo] = n0;
        v2f[mo + 1] = n1;
        return worst;
    }
    // any cardinality / degree: two passes over the neighbour list
    double z = 0.0;
    for (uint32_t i = 0; i < r; ++i) {
        double p = 1.0;
        for (uint32_t q = 0; q < n; ++q) p *= __ldcg(f2v + __ldg(a_nb + h.x + q) + i);
        z += p;
    }
    double worst = 0.0;
    for (uint32_t i = 0; i < r; ++i) {
        double p = 1.0;
        for (uint32_t q = 0; q < n; ++q) p *= __ldcg(f2v + __ldg(a_nb + h.x + q) + i);
        const double nv = p / z;
        const double ov = __ldcg(v2f + mo + i);
        const double err = fabs(ov - nv) / ov;
        if (err > worst) worst = err;
        v2f[mo + i] = nv;
    }
    return worst;
}

// factor -> variable for a small factor from the edge's record
__device__ __forceinline__ double fac_to_var_rec(int e, const uint4 *__restrict__ b_hdr, const uint32_t *__restrict__ b_terms,
                                                 const uint32_t *__restrict__ moff, const double *__restrict__ ftab, const double *v2f,
                                                 double *f2v, double *tmp)
{
    const uint4 h = __ldg(b_hdr + e);
    const uint32_t r = h.x & 0xffu, sub = (h.x >> 8) & 0xffu, w = h.x >> 16, mo = __ldg(moff + e);
    const double *tab = ftab + (((uint64_t)h.w << 32) | h.z);
    const uint32_t *t = b_terms + h.y;
    if (r == 2 && w <= 2 && sub <= 2) {
        // the pairwise / unary case (every factor of an Ising grid): at most 4 terms of at most 2 factors each
        uint32_t words[8];
        const uint32_t nw = 2u * sub * w;
#pragma unroll
        for (int q = 0; q < 8; ++q) words[q] = q < (int)nw ? __ldg(t + q) : 0u;
        double val[8];
#pragma unroll
        for (int q = 0; q < 8; ++q)
            if (q < (int)nw) val[q] = ((uint32_t)q % w == 0) ? __ldg(tab + words[q]) : __ldcg(v2f + words[q]);
        const double o0 = __ldcg(f2v + mo), o1 = __ldcg(f2v + mo + 1);
        double part[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            double acc = 0.0;
#pragma unroll
            for (int ts = 0; ts < 2; ++ts)
                if (ts < (int)sub) {
                    const int q = (i * (int)sub + ts) * (int)w;
                    double p = val[q];
                    if (w == 2) p *= val[q + 1];
                    acc += p;
                }
            part[i] = acc;
        }
        double z = 0.0;
        z += part[0];
        z += part[1];
        const double n0 = part[0] / z, n1 = part[1] / z;
        double worst = 0.0;
        const double e0 = fabs(o0 - n0) / o0, e1 = fabs(o1 - n1) / o1;
        if (e0 > worst) worst = e0;
        if (e1 > worst) worst = e1;
        f2v[mo] = n0;
        f2v[mo + 1] = n1;
        return worst;
    }
    double z = 0.0;
    for (uint32_t i = 0; i < r; ++i) {
        double part = 0.0;
        for (uint32_t ts = 0; ts < sub; ++ts) {
            const uint32_t *rec = t + (i * sub + ts) * w;
            double p = __ldg(tab + __ldg(rec));
            for (uint32_t u = 1; u < w; ++u) p *= __ldcg(v2f + __ldg(rec + u));
            part += p;
        }
        tmp[mo + i] = part;
        z += part;
    }
    double worst = 0.0;
    for (uint32_t i = 0; i < r; ++i) {
        const double nv = tmp[mo + i] / z;
        const double ov = __ldcg(f2v + mo + i);
        const double err = fabs(ov - nv) / ov;
        if (err > worst) worst = err;
        f2v[mo + i] = nv;
    }
    return worst;
}

__global__ void __launch_bounds__(128) fg_var_to_fac_kernel(int nedges, const uint2 *__restrict__ a_hdr,
                                                            const uint32_t *__restrict__ a_nb,
                                                            const uint32_t *__restrict__ moff,
                                                            const double *__restrict__ f2v, double *__restrict__ v2f,
                                                            unsigned long long *maxerr)
{
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gthreads = gridDim.x * blockDim.x;
    double worst = 0.0;
    for (int e = gtid; e < nedges; e += gthreads) {
        const double w_ = var_to_fac_rec(e, a_hdr, a_nb, moff, f2v, v2f);
        if (w_ > worst) worst = w_;
    }
    note_error_warp(maxerr, worst);
}

// factor -> variable, code/graph.cpp:364-391: one warp per edge (f, slot j):
// m_{f->v}[i] = sum over the factor entries with digit_j = i of  F * prod_{u != j} m_{u->f}
__device__ __forceinline__ double fac_to_var_edge(int e, int lane, const int32_t *__restrict__ efac,
                                                const uint32_t *__restrict__ evar, const int32_t *__restrict__ foff,
                                                const uint32_t *__restrict__ card, const uint32_t *__restrict__ moff,
                                                const uint64_t *__restrict__ toff, const uint32_t *__restrict__ estride,
                                                const uint32_t *__restrict__ fsize, const double *__restrict__ ftab,
                                                const double *v2f, double *f2v, double *tmp)
{
    const int f = efac[e];
    const int e0 = foff[f], w = foff[f + 1] - e0;
    const uint32_t r = card[evar[e]], stj = estride[e];
    const uint32_t sub = fsize[f] / r;
    const double *tab = ftab + toff[f];
    for (uint32_t i = 0; i < r; ++i) {
        double part = 0.0;
        for (uint32_t ts = lane; ts < sub; ts += 32) {
            const uint32_t t = (ts / stj) * (stj * r) + i * stj + (ts % stj);
            double p = tab[t];
            for (int u = 0; u < w; ++u) {
                const int eu = e0 + u;
                if (eu == e) continue;
                const uint32_t d = (t / estride[eu]) % card[evar[eu]];
                p *= __ldcg(v2f + moff[eu] + d);
            }
            part += p;
        }
        part = warp_sum(part);
        if (lane == 0) tmp[moff[e] + i] = part;
    }
    __syncwarp();
    double z = 0.0;
    for (uint32_t i = 0; i < r; ++i) z += __ldcg(tmp + moff[e] + i);   // same order in every lane
    double worst = 0.0;
    for (uint32_t i = lane; i < r; i += 32) {
        const double nv = __ldcg(tmp + moff[e] + i) / z;
        const double ov = __ldcg(f2v + moff[e] + i);
        const double err = fabs(ov - nv) / ov;
        if (err > worst) worst = err;
        f2v[moff[e] + i] = nv;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double other = __shfl_down_sync(0xffffffffu, worst, o);
        if (other > worst) worst = other;
    }
    return worst;       // lane 0 holds the message's worst relative change
}

constexpr uint32_t kSmallSub = 8;

// the same update by ONE thread (small factors: a warp per edge would leave 30 lanes idle)
__device__ __forceinline__ double fac_to_var_edge_small(int e, const int32_t *__restrict__ efac, const uint32_t *__restrict__ evar,
                                                      const int32_t *__restrict__ foff, const uint32_t *__restrict__ card,
                                                      const uint32_t *__restrict__ moff, const uint64_t *__restrict__ toff,
                                                      const uint32_t *__restrict__ estride, const uint32_t *__restrict__ fsize,
                                                      const double *__restrict__ ftab, const double *v2f, double *f2v, double *tmp)
{
    const int f = efac[e];
    const int e0 = foff[f], w = foff[f + 1] - e0;
    const uint32_t r = card[evar[e]], stj = estride[e];
    const uint32_t sub = fsize[f] / r;
    const double *tab = ftab + toff[f];
    double z = 0.0;
    for (uint32_t i = 0; i < r; ++i) {
        double part = 0.0;
        for (uint32_t ts = 0; ts < sub; ++ts) {
            const uint32_t t = (ts / stj) * (stj * r) + i * stj + (ts % stj);
            double p = tab[t];
            for (int u = 0; u < w; ++u) {
                const int eu = e0 + u;
                if (eu == e) continue;
                const uint32_t d = (t / estride[eu]) % card[evar[eu]];
                p *= __ldcg(v2f + moff[eu] + d);
            }
            part += p;
        }
        tmp[moff[e] + i] = part;
        z += part;
    }
    double worst = 0.0;
    for (uint32_t i = 0; i < r; ++i) {
        const double nv = tmp[moff[e] + i] / z;
        const double ov = __ldcg(f2v + moff[e] + i);
        const double err = fabs(ov - nv) / ov;
        if (err > worst) worst = err;
        f2v[moff[e] + i] = nv;
    }
    return worst;
}

__global__ void __launch_bounds__(128) fg_fac_to_var_kernel(int n_small, const int32_t *__restrict__ small_edges, int n_big,
                                                            const int32_t *__restrict__ big_edges,
                                                            const uint4 *__restrict__ b_hdr, const uint32_t *__restrict__ b_terms,
                                                            const int32_t *__restrict__ efac,
                                                            const uint32_t *__restrict__ evar,
                                                            const int32_t *__restrict__ foff,
                                                            const uint32_t *__restrict__ card,
                                                            const uint32_t *__restrict__ moff,
                                                            const uint64_t *__restrict__ toff,
                                                            const uint32_t *__restrict__ estride,
                                                            const uint32_t *__restrict__ fsize,
                                                            const double *__restrict__ ftab,
                                                            const double *__restrict__ v2f, double *__restrict__ f2v,
                                                            double *__restrict__ tmp, unsigned long long *maxerr)
{
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gthreads = gridDim.x * blockDim.x;
    double worst = 0.0;
    for (int i = gtid; i < n_small; i += gthreads) {
        const double w_ = fac_to_var_rec(small_edges[i], b_hdr, b_terms, moff, ftab, v2f, f2v, tmp);
        if (w_ > worst) worst = w_;
    }
    const int lane = threadIdx.x & 31;
    for (int i = gtid >> 5; i < n_big; i += gthreads >> 5) {
        const double w_ = fac_to_var_edge(big_edges[i], lane, efac, evar, foff, card, moff, toff, estride, fsize, ftab, v2f, f2v, tmp);
        if (lane == 0 && w_ > worst) worst = w_;
    }
    note_error_warp(maxerr, worst);
}

// FactorGraph::update (code/graph.cpp:298-332) as ONE cooperative launch: every sweep is the two floods above with a
// grid-wide barrier after each; the convergence test `maxerror < epsilon` and the sweep counter stay on the device, the
// host reads (sweeps) once.  A 40x40 Ising sweep is 15 680 two-entry messages: launch latency, not work, is what a
// launch per phase pays.  The max-error word rotates over three slots so that zeroing the next sweep's slot never
// races with a slow block still reading the previous one.
struct FgArgs {
    int nedges;
    const uint32_t *evar, *card, *moff;
    const int32_t *voff, *vedges, *efac, *foff;
    const uint64_t *toff;
    const uint32_t *estride, *fsize;
    const double *ftab;
    double *f2v, *v2f, *tmp;
    int n_small, n_big;
    const int32_t *small_edges, *big_edges;
    const uint2 *a_hdr;
    const uint32_t *a_nb;
    const uint4 *b_hdr;
    const uint32_t *b_terms;
    unsigned long long *err3;      // [3] max-error slots, [3] = sweeps (as uint32)
    uint32_t max_sweeps;
    double epsilon;
};

__global__ void __launch_bounds__(256) fg_update_kernel(const FgArgs a)
{
    cg::grid_group grid = cg::this_grid();
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gthreads = gridDim.x * blockDim.x;
    const int gwarp = gtid >> 5, gwarps = gthreads >> 5, lane = threadIdx.x & 31;
    uint32_t it = 0;
    for (; it < a.max_sweeps; ++it) {
        unsigned long long *err = a.err3 + it % 3;
        if (gtid == 0) a.err3[(it + 1) % 3] = 0ull;
        double worst = 0.0;
        for (int e = gtid; e < a.nedges; e += gthreads) {
            const double w_ = var_to_fac_rec(e, a.a_hdr, a.a_nb, a.moff, a.f2v, a.v2f);
            if (w_ > worst) worst = w_;
        }
        note_error_warp(err, worst);
        grid.sync();
        worst = 0.0;
        for (int i = gtid; i < a.n_small; i += gthreads) {
            const double w_ = fac_to_var_rec(a.small_edges[i], a.b_hdr, a.b_terms, a.moff, a.ftab, a.v2f, a.f2v, a.tmp);
            if (w_ > worst) worst = w_;
        }
        for (int i = gwarp; i < a.n_big; i += gwarps) {
            const double w_ = fac_to_var_edge(a.big_edges[i], lane, a.efac, a.evar, a.foff, a.card, a.moff, a.toff, a.estride, a.fsize,
                                              a.ftab, a.v2f, a.f2v, a.tmp);
            if (lane == 0 && w_ > worst) worst = w_;
        }
        note_error_warp(err, worst);
        grid.sync();
        const double maxerror = __longlong_as_double((long long)*reinterpret_cast<volatile unsigned long long *>(err));
        if (maxerror < a.epsilon) break;          // code/graph.cpp:328, the same decision in every thread
    }
    if (gtid == 0) *reinterpret_cast<uint32_t *>(a.err3 + 3) = it;
}

// FactorGraph::marginal, code/graph.cpp:393-403
__global__ void fg_marginal_kernel(int nvars, const uint32_t *__restrict__ card, const uint32_t *__restrict__ moff,
                                   const int32_t *__restrict__ voff, const int32_t *__restrict__ vedges,
                                   const double *__restrict__ f2v, const uint32_t *__restrict__ mvoff, double *out)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nvars) return;
    const uint32_t r = card[v];
    const int b = voff[v], n = voff[v + 1];
    if (b == n) {   // variable in no factor (e.g. observed and conditioned away): the width-0 factor [1]
        for (uint32_t i = 0; i < r; ++i) out[mvoff[v] + i] = (i == 0) ? 1.0 : 0.0;
        return;
    }
    double z = 0.0;
    for (uint32_t i = 0; i < r; ++i) {
        double p = 1.0;
        for (int q = b; q < n; ++q) p *= f2v[moff[vedges[q]] + i];
        z += p;
    }
    for (uint32_t i = 0; i < r; ++i) {
        double p = 1.0;
        for (int q = b; q < n; ++q) p *= f2v[moff[vedges[q]] + i];
        out[mvoff[v] + i] = p / z;
    }
}

template <typename T>
static int to_device(bnpp_ctx *ctx, T **dst, const std::vector<T> &src)
{
    BNPP_CUDA(ctx, cudaMalloc(dst, sizeof(T) * (src.empty() ? 1 : src.size())));
    if (!src.empty())
        BNPP_CUDA(ctx, cudaMemcpyAsync(*dst, src.data(), sizeof(T) * src.size(), cudaMemcpyHostToDevice, ctx->stream));
    return BNPP_OK;
}

}  // namespace bnpp

using namespace bnpp;

extern "C" {

int bnpp_fg_create(bnpp_ctx *ctx, int nvars, const uint32_t *card, int nfac, const int32_t *foff,
                   const uint32_t *fscope, const uint64_t *toff, const double *ftab_host, bnpp_fg **out)
{
    if (!ctx || !out || nvars < 0 || nfac < 0) return BNPP_EINVAL;
    *out = nullptr;
    const int nedges = nfac ? foff[nfac] : 0;
    std::vector<uint32_t> h_card(card, card + nvars), evar(nedges), moff(nedges + 1, 0), estride(nedges), fsize(nfac);
    std::vector<int32_t> efac(nedges), h_foff(foff, foff + nfac + 1), voff(nvars + 1, 0), vedges(nedges);
    std::vector<uint64_t> h_toff(toff, toff + nfac);
    uint64_t tab_total = 0;
    for (int f = 0; f < nfac; ++f) {
        uint64_t st = 1;
        for (int e = foff[f + 1] - 1; e >= foff[f]; --e) {
            const uint32_t v = fscope[e];
            if (v >= (uint32_t)nvars) return fail(ctx, BNPP_EINVAL, "factor scope names an unknown variable");
            evar[e] = v;
            efac[e] = f;
            estride[e] = (uint32_t)st;
            st *= card[v];
            if (st >= (1ull << 32)) return fail(ctx, BNPP_ETOOBIG, "factor table has >= 2^32 entries");
        }
        fsize[f] = (uint32_t)st;
        if (toff[f] + st > tab_total) tab_total = toff[f] + st;
    }
    for (int e = 0; e < nedges; ++e) {
        moff[e + 1] = moff[e] + card[evar[e]];
        voff[evar[e] + 1]++;
    }
    for (int v = 0; v < nvars; ++v) voff[v + 1] += voff[v];
    {
        std::vector<int32_t> fillp(voff.begin(), voff.end() - 1);
        for (int e = 0; e < nedges; ++e) vedges[fillp[evar[e]]++] = e;
    }
    std::vector<uint32_t> mvoff(nvars + 1, 0);
    for (int v = 0; v < nvars; ++v) mvoff[v + 1] = mvoff[v] + card[v];
    std::vector<int32_t> small_edges, big_edges;
    for (int e = 0; e < nedges; ++e) {
        const uint32_t r = card[evar[e]], sub = fsize[efac[e]] / r, w = (uint32_t)(foff[efac[e] + 1] - foff[efac[e]]);
        (sub <= kSmallSub && r < 256 && w < 65536 ? small_edges : big_edges).push_back(e);
    }
    // records (see bnpp_fg): the reads of every update, resolved once
    std::vector<uint2> a_hdr(nedges ? nedges : 1);
    std::vector<uint32_t> a_nb;
    for (int e = 0; e < nedges; ++e) {
        const uint32_t v = evar[e];
        a_hdr[e].x = (uint32_t)a_nb.size();
        uint32_t n = 0;
        for (int q = voff[v]; q < voff[v + 1]; ++q)
            if (vedges[q] != e) {
                a_nb.push_back(moff[vedges[q]]);
                ++n;
            }
        if (n >= 65536 || card[v] >= 65536) return fail(ctx, BNPP_ETOOBIG, "factor graph: a variable with >= 65536 factors or values");
        a_hdr[e].y = n | (card[v] << 16);
    }
    std::vector<uint4> b_hdr(nedges ? nedges : 1);
    std::vector<uint32_t> b_terms;
    for (int e : small_edges) {
        const int f = efac[e], e0 = foff[f], w = foff[f + 1] - e0;
        const uint32_t r = card[evar[e]], stj = estride[e], sub = fsize[f] / r;
        b_hdr[e] = make_uint4(r | (sub << 8) | ((uint32_t)w << 16), (uint32_t)b_terms.size(), (uint32_t)toff[f], (uint32_t)(toff[f] >> 32));
        for (uint32_t i = 0; i < r; ++i)
            for (uint32_t ts = 0; ts < sub; ++ts) {
                const uint32_t t = (ts / stj) * (stj * r) + i * stj + (ts % stj);
                b_terms.push_back(t);
                for (int u = 0; u < w; ++u) {
                    const int eu = e0 + u;
                    if (eu == e) continue;
                    b_terms.push_back(moff[eu] + (t / estride[eu]) % card[evar[eu]]);
                }
            }
    }
    while (b_terms.size() % 4) b_terms.push_back(0);
    if (a_nb.empty()) a_nb.push_back(0);

    bnpp_fg *g = new bnpp_fg();
    g->ctx = ctx;
    g->nvars = nvars;
    g->nfac = nfac;
    g->nedges = nedges;
    g->nmsg = moff[nedges];
    g->nmarg = mvoff[nvars];
    g->h_card = h_card;
    int rc;
#define UP(field, vec) if ((rc = to_device(ctx, &g->field, vec)) != BNPP_OK) { bnpp_fg_destroy(g); return rc; }
    UP(card, h_card) UP(foff, h_foff) UP(evar, evar) UP(efac, efac) UP(moff, moff) UP(voff, voff) UP(vedges, vedges)
    UP(toff, h_toff) UP(estride, estride) UP(fsize, fsize) UP(mvoff, mvoff) UP(small_edges, small_edges) UP(big_edges, big_edges)
    UP(a_hdr, a_hdr) UP(a_nb, a_nb) UP(b_hdr, b_hdr) UP(b_terms, b_terms)
    g->n_small = (int)small_edges.size();
    g->n_big = (int)big_edges.size();
#undef UP
    BNPP_CUDA(ctx, cudaMalloc(&g->ftab, sizeof(double) * (tab_total ? tab_total : 1)));
    if (tab_total)
        BNPP_CUDA(ctx, cudaMemcpyAsync(g->ftab, ftab_host, sizeof(double) * tab_total, cudaMemcpyHostToDevice, ctx->stream));
    const size_t mb = sizeof(double) * (g->nmsg ? g->nmsg : 1);
    BNPP_CUDA(ctx, cudaMalloc(&g->f2v, mb));
    BNPP_CUDA(ctx, cudaMalloc(&g->v2f, mb));
    BNPP_CUDA(ctx, cudaMalloc(&g->tmp, mb));
    BNPP_CUDA(ctx, cudaMalloc(&g->marg, sizeof(double) * (g->nmarg ? g->nmarg : 1)));
    BNPP_CUDA(ctx, cudaMalloc(&g->maxerr, sizeof(unsigned long long)));
    BNPP_CUDA(ctx, cudaMallocHost(&g->maxerr_host, sizeof(unsigned long long)));
    BNPP_CUDA(ctx, cudaMalloc(&g->err3, 4 * sizeof(unsigned long long)));
    BNPP_CUDA(ctx, cudaMallocHost(&g->sweeps_host, sizeof(uint32_t)));
    {
        int coop = 0, per_sm = 0;
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device);
        if (coop && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fg_update_kernel, 256, 0) == cudaSuccess && per_sm > 0) {
            // enough warps for one edge each where possible, never more than fits at once (a grid barrier needs all
            // blocks resident); fewer blocks make the barrier cheaper
            int want = (int)((small_edges.size() + 32 * big_edges.size() + 255) / 256);
            if (const char *env = getenv("BNPP_FG_BLOCKS")) want = atoi(env);
            const int cap = ctx->sm_count * (per_sm < 2 ? per_sm : 2);
            if (want > cap) want = cap;
            if (want < 1) want = 1;
            g->coop_blocks = want;
        }
        cudaGetLastError();
    }
    if (nedges) {
        fg_init_kernel<<<(nedges + 127) / 128, 128, 0, ctx->stream>>>(nedges, g->evar, g->card, g->moff, g->f2v, g->v2f);
        BNPP_CUDA(ctx, cudaGetLastError());
        ctx->launches++;
    }
    BNPP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // host vectors die at return
    *out = g;
    return BNPP_OK;
}

int bnpp_fg_destroy(bnpp_fg *g)
{
    if (!g) return BNPP_OK;
    cudaStreamSynchronize(g->ctx->stream);
    cudaFree(g->card); cudaFree(g->foff); cudaFree(g->evar); cudaFree(g->efac); cudaFree(g->moff);
    cudaFree(g->voff); cudaFree(g->vedges); cudaFree(g->toff); cudaFree(g->estride); cudaFree(g->fsize);
    cudaFree(g->mvoff); cudaFree(g->ftab); cudaFree(g->f2v); cudaFree(g->v2f); cudaFree(g->tmp);
    cudaFree(g->marg); cudaFree(g->maxerr); cudaFree(g->err3); cudaFree(g->small_edges); cudaFree(g->big_edges);
    cudaFree(g->a_hdr); cudaFree(g->a_nb); cudaFree(g->b_hdr); cudaFree(g->b_terms);
    if (g->maxerr_host) cudaFreeHost(g->maxerr_host);
    if (g->sweeps_host) cudaFreeHost(g->sweeps_host);
    delete g;
    return BNPP_OK;
}

static int fg_launch_sweep(bnpp_fg *g)
{
    bnpp_ctx *ctx = g->ctx;
    BNPP_CUDA(ctx, cudaMemsetAsync(g->maxerr, 0, sizeof(unsigned long long), ctx->stream));
    if (g->nedges) {
        fg_var_to_fac_kernel<<<(g->nedges + 127) / 128, 128, 0, ctx->stream>>>(g->nedges, g->a_hdr, g->a_nb, g->moff, g->f2v, g->v2f,
                                                                              g->maxerr);
        BNPP_CUDA(ctx, cudaGetLastError());
        fg_fac_to_var_kernel<<<(g->n_small + 32 * g->n_big + 127) / 128, 128, 0, ctx->stream>>>(
            g->n_small, g->small_edges, g->n_big, g->big_edges, g->b_hdr, g->b_terms, g->efac, g->evar, g->foff, g->card, g->moff, g->toff, g->estride, g->fsize, g->ftab, g->v2f,
            g->f2v, g->tmp, g->maxerr);
        BNPP_CUDA(ctx, cudaGetLastError());
        ctx->launches += 2;
    }
    BNPP_CUDA(ctx, cudaMemcpyAsync(g->maxerr_host, g->maxerr, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    return BNPP_OK;
}

int bnpp_fg_sweep(bnpp_fg *g, double *maxerror_host)
{
    if (!g) return BNPP_EINVAL;
    int rc = fg_launch_sweep(g);
    if (rc != BNPP_OK) return rc;
    BNPP_CUDA(g->ctx, cudaStreamSynchronize(g->ctx->stream));
    if (maxerror_host) {
        double d;
        memcpy(&d, g->maxerr_host, sizeof d);
        *maxerror_host = d;
    }
    return BNPP_OK;
}

int bnpp_fg_update(bnpp_fg *g, uint32_t max_sweeps, double epsilon, uint32_t *sweeps)
{
    if (!g) return BNPP_EINVAL;
    bnpp_ctx *ctx = g->ctx;
    if (g->coop_blocks > 0 && g->nedges > 0 && max_sweeps > 0) {
        // the whole loop in one cooperative launch; the host reads the sweep count once
        FgArgs a;
        a.nedges = g->nedges;
        a.evar = g->evar; a.card = g->card; a.moff = g->moff; a.voff = g->voff; a.vedges = g->vedges;
        a.efac = g->efac; a.foff = g->foff; a.toff = g->toff; a.estride = g->estride; a.fsize = g->fsize;
        a.ftab = g->ftab; a.f2v = g->f2v; a.v2f = g->v2f; a.tmp = g->tmp; a.err3 = g->err3;
        a.n_small = g->n_small; a.n_big = g->n_big; a.small_edges = g->small_edges; a.big_edges = g->big_edges;
        a.a_hdr = g->a_hdr; a.a_nb = g->a_nb; a.b_hdr = g->b_hdr; a.b_terms = g->b_terms;
        a.max_sweeps = max_sweeps;
        a.epsilon = epsilon;
        BNPP_CUDA(ctx, cudaMemsetAsync(g->err3, 0, 4 * sizeof(unsigned long long), ctx->stream));
        void *args[1] = {&a};
        cudaError_t e = cudaLaunchCooperativeKernel(reinterpret_cast<const void *>(fg_update_kernel), dim3(g->coop_blocks), dim3(256),
                                                    args, 0, ctx->stream);
        if (e == cudaSuccess) {
            ctx->launches++;
            BNPP_CUDA(ctx, cudaMemcpyAsync(g->sweeps_host, g->err3 + 3, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
            BNPP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            if (sweeps) *sweeps = *g->sweeps_host;
            return BNPP_OK;
        }
        cudaGetLastError();
        g->coop_blocks = 0;         // not launchable here (e.g. under a capture): a launch per phase from now on
    }
    uint32_t it;
    for (it = 0; it < max_sweeps; ++it) {
        double maxerror = 0.0;
        int rc = bnpp_fg_sweep(g, &maxerror);
        if (rc != BNPP_OK) return rc;
        if (maxerror < epsilon) break;   // code/graph.cpp:328
    }
    if (sweeps) *sweeps = it;
    return BNPP_OK;
}

// messages back to the uniform start of code/graph.cpp:261-274 (a second update() on the same graph)
int bnpp_fg_reset(bnpp_fg *g)
{
    if (!g) return BNPP_EINVAL;
    bnpp_ctx *ctx = g->ctx;
    if (g->nedges) {
        fg_init_kernel<<<(g->nedges + 127) / 128, 128, 0, ctx->stream>>>(g->nedges, g->evar, g->card, g->moff, g->f2v, g->v2f);
        BNPP_CUDA(ctx, cudaGetLastError());
        ctx->launches++;
    }
    return BNPP_OK;
}

int bnpp_fg_marginals(bnpp_fg *g, double *out_host)
{
    if (!g || !out_host) return BNPP_EINVAL;
    bnpp_ctx *ctx = g->ctx;
    if (g->nvars) {
        fg_marginal_kernel<<<(g->nvars + 127) / 128, 128, 0, ctx->stream>>>(g->nvars, g->card, g->moff, g->voff,
                                                                            g->vedges, g->f2v, g->mvoff, g->marg);
        BNPP_CUDA(ctx, cudaGetLastError());
        ctx->launches++;
    }
    BNPP_CUDA(ctx, cudaMemcpyAsync(out_host, g->marg, sizeof(double) * g->nmarg, cudaMemcpyDeviceToHost, ctx->stream));
    BNPP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return BNPP_OK;
}

}  // extern "C"
