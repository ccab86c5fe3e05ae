// K7 -- factor-graph sum-product (loopy BP) with every message of a phase updated in
// one launch.  Replaces FactorGraph (reference code/graph.cpp:256-403).
//
// The reference sweeps edge by edge; within a phase each update reads only the other
// direction's messages (code/graph.cpp:340-359, 367-388), so a flood over all edges
// is the same schedule and the converging sweep index is preserved.
// Messages live in two flat device arrays indexed by edge e = foff[f] + slot; the
// host never touches them between sweeps except for the one max-error word.
#include <vector>

#include "common.cuh"

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

// variable -> factor, code/graph.cpp:334-362: m_{v->f} = normalize(prod_{g in N(v)\f} m_{g->v})
__global__ void __launch_bounds__(128) fg_var_to_fac_kernel(int nedges, const uint32_t *__restrict__ evar,
                                                            const uint32_t *__restrict__ card,
                                                            const uint32_t *__restrict__ moff,
                                                            const int32_t *__restrict__ voff,
                                                            const int32_t *__restrict__ vedges,
                                                            const double *__restrict__ f2v, double *__restrict__ v2f,
                                                            unsigned long long *maxerr)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nedges) return;
    const uint32_t v = evar[e], r = card[v];
    const int b = voff[v], n = voff[v + 1];
    double z = 0.0;
    for (uint32_t i = 0; i < r; ++i) {
        double p = 1.0;
        for (int q = b; q < n; ++q) {
            const int e2 = vedges[q];
            if (e2 != e) p *= f2v[moff[e2] + i];
        }
        z += p;
    }
    double worst = 0.0;
    for (uint32_t i = 0; i < r; ++i) {
        double p = 1.0;
        for (int q = b; q < n; ++q) {
            const int e2 = vedges[q];
            if (e2 != e) p *= f2v[moff[e2] + i];
        }
        const double nv = p / z;
        const double ov = v2f[moff[e] + i];
        const double err = fabs(ov - nv) / ov;
        if (err > worst) worst = err;
        v2f[moff[e] + i] = nv;
    }
    note_error(maxerr, worst);
}

// factor -> variable, code/graph.cpp:364-391: one warp per edge (f, slot j):
// m_{f->v}[i] = sum over the factor entries with digit_j = i of  F * prod_{u != j} m_{u->f}
__global__ void __launch_bounds__(128) fg_fac_to_var_kernel(int nedges, const int32_t *__restrict__ efac,
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
    const int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (e >= nedges) return;
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
                p *= v2f[moff[eu] + d];
            }
            part += p;
        }
        part = warp_sum(part);
        if (lane == 0) tmp[moff[e] + i] = part;
    }
    __syncwarp();
    double z = 0.0;
    for (uint32_t i = 0; i < r; ++i) z += tmp[moff[e] + i];   // same order in every lane
    double worst = 0.0;
    for (uint32_t i = lane; i < r; i += 32) {
        const double nv = tmp[moff[e] + i] / z;
        const double ov = f2v[moff[e] + i];
        const double err = fabs(ov - nv) / ov;
        if (err > worst) worst = err;
        f2v[moff[e] + i] = nv;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double other = __shfl_down_sync(0xffffffffu, worst, o);
        if (other > worst) worst = other;
    }
    if (lane == 0) note_error(maxerr, worst);
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
    UP(toff, h_toff) UP(estride, estride) UP(fsize, fsize) UP(mvoff, mvoff)
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
    cudaFree(g->marg); cudaFree(g->maxerr);
    if (g->maxerr_host) cudaFreeHost(g->maxerr_host);
    delete g;
    return BNPP_OK;
}

static int fg_launch_sweep(bnpp_fg *g)
{
    bnpp_ctx *ctx = g->ctx;
    BNPP_CUDA(ctx, cudaMemsetAsync(g->maxerr, 0, sizeof(unsigned long long), ctx->stream));
    if (g->nedges) {
        fg_var_to_fac_kernel<<<(g->nedges + 127) / 128, 128, 0, ctx->stream>>>(
            g->nedges, g->evar, g->card, g->moff, g->voff, g->vedges, g->f2v, g->v2f, g->maxerr);
        BNPP_CUDA(ctx, cudaGetLastError());
        fg_fac_to_var_kernel<<<(g->nedges * 32 + 127) / 128, 128, 0, ctx->stream>>>(
            g->nedges, g->efac, g->evar, g->foff, g->card, g->moff, g->toff, g->estride, g->fsize, g->ftab, g->v2f,
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
