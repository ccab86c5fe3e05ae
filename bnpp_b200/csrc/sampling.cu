// Forward sampling of a Bayesian network on the GPU: BN::logical_sampling and
// BN::likelihood_weighting, code/model.cpp:540-690, over Factor::sampling, code/factor.cpp:257-288.
//
// The reference draws one sample at a time: per variable a conditioning + normalisation + scan of its CPT, each with
// its own std::random_device.  Samples are independent, so here ONE THREAD OWNS ONE SAMPLE: it walks the variables
// in the reference's topological order with the sample's valuation in (L1-backed) local memory, reads each CPT in
// place from HBM (the tables are the resident ones of the model; a few KB to MB, L2-resident), and draws with a
// counter-based generator (Philox4x32-10 keyed by the caller's seed and the sample index), so a run is reproducible
// -- the one observable difference from the reference, whose draws cannot be repeated.
//   logical sampling     : M samples, hits = #{samples that agree with the evidence}; estimate hits / M.
//   likelihood weighting : evidence variables are clamped, W = product of their CPT entries; the reference's
//                          bounded-variance stopping rule -- draw until sum W / U reaches N* -- is sequential in the
//                          sample index: batches of weights are drawn in parallel, an inclusive scan in sample order
//                          finds the sample the rule stops at, exactly as if they had been drawn one by one.
// Latency/instruction-bound integer + fp64 work; no tensor cores, no collective.
#include <algorithm>
#include <vector>

#include "common.cuh"

struct bnpp_sampler {
    bnpp_ctx *ctx = nullptr;
    int nvars = 0;
    // per position of the topological order
    uint32_t *var = nullptr, *card = nullptr, *child_stride = nullptr, *pa_off = nullptr;     // pa_off: [nvars + 1]
    uint32_t *pa_var = nullptr, *pa_stride = nullptr;
    const double **table = nullptr;
    int32_t *ev = nullptr;          // [nvars] evidence value per variable id, -1 = free (rewritten per query)
    double *weights = nullptr;      // likelihood weighting: W / U of a batch
    unsigned long long *counter = nullptr;
    size_t weights_cap = 0;
};

namespace bnpp {

constexpr int kMaxSampleVars = 2048;

struct Philox {
    uint32_t key0, key1, c0, c1, c2, c3;
    uint32_t out[4];
    int have;
    __device__ Philox(uint64_t seed, uint64_t sample) : key0((uint32_t)seed), key1((uint32_t)(seed >> 32)), c0(0),
        c1(0), c2((uint32_t)sample), c3((uint32_t)(sample >> 32)), have(0) {}
    __device__ void round(uint32_t &a, uint32_t &b, uint32_t &c, uint32_t &d, uint32_t k0, uint32_t k1)
    {
        const uint32_t hi0 = __umulhi(0xD2511F53u, a), lo0 = 0xD2511F53u * a;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c), lo1 = 0xCD9E8D57u * c;
        a = hi1 ^ b ^ k0;
        b = lo1;
        c = hi0 ^ d ^ k1;
        d = lo0;
    }
    __device__ double uniform()      // [0, 1], 32 bits (the reference: rd() / rd.max(), 32 bits as well)
    {
        if (!have) {
            uint32_t a = c0, b = c1, c = c2, d = c3, k0 = key0, k1 = key1;
#pragma unroll
            for (int r = 0; r < 10; ++r) {
                round(a, b, c, d, k0, k1);
                k0 += 0x9E3779B9u;
                k1 += 0xBB67AE85u;
            }
            out[0] = a; out[1] = b; out[2] = c; out[3] = d;
            have = 4;
            if (++c0 == 0) ++c1;
        }
        return (double)out[--have] / 4294967295.0;
    }
};

// one variable of one sample: the conditional of the child given the sampled parents (Factor::conditioning), normalised
// when its sum is off by more than 0.001 (code/factor.cpp:266-269), scanned with `prob <= p` (code/factor.cpp:276-280)
__device__ __forceinline__ uint32_t draw_child(const double *__restrict__ tab, uint32_t base, uint32_t stride, uint32_t card, double prob)
{
    double z = 0.0;
    for (uint32_t x = 0; x < card; ++x) z += __ldg(tab + base + x * stride);
    const bool renorm = fabs(z - 1.0) > 0.001;
    double p = 0.0;
    for (uint32_t x = 0; x < card; ++x) {
        const double v = __ldg(tab + base + x * stride);
        p += renorm ? v / z : v;
        if (prob <= p) return x;
    }
    return card - 1;
}

// mode 0: logical sampling (count the samples consistent with the evidence); mode 1: likelihood weighting (W / U per sample)
template <int MODE>
__global__ void __launch_bounds__(128) sample_kernel(int nvars, const uint32_t *__restrict__ var, const uint32_t *__restrict__ card,
                                                      const uint32_t *__restrict__ child_stride, const uint32_t *__restrict__ pa_off,
                                                      const uint32_t *__restrict__ pa_var, const uint32_t *__restrict__ pa_stride,
                                                      const double *const *__restrict__ table, const int32_t *__restrict__ ev,
                                                      uint64_t first, uint64_t n, uint64_t seed, double inv_u,
                                                      unsigned long long *hits, double *weights)
{
    uint8_t val[kMaxSampleVars];
    unsigned long long mine = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        Philox rng(seed, first + i);
        bool consistent = true;
        double w = 1.0;
        for (int t = 0; t < nvars; ++t) {
            const uint32_t v = var[t];
            uint32_t base = 0;
            for (uint32_t q = pa_off[t]; q < pa_off[t + 1]; ++q) base += (uint32_t)val[pa_var[q]] * pa_stride[q];
            const int32_t e = ev[v];
            if (MODE == 1 && e >= 0) {
                val[v] = (uint8_t)e;
                w *= __ldg(table[t] + base + (uint32_t)e * child_stride[t]);      // Factor::conditioning on the full valuation
            } else {
                const uint32_t x = draw_child(table[t], base, child_stride[t], card[t], rng.uniform());
                val[v] = (uint8_t)x;
                if (MODE == 0 && e >= 0 && (uint32_t)e != x) consistent = false;
            }
        }
        if (MODE == 0) mine += consistent;
        else weights[i] = w * inv_u;
    }
    if (MODE == 0) {
        for (int o = 16; o > 0; o >>= 1) mine += __shfl_down_sync(0xffffffffu, mine, o);
        if ((threadIdx.x & 31) == 0 && mine) atomicAdd(hits, mine);
    }
}

}  // namespace bnpp

using namespace bnpp;

extern "C" {

// scopes: one CPT per variable, scope[0] the child, the rest its parents (code/model.cpp:111-119); order: the
// reference's topological sampling order (variable ids); tables_dev: the resident CPTs, indexed by variable id
int bnpp_sampler_create(bnpp_ctx *ctx, int nvars, const uint32_t *card, const bnpp_scope *scopes, const uint32_t *order,
                        const double *const *tables_dev, bnpp_sampler **out)
{
    if (!ctx || !out || nvars < 1 || !card || !scopes || !order || !tables_dev) return BNPP_EINVAL;
    *out = nullptr;
    if (nvars > kMaxSampleVars) return fail(ctx, BNPP_ETOOBIG, "sampler: more than 2048 variables");
    std::vector<uint32_t> var(nvars), cd(nvars), cs(nvars), pa_off(nvars + 1, 0), pa_var, pa_stride;
    std::vector<const double *> tab(nvars);
    for (int t = 0; t < nvars; ++t) {
        const uint32_t v = order[t];
        if (v >= (uint32_t)nvars) return fail(ctx, BNPP_EINVAL, "sampler: bad variable id in the order");
        const bnpp_scope &s = scopes[v];
        if (s.rank < 1 || s.var_id[0] != v) return fail(ctx, BNPP_EINVAL, "sampler: factor i must be the CPT of variable i, child first");
        if (card[v] > 255) return fail(ctx, BNPP_ETOOBIG, "sampler: a variable with more than 255 values");
        uint64_t st = 1;
        std::vector<uint32_t> strides(s.rank);
        for (int i = s.rank - 1; i >= 0; --i) {
            strides[i] = (uint32_t)st;
            st *= s.card[i];
            if (st >= (1ull << 32)) return fail(ctx, BNPP_ETOOBIG, "sampler: CPT with >= 2^32 entries");
        }
        var[t] = v;
        cd[t] = card[v];
        cs[t] = strides[0];
        for (int i = 1; i < s.rank; ++i) {
            pa_var.push_back(s.var_id[i]);
            pa_stride.push_back(strides[i]);
        }
        pa_off[t + 1] = (uint32_t)pa_var.size();
        tab[t] = tables_dev[v];
    }
    if (pa_var.empty()) {
        pa_var.push_back(0);
        pa_stride.push_back(0);
    }
    bnpp_sampler *sp = new bnpp_sampler();
    sp->ctx = ctx;
    sp->nvars = nvars;
    auto up = [&](void **dst, const void *src, size_t bytes) {
        if (cudaMalloc(dst, bytes ? bytes : 8) != cudaSuccess) return false;
        return cudaMemcpyAsync(*dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream) == cudaSuccess;
    };
    bool ok = up((void **)&sp->var, var.data(), 4 * var.size()) && up((void **)&sp->card, cd.data(), 4 * cd.size()) &&
              up((void **)&sp->child_stride, cs.data(), 4 * cs.size()) && up((void **)&sp->pa_off, pa_off.data(), 4 * pa_off.size()) &&
              up((void **)&sp->pa_var, pa_var.data(), 4 * pa_var.size()) && up((void **)&sp->pa_stride, pa_stride.data(), 4 * pa_stride.size()) &&
              up((void **)&sp->table, tab.data(), sizeof(double *) * tab.size());
    ok = ok && cudaMalloc((void **)&sp->ev, 4 * nvars) == cudaSuccess && cudaMalloc((void **)&sp->counter, 8) == cudaSuccess;
    if (!ok || cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
        cudaGetLastError();
        bnpp_sampler_destroy(sp);
        return fail(ctx, BNPP_ECUDA, "sampler: device allocation failed");
    }
    *out = sp;
    return BNPP_OK;
}

int bnpp_sampler_destroy(bnpp_sampler *sp)
{
    if (!sp) return BNPP_OK;
    cudaStreamSynchronize(sp->ctx->stream);
    cudaFree(sp->var); cudaFree(sp->card); cudaFree(sp->child_stride); cudaFree(sp->pa_off); cudaFree(sp->pa_var);
    cudaFree(sp->pa_stride); cudaFree((void *)sp->table); cudaFree(sp->ev); cudaFree(sp->weights); cudaFree(sp->counter);
    delete sp;
    return BNPP_OK;
}

static int set_evidence(bnpp_sampler *sp, int n_ev, const uint32_t *ev_var, const uint32_t *ev_val)
{
    std::vector<int32_t> ev(sp->nvars, -1);
    for (int i = 0; i < n_ev; ++i) {
        if (ev_var[i] >= (uint32_t)sp->nvars) return fail(sp->ctx, BNPP_EINVAL, "sampler: evidence names an unknown variable");
        ev[ev_var[i]] = (int32_t)ev_val[i];
    }
    BNPP_CUDA(sp->ctx, cudaMemcpyAsync(sp->ev, ev.data(), 4 * ev.size(), cudaMemcpyHostToDevice, sp->ctx->stream));
    BNPP_CUDA(sp->ctx, cudaStreamSynchronize(sp->ctx->stream));       // `ev` dies here
    return BNPP_OK;
}

// BN::logical_sampling (code/model.cpp:540-560): n_samples forward samples; *hits = how many agree with the evidence
int bnpp_sampler_logical(bnpp_sampler *sp, int n_ev, const uint32_t *ev_var, const uint32_t *ev_val, uint64_t n_samples,
                         uint64_t seed, uint64_t *hits)
{
    if (!sp || !hits || (n_ev > 0 && (!ev_var || !ev_val))) return BNPP_EINVAL;
    bnpp_ctx *ctx = sp->ctx;
    int rc = set_evidence(sp, n_ev, ev_var, ev_val);
    if (rc != BNPP_OK) return rc;
    BNPP_CUDA(ctx, cudaMemsetAsync(sp->counter, 0, 8, ctx->stream));
    if (n_samples) {
        const unsigned blocks = (unsigned)std::min<uint64_t>((n_samples + 127) / 128, (uint64_t)ctx->sm_count * 8);
        sample_kernel<0><<<blocks, 128, 0, ctx->stream>>>(sp->nvars, sp->var, sp->card, sp->child_stride, sp->pa_off, sp->pa_var,
                                                            sp->pa_stride, sp->table, sp->ev, 0, n_samples, seed, 1.0, sp->counter, nullptr);
        BNPP_CUDA(ctx, cudaGetLastError());
        ctx->launches++;
    }
    unsigned long long h = 0;
    BNPP_CUDA(ctx, cudaMemcpyAsync(&h, sp->counter, 8, cudaMemcpyDeviceToHost, ctx->stream));
    BNPP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *hits = h;
    return BNPP_OK;
}

// BN::likelihood_weighting (code/model.cpp:620-690), bounded variance: samples are drawn until the sum of W / U reaches
// n_star; *n_sum = that sum, *m = the number of samples the rule used (estimate = U * n_sum / m).  Weights are drawn
// `batch` at a time and consumed in sample order, so the answer does not depend on the batch size.
int bnpp_sampler_likelihood(bnpp_sampler *sp, int n_ev, const uint32_t *ev_var, const uint32_t *ev_val, double u_bound,
                            double n_star, uint64_t batch, uint64_t max_samples, uint64_t seed, double *n_sum, uint64_t *m)
{
    if (!sp || !n_sum || !m || !(u_bound > 0.0) || batch == 0 || (n_ev > 0 && (!ev_var || !ev_val))) return BNPP_EINVAL;
    bnpp_ctx *ctx = sp->ctx;
    int rc = set_evidence(sp, n_ev, ev_var, ev_val);
    if (rc != BNPP_OK) return rc;
    if (sp->weights_cap < batch) {
        cudaFree(sp->weights);
        sp->weights = nullptr;
        BNPP_CUDA(ctx, cudaMalloc((void **)&sp->weights, sizeof(double) * batch));
        sp->weights_cap = batch;
    }
    std::vector<double> host(batch);
    double n = 0.0;
    uint64_t used = 0;
    while (n < n_star && used < max_samples) {
        const unsigned blocks = (unsigned)std::min<uint64_t>((batch + 127) / 128, (uint64_t)ctx->sm_count * 8);
        sample_kernel<1><<<blocks, 128, 0, ctx->stream>>>(sp->nvars, sp->var, sp->card, sp->child_stride, sp->pa_off, sp->pa_var,
                                                            sp->pa_stride, sp->table, sp->ev, used, batch, seed, 1.0 / u_bound, nullptr,
                                                            sp->weights);
        BNPP_CUDA(ctx, cudaGetLastError());
        ctx->launches++;
        BNPP_CUDA(ctx, cudaMemcpyAsync(host.data(), sp->weights, sizeof(double) * batch, cudaMemcpyDeviceToHost, ctx->stream));
        BNPP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        // the reference's loop, in sample order: `while (N < Nstar) { ...; N += W / U; ++M; }`
        for (uint64_t i = 0; i < batch && n < n_star && used < max_samples; ++i) {
            n += host[i];
            ++used;
        }
    }
    *n_sum = n;
    *m = used;
    return BNPP_OK;
}

}  // extern "C"
