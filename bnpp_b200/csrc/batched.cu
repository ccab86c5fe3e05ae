// K8 -- batched-evidence variable elimination (BASELINE config 5: tens of thousands of
// evidence sets on one network, PR per set).
//
// Replaces running BN::partition (reference code/model.cpp:250-301) once per evidence
// set: with the observed variable IDS fixed, every set shares one elimination order and
// one plan (ve.cu); only the evidence VALUES differ.  Each intermediate gets the batch as
// its FASTEST axis, [entries][batch], so consecutive threads touch consecutive doubles;
// the resident CPTs are shared by all sets and read through a per-set base offset
//     base_k(b) = sum_j stride_kj * value[b][obs_kj]
// which is the whole of Factor::conditioning (code/factor.cpp:214-242) here.  One launch
// per bucket for the entire batch; nothing is materialised per evidence set.
#include <cstring>
#include <vector>

#include "batched.hpp"
#include "common.cuh"

namespace bnpp {

constexpr int kMaxObsPerOperand = 8;

struct BOperand {
    const double *ptr;
    uint32_t batched;                       // 1: intermediate [entries][batch]; 0: resident CPT view
    uint32_t sx;                            // stride of the eliminated variable
    uint32_t nobs;
    uint32_t ostride[kMaxObsPerOperand];    // stride of an observed axis ...
    uint32_t oidx[kMaxObsPerOperand];       // ... and its column in the evidence matrix
};

struct BParams {
    BOperand op[kMaxK];
    double *out;                            // [n_out][nb]
    const uint8_t *ev;                      // [nb][n_obs] evidence values of this launch's sets
    uint32_t n_obs, nb, n_out, cx, R;
    FastDiv nbdiv;
    FastDiv div[kMaxR];
    uint32_t s[kMaxK][kMaxR];
};

template <int K>
__global__ void __launch_bounds__(kBlock) contract_batched(const __grid_constant__ BParams p)
{
    const uint64_t total = (uint64_t)p.n_out * p.nb;
    const uint64_t step = (uint64_t)gridDim.x * kBlock;
    for (uint64_t idx = (uint64_t)blockIdx.x * kBlock + threadIdx.x; idx < total; idx += step) {
        // idx = o * nb + b with the batch fastest
        uint32_t o, b;
        if (p.nb == 1) { o = (uint32_t)idx; b = 0; }
        else if (total < (1ull << 32)) { o = fastdiv((uint32_t)idx, p.nbdiv); b = (uint32_t)idx - o * p.nb; }
        else { o = (uint32_t)(idx / p.nb); b = (uint32_t)(idx - (uint64_t)o * p.nb); }
        uint32_t off[K];
#pragma unroll
        for (int k = 0; k < K; ++k) off[k] = 0;
        uint32_t rem = o;
        for (int a = (int)p.R - 1; a > 0; --a) {
            const uint32_t q = fastdiv(rem, p.div[a]);
            const uint32_t d = rem - q * p.div[a].d;
            rem = q;
#pragma unroll
            for (int k = 0; k < K; ++k) off[k] += d * p.s[k][a];
        }
        if (p.R > 0) {
#pragma unroll
            for (int k = 0; k < K; ++k) off[k] += rem * p.s[k][0];
        }
        const uint8_t *ev = p.ev + (uint64_t)b * p.n_obs;
        const double *src[K];
        uint64_t xs[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const BOperand &op = p.op[k];
            if (op.batched) {
                src[k] = op.ptr + ((uint64_t)off[k] * p.nb + b);
                xs[k] = (uint64_t)op.sx * p.nb;
            } else {
                uint32_t base = off[k];
                for (uint32_t j = 0; j < op.nobs; ++j) base += op.ostride[j] * ev[op.oidx[j]];
                src[k] = op.ptr + base;
                xs[k] = op.sx;
            }
        }
        double acc = 0.0;
        for (uint32_t x = 0; x < p.cx; ++x) {
            double v = ld1(src[0] + x * xs[0]);
#pragma unroll
            for (int k = 1; k < K; ++k) v = __dmul_rn(v, ld1(src[k] + x * xs[k]));
            acc = (x == 0) ? v : __dadd_rn(acc, v);
        }
        p.out[idx] = acc;
    }
}

typedef void (*batched_fn)(const BParams);

static batched_fn pick_batched(int K)
{
    switch (K) {
    case 1: return contract_batched<1>;
    case 2: return contract_batched<2>;
    case 3: return contract_batched<3>;
    case 4: return contract_batched<4>;
    case 5: return contract_batched<5>;
    default: return contract_batched<6>;
    }
}

int contract_batched_step(bnpp_ctx *ctx, int k, const BatchedOperandDesc *ops, const std::vector<uint32_t> &out_var,
                          const std::vector<uint32_t> &out_card, int64_t elim, uint32_t nb, const uint8_t *ev_dev,
                          uint32_t n_obs, double *out_dev)
{
    if (k < 1 || k > kMaxK) return fail(ctx, BNPP_ERANK, "batched step: operand count");
    BParams p;
    memset(&p, 0, sizeof p);
    const int wr = (int)out_var.size();
    struct Ax { uint32_t ext; uint64_t s[kMaxK]; };
    std::vector<Ax> axes(wr);
    uint64_t n_out = 1;
    for (int i = wr - 1; i >= 0; --i) {
        axes[i].ext = out_card[i];
        for (int q = 0; q < kMaxK; ++q) axes[i].s[q] = 0;
        n_out *= out_card[i];
    }
    if (n_out >= (1ull << 32)) return fail(ctx, BNPP_ETOOBIG, "batched step: output too large");
    uint32_t cx = 1;
    for (int q = 0; q < k; ++q) {
        const BatchedOperandDesc &op = ops[q];
        uint64_t dense = 1;
        for (int i = (int)op.var->size() - 1; i >= 0; --i) {
            const uint64_t st = op.stride->empty() ? dense : (uint64_t)(*op.stride)[i];
            dense *= (*op.card)[i];
            if (elim >= 0 && (*op.var)[i] == (uint64_t)elim) {
                p.op[q].sx = (uint32_t)st;
                cx = (*op.card)[i];
                continue;
            }
            int pos = -1;
            for (int j = 0; j < wr; ++j)
                if (out_var[j] == (*op.var)[i]) { pos = j; break; }
            if (pos < 0) return fail(ctx, BNPP_EINVAL, "batched step: operand variable not in the output");
            axes[pos].s[q] = st;
        }
        p.op[q].ptr = op.ptr;
        p.op[q].batched = op.batched ? 1 : 0;
        if (op.obs->size() > (size_t)kMaxObsPerOperand) return fail(ctx, BNPP_ERANK, "batched step: too many observed axes in one table");
        p.op[q].nobs = (uint32_t)op.obs->size();
        for (size_t j = 0; j < op.obs->size(); ++j) {
            p.op[q].ostride[j] = (uint32_t)(*op.obs)[j].first;
            p.op[q].oidx[j] = (uint32_t)(*op.obs)[j].second;
        }
    }
    // merge contiguous axes (output dense, so only the operands decide)
    std::vector<Ax> m;
    for (int i = 0; i < wr; ++i) {
        if (axes[i].ext == 1) continue;
        if (!m.empty()) {
            Ax &o = m.back();
            bool ok = true;
            for (int q = 0; q < k && ok; ++q) ok = (o.s[q] == axes[i].s[q] * axes[i].ext);
            if (ok) {
                o.ext *= axes[i].ext;
                for (int q = 0; q < k; ++q) o.s[q] = axes[i].s[q];
                continue;
            }
        }
        m.push_back(axes[i]);
    }
    if ((int)m.size() > kMaxR) return fail(ctx, BNPP_ERANK, "batched step: too many axes");
    p.R = (uint32_t)m.size();
    for (uint32_t a = 0; a < p.R; ++a) {
        p.div[a] = make_fastdiv(m[a].ext);
        for (int q = 0; q < k; ++q) p.s[q][a] = (uint32_t)m[a].s[q];
    }
    p.out = out_dev;
    p.ev = ev_dev;
    p.n_obs = n_obs;
    p.nb = nb;
    p.nbdiv = make_fastdiv(nb > 1 ? nb : 2);
    p.n_out = (uint32_t)n_out;
    p.cx = cx;
    const uint64_t total = n_out * nb;
    uint64_t blocks = (total + kBlock - 1) / kBlock;
    const uint64_t cap = (uint64_t)ctx->sm_count * 8;
    if (blocks > cap) blocks = cap;
    pick_batched(k)<<<(unsigned)blocks, kBlock, 0, ctx->stream>>>(p);
    BNPP_CUDA(ctx, cudaGetLastError());
    ctx->launches++;
    return BNPP_OK;
}

}  // namespace bnpp
