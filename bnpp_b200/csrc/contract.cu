// K1/K2/K3/K4 -- the fused elimination-step kernel and its host-side plan.
//
//   out[o] = sum_{x < card(X)}  prod_k  F_k[ pi_k(o, x) ]        Z = sum_o out[o]
//
// replaces `prod *= *pf` over a bucket followed by `prod.sum_out(var)`
// (reference code/model.cpp:414-418, code/factor.cpp:117-147 and :182-212) without
// materialising the product table.  Factor::product / divide / sum_out / conditioning
// are the k=2 / k=1 special cases of the same kernel.
//
// Where the reference walks an odometer and does two hash lookups per scope variable
// per entry (code/domain.cpp:113-123,162-179), the plan below turns every operand
// into a stride vector over the OUTPUT's axes (stride 0 = axis absent = broadcast,
// SURVEY A.1), merges axes that are contiguous in every operand, and the kernel
// recovers the mixed-radix digits of its linear index with multiply-high divisions.
// Each thread owns a [V x C] micro-tile: V (1 or 2) consecutive output entries times the
// C values of the eliminated variable, loaded with 128/256-bit accesses whenever the
// operand's layout makes them contiguous.  HBM-bound: algorithmic bytes per launch =
// 8 * (sum_k #F_k + #out)   (SURVEY §8d).
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace bnpp {

enum LoadClass : uint8_t {
    LC_SCALAR = 0,  // V*C independent 8-byte loads (duplicates skipped when a stride is 0)
    LC_BCAST,       // one value for the whole micro-tile
    LC_VX,          // x contiguous (stride 1, C == 2): one 16-byte load per output entry
    LC_VX_B,        //   ... and the operand does not depend on the innermost output axis
    LC_VL,          // innermost output axis contiguous (V == 2): one 16-byte load per x
    LC_VL_B,        //   ... and the operand does not depend on x
    LC_V4,          // [j][x] contiguous: one 32-byte load, element 2*j + x
    LC_V4T,         // [x][j] contiguous: one 32-byte load, element 2*x + j
};

struct ContractParams {
    const double *in[kMaxK];
    double *out;
    double *partials;
    unsigned int *ticket;
    double *z;
    unsigned int *status;
    uint64_t n_items;           // output entries / V
    uint32_t R;                 // iteration axes, outermost first
    uint32_t cx;                // cardinality of the eliminated variable (1 = none)
    FastDiv div[kMaxR];
    uint32_t so[kMaxR];         // output stride per axis (elements, per item on the last axis)
    uint32_t s[kMaxK][kMaxR];   // operand stride per axis
    uint32_t sx[kMaxK];         // operand stride of the eliminated variable
    uint32_t sl[kMaxK];         // operand stride between the V entries of an item
    uint8_t cls[kMaxK];
    uint8_t out_vec;            // 16-byte store allowed
};

template <int K>
__device__ __forceinline__ void decompose(const ContractParams &p, uint32_t item, uint32_t (&off)[K], uint32_t &ooff)
{
#pragma unroll
    for (int k = 0; k < K; ++k) off[k] = 0;
    ooff = 0;
    uint32_t rem = item;
    for (int a = (int)p.R - 1; a > 0; --a) {
        const uint32_t q = fastdiv(rem, p.div[a]);
        const uint32_t d = rem - q * p.div[a].d;
        rem = q;
#pragma unroll
        for (int k = 0; k < K; ++k) off[k] += d * p.s[k][a];
        ooff += d * p.so[a];
    }
    if (p.R > 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) off[k] += rem * p.s[k][0];
        ooff += rem * p.so[0];
    }
}

template <int C, int V>
__device__ __forceinline__ void load_tile(const double *__restrict__ base, uint32_t off, uint32_t sx, uint32_t sl,
                                          uint8_t cls, double (&t)[V][C])
{
    const double *p = base + off;
    switch (cls) {
    case LC_BCAST: {
        const double v = ld1(p);
#pragma unroll
        for (int j = 0; j < V; ++j)
#pragma unroll
            for (int x = 0; x < C; ++x) t[j][x] = v;
        break;
    }
    case LC_VX:
        if (C == 2) {
#pragma unroll
            for (int j = 0; j < V; ++j) {
                const double2 v = ld2(p + j * sl);
                t[j][0] = v.x;
                t[j][C - 1] = v.y;
            }
        }
        break;
    case LC_VX_B:
        if (C == 2) {
            const double2 v = ld2(p);
#pragma unroll
            for (int j = 0; j < V; ++j) {
                t[j][0] = v.x;
                t[j][C - 1] = v.y;
            }
        }
        break;
    case LC_VL:
        if (V == 2) {
#pragma unroll
            for (int x = 0; x < C; ++x) {
                const double2 v = ld2(p + x * sx);
                t[0][x] = v.x;
                t[V - 1][x] = v.y;
            }
        }
        break;
    case LC_VL_B:
        if (V == 2) {
            const double2 v = ld2(p);
#pragma unroll
            for (int x = 0; x < C; ++x) {
                t[0][x] = v.x;
                t[V - 1][x] = v.y;
            }
        }
        break;
    case LC_V4:
        if (V == 2 && C == 2) {
            const double4_t v = ld4(p);
            t[0][0] = v.x; t[0][C - 1] = v.y; t[V - 1][0] = v.z; t[V - 1][C - 1] = v.w;
        }
        break;
    case LC_V4T:
        if (V == 2 && C == 2) {
            const double4_t v = ld4(p);
            t[0][0] = v.x; t[V - 1][0] = v.y; t[0][C - 1] = v.z; t[V - 1][C - 1] = v.w;
        }
        break;
    default: {
#pragma unroll
        for (int j = 0; j < V; ++j)
#pragma unroll
            for (int x = 0; x < C; ++x) {
                if (x > 0 && sx == 0) t[j][x] = t[j][0];
                else if (j > 0 && sl == 0) t[j][x] = t[0][x];
                else t[j][x] = ld1(p + j * sl + x * sx);
            }
    }
    }
}

// Fast path: eliminated variable binary (C = 2) or absent (C = 1).
template <int K, int C, int V, bool DIV>
__global__ void __launch_bounds__(kBlock) contract_fast(const __grid_constant__ ContractParams p)
{
    const uint64_t step = (uint64_t)gridDim.x * kBlock;
    double zacc = 0.0;
    bool zero_div = false;
    for (uint64_t it = (uint64_t)blockIdx.x * kBlock + threadIdx.x; it < p.n_items; it += step) {
        uint32_t off[K], ooff;
        decompose<K>(p, (uint32_t)it, off, ooff);
        double acc[V][C];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            double t[V][C];
            load_tile<C, V>(p.in[k], off[k], p.sx[k], p.sl[k], p.cls[k], t);
#pragma unroll
            for (int j = 0; j < V; ++j)
#pragma unroll
                for (int x = 0; x < C; ++x) {
                    if (k == 0) acc[j][x] = t[j][x];
                    else if (DIV) { zero_div |= (t[j][x] == 0.0); acc[j][x] = acc[j][x] / t[j][x]; }
                    else acc[j][x] = acc[j][x] * t[j][x];
                }
        }
        double r[V];
#pragma unroll
        for (int j = 0; j < V; ++j) {
            r[j] = acc[j][0];
#pragma unroll
            for (int x = 1; x < C; ++x) r[j] += acc[j][x];
            zacc += r[j];
        }
        double *o = p.out + ooff;
        if (V == 2) {
            if (p.out_vec) *reinterpret_cast<double2 *>(o) = make_double2(r[0], r[V - 1]);
            else { o[0] = r[0]; o[1] = r[V - 1]; }
        } else {
            o[0] = r[0];
        }
    }
    if (DIV && zero_div) atomicOr(p.status, BNPP_STATUS_ZERO_DIVISOR);
    grid_sum_to(zacc, p.partials, p.ticket, p.z);
}

// Generic path: any cardinality of the eliminated variable, one output entry per thread.
template <int K, bool DIV>
__global__ void __launch_bounds__(kBlock) contract_generic(const __grid_constant__ ContractParams p)
{
    const uint64_t step = (uint64_t)gridDim.x * kBlock;
    double zacc = 0.0;
    bool zero_div = false;
    for (uint64_t it = (uint64_t)blockIdx.x * kBlock + threadIdx.x; it < p.n_items; it += step) {
        uint32_t off[K], ooff;
        decompose<K>(p, (uint32_t)it, off, ooff);
        double acc = 0.0;
        for (uint32_t x = 0; x < p.cx; ++x) {
            double v = ld1(p.in[0] + off[0] + x * p.sx[0]);
#pragma unroll
            for (int k = 1; k < K; ++k) {
                const double t = ld1(p.in[k] + off[k] + x * p.sx[k]);
                if (DIV) { zero_div |= (t == 0.0); v = v / t; }
                else v = v * t;
            }
            acc += v;
        }
        p.out[ooff] = acc;
        zacc += acc;
    }
    if (DIV && zero_div) atomicOr(p.status, BNPP_STATUS_ZERO_DIVISOR);
    grid_sum_to(zacc, p.partials, p.ticket, p.z);
}

typedef void (*kernel_fn)(const ContractParams);

template <int K>
static kernel_fn pick_fast(int C, int V)
{
    if (C == 1) return V == 2 ? contract_fast<K, 1, 2, false> : contract_fast<K, 1, 1, false>;
    return V == 2 ? contract_fast<K, 2, 2, false> : contract_fast<K, 2, 1, false>;
}

static kernel_fn pick(int K, int C, int V, bool div, bool generic)
{
    if (generic) {
        if (div) return contract_generic<2, true>;
        switch (K) {
        case 1: return contract_generic<1, false>;
        case 2: return contract_generic<2, false>;
        case 3: return contract_generic<3, false>;
        case 4: return contract_generic<4, false>;
        case 5: return contract_generic<5, false>;
        default: return contract_generic<6, false>;
        }
    }
    if (div) {
        if (C == 1) return V == 2 ? contract_fast<2, 1, 2, true> : contract_fast<2, 1, 1, true>;
        return V == 2 ? contract_fast<2, 2, 2, true> : contract_fast<2, 2, 1, true>;
    }
    switch (K) {
    case 1: return pick_fast<1>(C, V);
    case 2: return pick_fast<2>(C, V);
    case 3: return pick_fast<3>(C, V);
    case 4: return pick_fast<4>(C, V);
    case 5: return pick_fast<5>(C, V);
    default: return pick_fast<6>(C, V);
    }
}

// ---------------------------------------------------------------------------
// host-side plan
// ---------------------------------------------------------------------------
struct Axis {
    uint32_t ext;
    uint64_t so;
    uint64_t s[kMaxK];
};

static bool aligned(const void *p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

int contract(bnpp_ctx *ctx, int k, const bnpp_operand *ops, const bnpp_scope *out_scope,
             int64_t elim_var, int divide, double *out_dev, double *z_dev)
{
    if (!ctx) return BNPP_EINVAL;
    if (k < 1 || k > kMaxK) return fail(ctx, BNPP_ERANK, "product_sum_out: operand count must be 1..BNPP_MAX_OPERANDS");
    if (divide && k != 2) return fail(ctx, BNPP_EINVAL, "divide needs exactly two operands");
    if (!out_scope || out_scope->rank < 0 || out_scope->rank > BNPP_MAX_RANK)
        return fail(ctx, BNPP_EINVAL, "bad output scope");
    const int wr = out_scope->rank;

    // output axes, dense row-major, last fastest (code/domain.cpp:20-24)
    std::vector<Axis> axes(wr);
    uint64_t n_out = 1;
    for (int i = wr - 1; i >= 0; --i) {
        if (out_scope->card[i] == 0) return fail(ctx, BNPP_EINVAL, "zero cardinality");
        for (int j = i + 1; j < wr; ++j)
            if (out_scope->var_id[i] == out_scope->var_id[j]) return fail(ctx, BNPP_EINVAL, "duplicate variable in output scope");
        axes[i].ext = out_scope->card[i];
        axes[i].so = n_out;
        for (int q = 0; q < kMaxK; ++q) axes[i].s[q] = 0;
        n_out *= out_scope->card[i];
        if (n_out >= (1ull << 32)) return fail(ctx, BNPP_ETOOBIG, "output table has >= 2^32 entries");
    }

    uint64_t sx[kMaxK] = {0};
    uint32_t cx = 1;
    for (int q = 0; q < k; ++q) {
        const bnpp_operand &op = ops[q];
        if (op.scope.rank < 0 || op.scope.rank > BNPP_MAX_RANK) return fail(ctx, BNPP_EINVAL, "bad operand scope");
        uint64_t dense = 1, maxoff = 0;
        for (int i = op.scope.rank - 1; i >= 0; --i) {
            const uint32_t var = op.scope.var_id[i], card = op.scope.card[i];
            const uint64_t st = op.stride ? (uint64_t)op.stride[i] : dense;
            if (op.stride && op.stride[i] < 0) return fail(ctx, BNPP_EINVAL, "negative stride");
            dense *= card;
            maxoff += (uint64_t)(card - 1) * st;
            for (int j = i + 1; j < op.scope.rank; ++j)
                if (op.scope.var_id[j] == var) return fail(ctx, BNPP_EINVAL, "duplicate variable in operand scope");
            if (elim_var >= 0 && var == (uint64_t)elim_var) {
                sx[q] = st;
                if (cx != 1 && cx != card) return fail(ctx, BNPP_EINVAL, "inconsistent cardinality of the eliminated variable");
                cx = card;
                continue;
            }
            int pos = -1;
            for (int j = 0; j < wr; ++j)
                if (out_scope->var_id[j] == var) { pos = j; break; }
            if (pos < 0) return fail(ctx, BNPP_EINVAL, "operand variable is neither in the output scope nor eliminated");
            if (out_scope->card[pos] != card) return fail(ctx, BNPP_EINVAL, "cardinality mismatch between operand and output");
            axes[pos].s[q] = st;
        }
        if (maxoff >= (1ull << 32)) return fail(ctx, BNPP_ETOOBIG, "operand table has >= 2^32 entries");
    }

    // drop size-1 axes, then merge neighbours that are contiguous in the output and in every operand
    std::vector<Axis> m;
    for (int i = 0; i < wr; ++i) {
        if (axes[i].ext == 1) continue;
        if (!m.empty()) {
            Axis &o = m.back();   // o is OUTER to axes[i]
            bool ok = (o.so == axes[i].so * axes[i].ext);
            for (int q = 0; q < k && ok; ++q) ok = (o.s[q] == axes[i].s[q] * axes[i].ext);
            if (ok && (uint64_t)o.ext * axes[i].ext < (1ull << 32)) {
                o.ext *= axes[i].ext;
                o.so = axes[i].so;
                for (int q = 0; q < k; ++q) o.s[q] = axes[i].s[q];
                continue;
            }
        }
        m.push_back(axes[i]);
    }
    if ((int)m.size() > kMaxR) return fail(ctx, BNPP_ERANK, "more than BNPP_MAX_AXES non-mergeable axes");

    ContractParams p;
    memset(&p, 0, sizeof p);
    const bool generic = (cx > 2);
    const int C = generic ? 0 : (int)cx;
    int V = 1;
    if (!generic && !m.empty() && (m.back().ext % 2 == 0)) V = 2;

    for (int q = 0; q < k; ++q) {
        p.in[q] = ops[q].data;
        p.sx[q] = (uint32_t)sx[q];
        p.sl[q] = m.empty() ? 0 : (uint32_t)m.back().s[q];
    }
    const uint32_t sol = m.empty() ? 0 : (uint32_t)m.back().so;
    if (V == 2) {
        Axis &l = m.back();
        l.ext /= 2;
        l.so *= 2;
        for (int q = 0; q < k; ++q) l.s[q] *= 2;
        if (l.ext == 1) m.pop_back();
    }
    p.R = (uint32_t)m.size();
    for (uint32_t a = 0; a < p.R; ++a) {
        p.div[a] = make_fastdiv(m[a].ext);
        p.so[a] = (uint32_t)m[a].so;
        for (int q = 0; q < k; ++q) p.s[q][a] = (uint32_t)m[a].s[q];
    }
    p.n_items = n_out / V;
    p.cx = cx;
    p.out = out_dev;
    p.out_vec = (V == 2 && sol == 1 && aligned(out_dev, 16)) ? 1 : 0;

    // load class per operand: how the [V x C] micro-tile sits in the operand's memory
    for (int q = 0; q < k && !generic; ++q) {
        const uint32_t x = p.sx[q], l = (V == 2) ? p.sl[q] : 0;
        uint32_t g = 0;   // gcd-like: every item offset is a multiple of 2 / 4 elements?
        bool mult2 = true, mult4 = true;
        for (uint32_t a = 0; a < p.R; ++a) {
            if (p.s[q][a] % 2) mult2 = false;
            if (p.s[q][a] % 4) mult4 = false;
        }
        (void)g;
        const bool has_x = (C == 2 && x != 0), has_l = (V == 2 && l != 0);
        uint8_t c = LC_SCALAR;
        if (!has_x && !has_l) c = LC_BCAST;
        else if (has_x && x == 1 && mult2 && aligned(p.in[q], 16) && (!has_l || l % 2 == 0)) {
            if (!has_l) c = LC_VX_B;
            else if (l == 2 && mult4 && aligned(p.in[q], 32)) c = LC_V4;
            else c = LC_VX;
        } else if (has_l && l == 1 && mult2 && aligned(p.in[q], 16) && (!has_x || x % 2 == 0)) {
            if (!has_x) c = LC_VL_B;
            else if (x == 2 && mult4 && aligned(p.in[q], 32)) c = LC_V4T;
            else c = LC_VL;
        }
        p.cls[q] = c;
    }

    p.partials = ctx->partials;
    p.ticket = ctx->ticket;
    p.z = z_dev ? z_dev : ctx->scratch_z;
    p.status = ctx->status;

    uint64_t blocks = (p.n_items + kBlock - 1) / kBlock;
    const uint64_t cap = (uint64_t)ctx->sm_count * 8;   // persistent grid: 8 CTAs of 256 threads per SM
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;

    kernel_fn fn = pick(k, C, V, divide != 0, generic);
    fn<<<(unsigned)blocks, kBlock, 0, ctx->stream>>>(p);
    BNPP_CUDA(ctx, cudaGetLastError());
    ctx->launches++;
    char nm[96];
    if (generic) snprintf(nm, sizeof nm, "contract_generic<K=%d,div=%d> cx=%u R=%u", k, divide != 0, cx, p.R);
    else snprintf(nm, sizeof nm, "contract_fast<K=%d,C=%d,V=%d,div=%d> R=%u cls=%d,%d,%d", k, C, V, divide != 0, p.R,
                  p.cls[0], k > 1 ? p.cls[1] : -1, k > 2 ? p.cls[2] : -1);
    ctx->last_kernel = nm;
    ctx->last_grid = (uint32_t)blocks;
    ctx->last_block = kBlock;
    return BNPP_OK;
}

}  // namespace bnpp
