/* TEST INFRASTRUCTURE ONLY -- CPU restatement of bn-pp's factor algebra.
 *
 * Plain-C restatement of the reference algorithm for the hot path, used as the
 * checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.
 * The shipped product never links or calls this file.
 *
 * Parity is PINNED: tests/test_oracle.py checks every function here against
 * outputs of the unmodified reference (oracle/_ref/ref_harness, fixtures in
 * the JSON fixtures under tests/golden/ made by oracle/make_golden.py) and against the
 * reference's own shipped goldens (grid3x3.uai.PR/.MAR, network.uai.PR/.MAR).
 *
 * Conventions follow the reference literally:
 *   - tables are row-major with the LAST scope variable fastest
 *     (code/domain.cpp:20-24);
 *   - every loop enumerates valuations with the odometer of
 *     code/domain.cpp:113-123 and maps them with the "consistent valuation"
 *     rule of code/domain.cpp:162-190 (an axis absent from the other domain
 *     contributes stride 0);
 *   - partitions are sequential left-to-right sums in index order
 *     (code/factor.cpp:139,172,204,234).
 * A scope is an array of variable ids; card[] is indexed by variable id.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_MAXW 64

/* code/domain.cpp:15-26 -- strides ("_offset"), returns the table size */
static uint64_t orc_strides(int w, const unsigned *ids, const unsigned *card, uint64_t *stride)
{
    uint64_t size = 1;
    for (int i = w - 1; i >= 0; --i) {
        stride[i] = size;
        size *= card[ids[i]];
    }
    return size;
}

uint64_t orc_domain_size(int w, const unsigned *ids, const unsigned *card)
{
    uint64_t s[ORC_MAXW];
    return orc_strides(w, ids, card, s);
}

static int orc_find(int w, const unsigned *ids, unsigned id)
{
    for (int i = 0; i < w; ++i)
        if (ids[i] == id) return i;
    return -1;
}

/* code/domain.cpp:32-52 -- union scope: d1 in order, then d2-only variables in d2 order */
int orc_union_scope(int wa, const unsigned *ida, int wb, const unsigned *idb, unsigned *out)
{
    int w = 0;
    for (int i = 0; i < wa; ++i) out[w++] = ida[i];
    for (int i = 0; i < wb; ++i)
        if (orc_find(wa, ida, idb[i]) < 0) out[w++] = idb[i];
    return w;
}

/* code/domain.cpp:113-123 */
static void orc_next(int w, const unsigned *ids, const unsigned *card, unsigned *val)
{
    int j;
    for (j = w - 1; j >= 0 && val[j] == card[ids[j]] - 1; --j) val[j] = 0;
    if (j >= 0) val[j]++;
}

/* code/domain.cpp:162-179 -- index into (w,ids) of the valuation `val` given in (wd,idd) order */
static uint64_t orc_pos_consistent(int w, const unsigned *ids, const uint64_t *stride,
                                   int wd, const unsigned *idd, const unsigned *val)
{
    uint64_t pos = 0;
    for (int i = 0; i < w; ++i) {
        int j = orc_find(wd, idd, ids[i]);
        if (j >= 0) pos += stride[i] * val[j];
    }
    return pos;
}

/* code/factor.cpp:117-147 (product) and :149-180 (divide).
 * out scope must be orc_union_scope(a, b). Returns 0, or -1 on a zero divisor
 * (the reference asserts, code/factor.cpp:169). */
int orc_product(int wa, const unsigned *ida, const double *va,
                int wb, const unsigned *idb, const double *vb,
                const unsigned *card, int divide,
                double *out, double *partition)
{
    unsigned idu[ORC_MAXW], val[ORC_MAXW];
    uint64_t sa[ORC_MAXW], sb[ORC_MAXW], su[ORC_MAXW];
    int wu = orc_union_scope(wa, ida, wb, idb, idu);
    uint64_t size = orc_strides(wu, idu, card, su);
    orc_strides(wa, ida, card, sa);
    orc_strides(wb, idb, card, sb);
    memset(val, 0, sizeof val);
    double z = 0;
    for (uint64_t i = 0; i < size; ++i) {
        uint64_t p1 = orc_pos_consistent(wa, ida, sa, wu, idu, val);
        uint64_t p2 = orc_pos_consistent(wb, idb, sb, wu, idu, val);
        double v;
        if (divide) {
            if (vb[p2] == 0) return -1;
            v = va[p1] / vb[p2];
        } else {
            v = va[p1] * vb[p2];
        }
        out[i] = v;
        z += v;
        orc_next(wu, idu, card, val);
    }
    *partition = z;
    return 0;
}

/* code/domain.cpp:54-72 -- scope minus one variable; returns new width */
int orc_scope_minus(int w, const unsigned *ids, unsigned var, unsigned *out)
{
    int n = 0;
    for (int i = 0; i < w; ++i)
        if (ids[i] != var) out[n++] = ids[i];
    return n;
}

/* code/factor.cpp:182-212 -- returns the output width (== w when var is not in scope: deep copy) */
int orc_sum_out(int w, const unsigned *ids, const double *v, unsigned var,
                const unsigned *card, double *out, double *partition, double partition_in)
{
    uint64_t s[ORC_MAXW], so[ORC_MAXW];
    unsigned ido[ORC_MAXW], val[ORC_MAXW];
    uint64_t size = orc_strides(w, ids, card, s);
    int k = orc_find(w, ids, var);
    if (k < 0) {
        memcpy(out, v, size * sizeof(double));
        *partition = partition_in;
        return w;
    }
    int wo = orc_scope_minus(w, ids, var, ido);
    uint64_t osize = orc_strides(wo, ido, card, so);
    unsigned c = card[var];
    memset(val, 0, sizeof val);
    double z = 0;
    for (uint64_t i = 0; i < osize; ++i) {
        double acc = 0.0;
        for (unsigned x = 0; x < c; ++x) {
            /* code/domain.cpp:181-190 */
            uint64_t pos = orc_pos_consistent(w, ids, s, wo, ido, val) + s[k] * x;
            acc += v[pos];
            z += v[pos];
        }
        out[i] = acc;
        orc_next(wo, ido, card, val);
    }
    *partition = z;
    return wo;
}

/* code/domain.cpp:74-90 + code/factor.cpp:214-242 -- evidence slice.
 * ev_val[id] < 0 means "not observed". Returns the output width. */
int orc_condition(int w, const unsigned *ids, const double *v,
                  const int *ev_val, const unsigned *card,
                  unsigned *out_ids, double *out, double *partition)
{
    uint64_t s[ORC_MAXW];
    unsigned val[ORC_MAXW];
    orc_strides(w, ids, card, s);
    int wo = 0;
    uint64_t osize = 1;
    for (int i = 0; i < w; ++i) {
        if (ev_val[ids[i]] < 0) { out_ids[wo++] = ids[i]; osize *= card[ids[i]]; }
    }
    /* code/domain.cpp:138-150 */
    for (int i = 0; i < w; ++i) val[i] = ev_val[ids[i]] < 0 ? 0 : (unsigned)ev_val[ids[i]];
    double z = 0;
    for (uint64_t i = 0; i < osize; ++i) {
        uint64_t pos = 0;                       /* code/domain.cpp:152-160 */
        for (int j = w - 1; j >= 0; --j) pos += val[j] * s[j];
        out[i] = v[pos];
        z += v[pos];
        /* code/domain.cpp:125-136 -- odometer over the free digits only */
        int j;
        for (j = w - 1; j >= 0 && (ev_val[ids[j]] >= 0 || val[j] == card[ids[j]] - 1); --j) {
            if (ev_val[ids[j]] >= 0) continue;
            val[j] = 0;
        }
        if (j >= 0) val[j]++;
    }
    *partition = z;
    return wo;
}

/* code/factor.cpp:244-255 -- true division by the cached partition */
void orc_normalize(uint64_t n, const double *v, double partition, double *out)
{
    for (uint64_t i = 0; i < n; ++i) out[i] = v[i] / partition;
}

/* code/factor.cpp:97-105 (starts from 0.0) */
double orc_max(uint64_t n, const double *v)
{
    double m = 0.0;
    for (uint64_t i = 0; i < n; ++i) if (v[i] > m) m = v[i];
    return m;
}

/* code/factor.cpp:107-115 (starts from the partition) */
double orc_min(uint64_t n, const double *v, double partition)
{
    double m = partition;
    for (uint64_t i = 0; i < n; ++i) if (v[i] < m) m = v[i];
    return m;
}

double orc_sum(uint64_t n, const double *v)
{
    double z = 0;
    for (uint64_t i = 0; i < n; ++i) z += v[i];
    return z;
}

/* Fused elimination step = product of k operands followed by sum_out(var)
 * (code/model.cpp:414-418), evaluated without materialising the product and with
 * an explicit output scope (so tests can pin any layout).  Operand scopes are
 * concatenated in ids_cat with widths[]; tables in tabs[]. */
int orc_product_sum_out(int k, const int *widths, const unsigned *ids_cat, const double *const *tabs,
                        int wo, const unsigned *ido, unsigned var, int has_var,
                        const unsigned *card, double *out, double *partition)
{
    uint64_t so[ORC_MAXW];
    unsigned val[ORC_MAXW];
    uint64_t osize = orc_strides(wo, ido, card, so);
    unsigned c = has_var ? card[var] : 1;
    uint64_t (*st)[ORC_MAXW] = malloc(sizeof(uint64_t[ORC_MAXW]) * (size_t)k);
    const unsigned **sc = malloc(sizeof(unsigned *) * (size_t)k);
    int *xpos = malloc(sizeof(int) * (size_t)k);
    const unsigned *p = ids_cat;
    for (int f = 0; f < k; ++f) {
        sc[f] = p;
        orc_strides(widths[f], p, card, st[f]);
        xpos[f] = has_var ? orc_find(widths[f], p, var) : -1;
        p += widths[f];
    }
    memset(val, 0, sizeof val);
    double z = 0;
    for (uint64_t i = 0; i < osize; ++i) {
        double acc = 0.0;
        for (unsigned x = 0; x < c; ++x) {
            double prod = 1.0;
            for (int f = 0; f < k; ++f) {
                uint64_t pos = orc_pos_consistent(widths[f], sc[f], st[f], wo, ido, val);
                if (xpos[f] >= 0) pos += st[f][xpos[f]] * x;
                prod *= tabs[f][pos];
            }
            acc += prod;
        }
        out[i] = acc;
        z += acc;
        orc_next(wo, ido, card, val);
    }
    *partition = z;
    free(st); free(sc); free(xpos);
    return 0;
}

/* ------------------------------------------------------------------------
 * Factor-graph sum-product (loopy BP), code/graph.cpp:256-403.
 *
 * Flat layout: factor f has scope fscope[foff[f] .. foff[f+1]) and table
 * ftab + toff[f].  Edge e = (f, slot j) is numbered foff[f]+j; messages of edge
 * e live at moff[e] .. moff[e]+card(var).  f2v / v2f hold the two directions.
 * Within a phase every update reads only the other direction's messages
 * (code/graph.cpp:340-359, 367-388), so updating edges in any order is
 * exactly the reference's sequential sweep; only the
 * multiplication ORDER inside one message (unordered_map iteration in the
 * reference) differs, which moves results by rounding (~1e-16) only.
 * ---------------------------------------------------------------------- */
typedef struct {
    int nvars, nfac;
    const unsigned *card;
    const int *foff;       /* nfac+1 */
    const unsigned *fscope;
    const uint64_t *toff;  /* nfac */
    const double *ftab;
    const int *moff;       /* nedges+1 */
    const int *voff;       /* nvars+1: edges of variable v are vedges[voff[v] .. voff[v+1]) */
    const int *vedges;
    double *f2v, *v2f;
} orc_fg;

/* code/graph.cpp:261-274 */
void orc_bp_init(const orc_fg *g)
{
    int nedges = g->foff[g->nfac];
    for (int e = 0; e < nedges; ++e) {
        unsigned r = g->card[g->fscope[e]];
        for (unsigned i = 0; i < r; ++i) {
            g->f2v[g->moff[e] + i] = 1.0 / r;
            g->v2f[g->moff[e] + i] = 1.0 / r;
        }
    }
}

static double orc_msg_err(unsigned r, const double *oldm, const double *newm, double maxerror)
{
    /* code/graph.cpp:349-356: NaN (0/0) never raises maxerror, inf does */
    for (unsigned i = 0; i < r; ++i) {
        double err = fabs(oldm[i] - newm[i]) / oldm[i];
        if (err > maxerror) maxerror = err;
    }
    return maxerror;
}

/* one sweep of code/graph.cpp:298-332; returns maxerror */
double orc_bp_sweep(const orc_fg *g)
{
    int nedges = g->foff[g->nfac];
    double maxerror = 0.0;
    double tmp[ORC_MAXW * 4];
    /* variable -> factor, code/graph.cpp:334-362 */
    double *nv2f = malloc(sizeof(double) * (size_t)g->moff[nedges]);
    for (int e = 0; e < nedges; ++e) {
        unsigned v = g->fscope[e], r = g->card[v];
        for (unsigned i = 0; i < r; ++i) tmp[i] = 1.0;
        for (int q = g->voff[v]; q < g->voff[v + 1]; ++q) {
            int e2 = g->vedges[q];
            if (e2 == e) continue;
            for (unsigned i = 0; i < r; ++i) tmp[i] *= g->f2v[g->moff[e2] + i];
        }
        double z = 0;
        for (unsigned i = 0; i < r; ++i) z += tmp[i];
        for (unsigned i = 0; i < r; ++i) nv2f[g->moff[e] + i] = tmp[i] / z;
        maxerror = orc_msg_err(r, g->v2f + g->moff[e], nv2f + g->moff[e], maxerror);
    }
    memcpy(g->v2f, nv2f, sizeof(double) * (size_t)g->moff[nedges]);
    free(nv2f);
    /* factor -> variable, code/graph.cpp:364-391: multiply in every other
     * variable's message and sum it out, then normalise */
    for (int f = 0; f < g->nfac; ++f) {
        int w = g->foff[f + 1] - g->foff[f];
        const unsigned *sc = g->fscope + g->foff[f];
        uint64_t st[ORC_MAXW];
        uint64_t size = orc_strides(w, sc, g->card, st);
        const double *tab = g->ftab + g->toff[f];
        for (int j = 0; j < w; ++j) {
            int e = g->foff[f] + j;
            unsigned r = g->card[sc[j]];
            for (unsigned i = 0; i < r; ++i) tmp[i] = 0.0;
            for (uint64_t t = 0; t < size; ++t) {
                double p = tab[t];
                for (int u = 0; u < w; ++u) {
                    if (u == j) continue;
                    unsigned d = (unsigned)((t / st[u]) % g->card[sc[u]]);
                    p *= g->v2f[g->moff[g->foff[f] + u] + d];
                }
                tmp[(t / st[j]) % r] += p;
            }
            double z = 0;
            for (unsigned i = 0; i < r; ++i) z += tmp[i];
            double newm[ORC_MAXW * 4];
            for (unsigned i = 0; i < r; ++i) newm[i] = tmp[i] / z;
            maxerror = orc_msg_err(r, g->f2v + g->moff[e], newm, maxerror);
            for (unsigned i = 0; i < r; ++i) g->f2v[g->moff[e] + i] = newm[i];
        }
    }
    return maxerror;
}

/* code/graph.cpp:298-332 -- returns the 0-based index of the converging sweep, or maxit */
unsigned orc_bp_update(const orc_fg *g, unsigned maxit, double eps)
{
    unsigned it;
    for (it = 0; it < maxit; ++it) {
        double maxerror = orc_bp_sweep(g);
        if (maxerror < eps) break;
    }
    return it;
}

/* code/graph.cpp:393-403 */
void orc_bp_marginal(const orc_fg *g, unsigned var, double *out)
{
    unsigned r = g->card[var];
    for (unsigned i = 0; i < r; ++i) out[i] = 1.0;
    for (int q = g->voff[var]; q < g->voff[var + 1]; ++q) {
        int e = g->vedges[q];
        for (unsigned i = 0; i < r; ++i) out[i] *= g->f2v[g->moff[e] + i];
    }
    double z = 0;
    for (unsigned i = 0; i < r; ++i) z += out[i];
    for (unsigned i = 0; i < r; ++i) out[i] = out[i] / z;
}
