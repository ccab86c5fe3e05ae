"""Synthetic UAI workloads named by BASELINE.json `configs`.

Pure-Python generators (stdlib `random.Random(seed)` so the files are
reproducible byte for byte); values are written with `%.17g` so the reference's
`stod` reader (code/io.cpp:35-41) recovers them exactly.

* `ising_uai`      -- config 3: n x n binary Ising grid, unary factors first.
* `random_bn_uai`  -- configs 4/5: "window" Bayesian network in the reference's
                      CHILD-FIRST scope convention (code/model.cpp:111-119).
* `evidence_batch` -- config 5: evidence sets over one fixed set of observed ids.
"""
import math
import random


def ising_uai(n, h=0.5, J=0.5, seed=7, header="MARKOV"):
    """n*n binary variables, row-major id = r*n+c.  Factors: n*n unary
    [e^t, e^-t], t~U(-h,h); then for every cell its right and its down
    neighbour (if any) a pairwise [e^w, e^-w, e^-w, e^w], w~U(-J,J)."""
    rng = random.Random(seed)
    N = n * n
    scopes = [[i] for i in range(N)]
    tables = []
    for _ in range(N):
        t = rng.uniform(-h, h)
        tables.append([math.exp(t), math.exp(-t)])
    for r in range(n):
        for c in range(n):
            i = r * n + c
            for j in ([i + 1] if c + 1 < n else []) + ([i + n] if r + 1 < n else []):
                w = rng.uniform(-J, J)
                scopes.append([i, j])
                tables.append([math.exp(w), math.exp(-w), math.exp(-w), math.exp(w)])
    return _emit(header, [2] * N, scopes, tables)


def random_bn_scopes(N, W, K, seed):
    rng = random.Random(seed)
    scopes = []
    for i in range(N):
        lo = max(0, i - W)
        k = min(K, i - lo)
        parents = sorted(rng.sample(range(lo, i), k)) if k > 0 else []
        scopes.append([i] + parents)
    return scopes, rng


def random_bn_uai(N, W, K, seed, header="BAYES"):
    """Variable i has min(K, i) parents drawn uniformly from the previous W
    variables; CPT P(x_i = 0 | pa) ~ U(0.05, 0.95).  Scope = [child, parents...]
    so the child digit is the MOST significant one of its table."""
    scopes, rng = random_bn_scopes(N, W, K, seed)
    tables = []
    for sc in scopes:
        npa = 1 << (len(sc) - 1)
        p0 = [rng.uniform(0.05, 0.95) for _ in range(npa)]
        tables.append(p0 + [1.0 - p for p in p0])
    return _emit(header, [2] * N, scopes, tables)


def evidence_batch(nvars, nobs, nsets, seed=5, fixed_ids=True, card=2):
    """-> list of {id: value} dicts.  fixed_ids: one observed-id set shared by all
    evidence sets (one elimination order, one plan), values random per set."""
    rng = random.Random(seed)
    ids = sorted(rng.sample(range(nvars), nobs))
    out = []
    for _ in range(nsets):
        if not fixed_ids:
            ids = sorted(rng.sample(range(nvars), nobs))
        out.append({i: rng.randrange(card) for i in ids})
    return out


def evidence_text(ev):
    """UAI evidence file body understood by code/io.cpp:157-180."""
    return "1\n%d %s\n" % (len(ev), " ".join("%d %d" % kv for kv in sorted(ev.items())))


def _emit(header, card, scopes, tables):
    out = [header, str(len(card)), " ".join(str(c) for c in card), str(len(scopes))]
    for sc in scopes:
        out.append("%d %s" % (len(sc), " ".join(str(v) for v in sc)))
    for t in tables:
        out.append("%d %s" % (len(t), " ".join("%.17g" % v for v in t)))
    return "\n".join(out) + "\n"


# config 4 per GPU count: generator parameters (N, W, K, seed) and how many leaves are observed.  Chosen with the host
# orderer so that the min-fill width is 27 + log2(G) and fixing the log2(G) shard variables leaves every rank a
# width-27 problem of ~2.0-2.2e9 union entries.  The 1-GPU network has six leaves only and observing more than four
# of them tips a min-fill tie into a width-29 order, hence 4 there.
WIDE = {1: (64, 40, 4, 5, 4), 2: (68, 40, 4, 27, 8), 4: (72, 40, 4, 23, 8), 8: (76, 44, 4, 3, 8)}
# the strong-scaling network: ONE fixed network (the 8-GPU one, min-fill width 30) at every GPU count
STRONG = WIDE[8]


def wide_bn(key):
    """-> (N, W, K, seed, evidence) of config 4: `key` is a GPU count of WIDE or the string "strong".
    evidence = observed leaves (variables that are nobody's parent) with values from Random(seed + 100)."""
    N, W, K, seed, nobs = STRONG if key == "strong" else WIDE[key]
    scopes, _ = random_bn_scopes(N, W, K, seed)
    parents = set(v for sc in scopes for v in sc[1:])
    leaves = [v for v in range(N) if v not in parents]
    rng = random.Random(seed + 100)
    obs = sorted(rng.sample(leaves, min(nobs, len(leaves))))
    return N, W, K, seed, {v: rng.randrange(2) for v in obs}
