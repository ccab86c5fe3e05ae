"""TEST INFRASTRUCTURE ONLY -- Python face of the CPU oracle.

Two checkers live here; neither is ever imported by the shipped package
(`bnpp_b200/`), only by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs:

* `OFactor` + the free functions below wrap oracle/factor_oracle.c (plain-C
  restatement of code/factor.cpp:97-255 and code/domain.cpp:15-190) and restate
  the reference's drivers on top of it: bucket elimination
  (code/model.cpp:348-446), BN::partition / BN::marginals (code/model.cpp:250-346),
  Model::joint_distribution (code/model.cpp:41-49), loopy sum-product
  (code/graph.cpp:256-403) and the UAI reader (code/io.cpp:43-180).
* `RefHarness` drives oracle/_ref/ref_harness, the UNMODIFIED reference compiled
  from its own sources (see oracle/Makefile); it is the 1e-9 source of truth and
  what the restatement itself is pinned against (tests/test_oracle.py).

Elimination ORDERS are an input here (taken from the golden fixtures or from the
product's host-side Graph): the reference's tie-breaks follow libstdc++ hash-table
iteration order, which only the real reference pins.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_double_p = ctypes.POINTER(ctypes.c_double)
c_uint_p = ctypes.POINTER(ctypes.c_uint)
c_int_p = ctypes.POINTER(ctypes.c_int)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle_c.so")
        if not os.path.exists(path):
            subprocess.check_call(["make", "-C", _HERE, "liboracle_c.so"])
        L = ctypes.CDLL(path)
        L.orc_domain_size.restype = ctypes.c_uint64
        L.orc_max.restype = ctypes.c_double
        L.orc_min.restype = ctypes.c_double
        L.orc_sum.restype = ctypes.c_double
        L.orc_bp_sweep.restype = ctypes.c_double
        L.orc_bp_update.restype = ctypes.c_uint
        _LIB = L
    return _LIB


def _u(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def _up(a):
    return a.ctypes.data_as(c_uint_p)


def _dp(a):
    return a.ctypes.data_as(c_double_p)


class OFactor:
    """Dense fp64 potential table: scope ids (last fastest) + values + cached partition."""

    def __init__(self, scope, values, partition=None):
        self.scope = [int(s) for s in scope]
        self.values = np.ascontiguousarray(values, dtype=np.float64).reshape(-1)
        # code/io.cpp:90-97: sequential sum in file order
        self.partition = float(lib().orc_sum(ctypes.c_uint64(self.values.size), _dp(self.values))) \
            if partition is None else float(partition)

    @property
    def width(self):
        return len(self.scope)

    @property
    def size(self):
        return int(self.values.size)

    def by_valuation(self, card):
        """dict valuation-tuple(sorted by var id) -> value; layout-independent comparison."""
        shape = [int(card[v]) for v in self.scope]
        arr = self.values.reshape(shape) if shape else self.values.reshape(())
        order = np.argsort(self.scope)
        return sorted(self.scope), np.transpose(arr, order) if shape else arr


def scalar(value=1.0):
    """code/factor.cpp:25-30"""
    return OFactor([], [value], value)


def domain_size(scope, card):
    s = _u(scope)
    return int(lib().orc_domain_size(len(scope), _up(s), _up(_u(card))))


def _binary(a, b, card, divide):
    L = lib()
    card = _u(card)
    ida, idb = _u(a.scope), _u(b.scope)
    idu = np.zeros(len(a.scope) + len(b.scope) + 1, dtype=np.uint32)
    wu = L.orc_union_scope(a.width, _up(ida), b.width, _up(idb), _up(idu))
    scope = [int(x) for x in idu[:wu]]
    out = np.empty(domain_size(scope, card), dtype=np.float64)
    z = ctypes.c_double()
    rc = L.orc_product(a.width, _up(ida), _dp(a.values), b.width, _up(idb), _dp(b.values),
                       _up(card), int(divide), _dp(out), ctypes.byref(z))
    if rc != 0:
        raise ZeroDivisionError("Factor::divide: zero divisor (code/factor.cpp:169 asserts)")
    return OFactor(scope, out, z.value)


def product(a, b, card):
    """code/factor.cpp:117-147"""
    return _binary(a, b, card, False)


def divide(a, b, card):
    """code/factor.cpp:149-180"""
    return _binary(a, b, card, True)


def sum_out(a, var, card):
    """code/factor.cpp:182-212"""
    L = lib()
    card = _u(card)
    ids = _u(a.scope)
    if var in a.scope:
        scope = [s for s in a.scope if s != var]
    else:
        scope = list(a.scope)
    out = np.empty(domain_size(scope, card), dtype=np.float64)
    z = ctypes.c_double()
    L.orc_sum_out(a.width, _up(ids), _dp(a.values), ctypes.c_uint(var), _up(card), _dp(out),
                  ctypes.byref(z), ctypes.c_double(a.partition))
    return OFactor(scope, out, z.value)


def condition(a, evidence, card):
    """code/factor.cpp:214-242; evidence = {var id: value}"""
    L = lib()
    card = _u(card)
    ids = _u(a.scope)
    ev = np.full(len(card), -1, dtype=np.int32)
    for k, v in evidence.items():
        if k < len(card):
            ev[k] = v
    scope = [s for s in a.scope if s not in evidence]
    out_ids = np.zeros(a.width + 1, dtype=np.uint32)
    out = np.empty(domain_size(scope, card), dtype=np.float64)
    z = ctypes.c_double()
    wo = L.orc_condition(a.width, _up(ids), _dp(a.values), ev.ctypes.data_as(c_int_p), _up(card),
                         _up(out_ids), _dp(out), ctypes.byref(z))
    assert [int(x) for x in out_ids[:wo]] == scope
    return OFactor(scope, out, z.value)


def normalize(a):
    """code/factor.cpp:244-255"""
    out = np.empty_like(a.values)
    lib().orc_normalize(ctypes.c_uint64(a.size), _dp(a.values), ctypes.c_double(a.partition), _dp(out))
    return OFactor(a.scope, out, 1.0)


def fmax(a):
    return float(lib().orc_max(ctypes.c_uint64(a.size), _dp(a.values)))


def fmin(a):
    return float(lib().orc_min(ctypes.c_uint64(a.size), _dp(a.values), ctypes.c_double(a.partition)))


def product_sum_out(operands, out_scope, var, card):
    """Fused elimination step (code/model.cpp:414-418) with an explicit output scope.

    var=None means a pure k-ary product laid out in `out_scope`."""
    L = lib()
    card = _u(card)
    k = len(operands)
    widths = np.array([f.width for f in operands], dtype=np.int32)
    ids_cat = _u([s for f in operands for s in f.scope] or [0])
    tabs = (c_double_p * k)(*[_dp(f.values) for f in operands])
    ido = _u(out_scope if len(out_scope) else [0])
    out = np.empty(domain_size(out_scope, card), dtype=np.float64)
    z = ctypes.c_double()
    L.orc_product_sum_out(k, widths.ctypes.data_as(c_int_p), _up(ids_cat), tabs,
                          len(out_scope), _up(ido), ctypes.c_uint(0 if var is None else var),
                          int(var is not None), _up(card), _dp(out), ctypes.byref(z))
    return OFactor(out_scope, out, z.value)


# --------------------------------------------------------------------------
# UAI reader, code/io.cpp:14-180
# --------------------------------------------------------------------------
def _tokens(text):
    for line in text.splitlines():
        for tok in line.split():
            if tok.startswith("#"):
                break  # code/io.cpp:19-20: rest of the line is a comment
            yield tok


class OModel:
    def __init__(self, kind, card, factors):
        self.kind = kind
        self.card = _u(card)
        self.factors = factors

    @property
    def nvars(self):
        return len(self.card)


def parse_uai(text):
    """code/io.cpp:43-100"""
    it = _tokens(text)
    kind = next(it)
    n = int(next(it))
    card = [int(next(it)) for _ in range(n)]
    m = int(next(it))
    scopes = []
    for _ in range(m):
        w = int(next(it))
        scopes.append([int(next(it)) for _ in range(w)])
    factors = []
    for sc in scopes:
        sz = int(next(it))
        vals = [float(next(it)) for _ in range(sz)]
        factors.append(OFactor(sc, vals))
    return OModel(kind, card, factors)


def read_uai(path):
    with open(path) as f:
        return parse_uai(f.read())


def parse_evidence(text):
    """code/io.cpp:157-180 -- honoured only when the leading sample count is exactly 1"""
    it = _tokens(text)
    ev = {}
    n = int(next(it))
    if n == 1:
        k = int(next(it))
        for _ in range(k):
            i = int(next(it))
            ev[i] = int(next(it))
    return ev


# --------------------------------------------------------------------------
# Drivers, code/model.cpp
# --------------------------------------------------------------------------
def variable_elimination(order, factors, card):
    """code/model.cpp:382-445 with buckets in insertion order."""
    result = scalar(1.0)
    buckets = {v: [] for v in order}
    remaining = list(order)
    for f in factors:
        for v in remaining:
            if v in f.scope:
                buckets[v].append(f)
                break
        else:
            result = product(result, f, card)
    while remaining:
        var = remaining.pop(0)
        prod = scalar(1.0)
        for f in buckets[var]:
            prod = product(prod, f, card)
        newf = sum_out(prod, var, card)
        for v in remaining:
            if v in newf.scope:
                buckets[v].append(newf)
                break
        else:
            result = product(result, newf, card)
    return result


def partition(model, evidence, order=None):
    """code/model.cpp:275-294; `order` = elimination order over the unobserved variables."""
    card = model.card
    if order is None:
        order = [v for v in range(model.nvars) if v not in evidence]
    factors = [condition(f, evidence, card) for f in model.factors]
    part = variable_elimination(order, factors, card)
    assert part.size == 1
    return part.partition


def marginals(model, evidence, order_for=None):
    """code/model.cpp:320-339; order_for(v) -> elimination order for the VE pass of variable v."""
    card = model.card
    factors = [condition(f, evidence, card) for f in model.factors]
    out = []
    for v in range(model.nvars):
        order = [u for u in range(model.nvars) if u != v] if order_for is None else order_for(v)
        out.append(normalize(variable_elimination(order, factors, card)))
    return out


def joint(model, evidence):
    """code/model.cpp:41-49 (what `mn` runs)"""
    f = scalar(1.0)
    for pf in model.factors:
        f = product(f, condition(pf, evidence, model.card), model.card)
    return f


def joint_marginals(model, evidence):
    """code/model.cpp:69-101"""
    j = normalize(joint(model, evidence))
    out = []
    for v in range(model.nvars):
        f = j
        for u in range(model.nvars):
            if u != v:
                f = sum_out(f, u, model.card)
        out.append(f)
    return out


# --------------------------------------------------------------------------
# Loopy sum-product, code/graph.cpp:256-403
# --------------------------------------------------------------------------
class _FG(ctypes.Structure):
    _fields_ = [("nvars", ctypes.c_int), ("nfac", ctypes.c_int), ("card", c_uint_p),
                ("foff", c_int_p), ("fscope", c_uint_p),
                ("toff", ctypes.POINTER(ctypes.c_uint64)), ("ftab", c_double_p),
                ("moff", c_int_p), ("voff", c_int_p), ("vedges", c_int_p),
                ("f2v", c_double_p), ("v2f", c_double_p)]


class OFactorGraph:
    def __init__(self, card, factors):
        self.card = _u(card)
        self.factors = factors
        foff = [0]
        fscope = []
        toff = []
        tabs = []
        t = 0
        for f in factors:
            fscope += f.scope
            foff.append(len(fscope))
            toff.append(t)
            tabs.append(f.values)
            t += f.size
        self.foff = np.array(foff, dtype=np.int32)
        self.fscope = _u(fscope if fscope else [0])
        self.toff = np.array(toff if toff else [0], dtype=np.uint64)
        self.ftab = np.concatenate(tabs) if tabs else np.zeros(1)
        moff = [0]
        for v in fscope:
            moff.append(moff[-1] + int(self.card[v]))
        self.moff = np.array(moff, dtype=np.int32)
        by_var = [[] for _ in range(len(card))]
        for e, v in enumerate(fscope):
            by_var[v].append(e)
        voff = [0]
        vedges = []
        for v in range(len(card)):
            vedges += by_var[v]
            voff.append(len(vedges))
        self.voff = np.array(voff, dtype=np.int32)
        self.vedges = np.array(vedges if vedges else [0], dtype=np.int32)
        self.f2v = np.zeros(max(1, moff[-1]))
        self.v2f = np.zeros(max(1, moff[-1]))
        self.g = _FG(len(card), len(factors), _up(self.card),
                     self.foff.ctypes.data_as(c_int_p), _up(self.fscope),
                     self.toff.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), _dp(self.ftab),
                     self.moff.ctypes.data_as(c_int_p), self.voff.ctypes.data_as(c_int_p),
                     self.vedges.ctypes.data_as(c_int_p), _dp(self.f2v), _dp(self.v2f))
        lib().orc_bp_init(ctypes.byref(self.g))

    def update(self, maxit=10000, eps=0.001):
        """code/graph.cpp:298-332 -> index of the converging sweep"""
        return int(lib().orc_bp_update(ctypes.byref(self.g), ctypes.c_uint(maxit), ctypes.c_double(eps)))

    def sweep(self):
        return float(lib().orc_bp_sweep(ctypes.byref(self.g)))

    def marginal(self, var):
        """code/graph.cpp:393-403"""
        out = np.zeros(int(self.card[var]))
        lib().orc_bp_marginal(ctypes.byref(self.g), ctypes.c_uint(var), _dp(out))
        return out


# --------------------------------------------------------------------------
# The real reference, compiled in place (oracle/Makefile -> oracle/_ref/)
# --------------------------------------------------------------------------
REF_HARNESS = os.path.join(_HERE, "_ref", "ref_harness")


def have_ref():
    return os.path.exists(REF_HARNESS)


class RefHarness:
    """Runs a command script through oracle/_ref/ref_harness and parses the replies."""

    def __init__(self):
        if not have_ref():
            raise RuntimeError("oracle/_ref/ref_harness not built (needs /root/reference; run make -C oracle)")

    def run(self, script, timeout=600):
        if isinstance(script, (list, tuple)):
            script = "\n".join(script)
        p = subprocess.run([REF_HARNESS], input=script + "\n", capture_output=True, text=True, timeout=timeout)
        if p.returncode != 0:
            raise RuntimeError("ref_harness failed rc=%d: %s" % (p.returncode, p.stderr[-400:]))
        return [ln.split() for ln in p.stdout.splitlines() if ln.strip()]

    @staticmethod
    def factors(rows):
        """Collect FACTOR/V reply pairs -> list of (scope, size, partition, values)."""
        out = []
        for i, r in enumerate(rows):
            if r[0] == "FACTOR":
                w = int(r[1])
                scope = [int(x) for x in r[2:2 + w]]
                size = int(r[2 + w])
                z = float(r[3 + w])
                vals = np.array([float(x) for x in rows[i + 1][1:]])
                out.append((scope, size, z, vals))
        return out

    @staticmethod
    def marginals(rows):
        return [np.array([float(x) for x in r[3:]]) for r in rows if r[0] == "M"]
