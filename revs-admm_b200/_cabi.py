"""ctypes binding of include/revs_admm.h.  No torch types cross this boundary.

The library is the product: if it is missing or no B200 is visible every call raises --
there is no CPU path in this package.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("REVS_LIB") or os.path.join(HERE, "librevs_admm.so")     # REVS_LIB: a build variant (profiles/build_variants.sh)

REVS_REL_VOLTAGE, REVS_REL_FLOW, REVS_REL_DROP = 0, 1, 2


class RevsError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"revs_admm error {code}: {msg}")
        self.code = code


class Stats(C.Structure):
    _fields_ = [("kernel_launches", C.c_int64), ("gemm_launches", C.c_int64), ("gemm_full_launches", C.c_int64),
                ("qp_outer_iterations", C.c_int64), ("qp_newton_iterations", C.c_int64),
                ("admm_iterations", C.c_int32), ("max_working_set", C.c_int32),
                ("primal_residual", C.c_double), ("dual_residual", C.c_double), ("qp_flops", C.c_double),
                ("gemm_ms", C.c_float), ("gemm_full_ms", C.c_float), ("home_ms", C.c_float), ("dual_ms", C.c_float),
                ("qp_ms", C.c_float), ("qp_big_ms", C.c_float), ("total_ms", C.c_float),
                ("qp_warp_ms", C.c_float), ("qp_init_ms", C.c_float), ("qp_columns", C.c_int64),
                ("qp_warp_rounds", C.c_int64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_P = C.c_void_p
_D = C.POINTER(C.c_double)
# name -> argtypes   (every symbol include/revs_admm.h declares)
SIGNATURES = {
    "revs_last_error": ([], C.c_char_p),
    "revs_version": ([], C.c_int),
    "revs_device_count": ([C.POINTER(C.c_int)], C.c_int),
    "revs_create": ([C.POINTER(_P), C.c_int, C.c_int, C.POINTER(C.c_int64), C.c_int], C.c_int),
    "revs_destroy": ([_P], C.c_int),
    "revs_set_sensitivity": ([_P, C.c_int, _D], C.c_int),
    "revs_set_feeder_tree": ([_P, C.c_int, C.c_int, C.POINTER(C.c_int32), _D, C.POINTER(C.c_int32)], C.c_int),
    "revs_set_feeder_trees": ([_P, C.POINTER(C.c_int64), C.POINTER(C.c_int32), _D, C.POINTER(C.c_int32)], C.c_int),
    "revs_set_homes": ([_P, _D, C.POINTER(C.c_uint8), _D, _D, _D, C.POINTER(C.c_int32),
                        C.POINTER(C.c_int32)], C.c_int),
    "revs_set_tariff": ([_P, _D], C.c_int),
    "revs_solve_admm": ([_P, C.c_double, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double,
                         C.POINTER(C.c_int)], C.c_int),
    "revs_admm_begin": ([_P, C.c_double, C.c_int, C.c_double, C.c_double, C.c_double], C.c_int),
    "revs_admm_step": ([_P, _D], C.c_int),
    "revs_home_step": ([_P, C.c_double, _D, _D, _D, _D, _D], C.c_int),
    "revs_utility_step": ([_P, C.c_double, C.c_double, C.c_double, C.c_double, _D, _D, _D, _D, _D, _D], C.c_int),
    "revs_get_results": ([_P, _D, _D, _D, _D, C.c_int], C.c_int),
    "revs_get_schedule": ([_P, _D, C.POINTER(C.c_uint64), C.c_int, _D, C.c_int], C.c_int),
    "revs_get_schedule_ld": ([_P, _D, C.POINTER(C.c_uint64), C.c_int, _D, C.c_int, C.c_int64], C.c_int),
    "revs_get_estimate": ([_P, _D, _D], C.c_int),
    "revs_solve_individual": ([_P, _D, _D, _D], C.c_int),
    "revs_reliability": ([_P, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int32), _D, C.c_double, _D, _D], C.c_int),
    "revs_contract": ([C.c_int, C.c_int, C.c_int, C.c_int, _D, _D, _D], C.c_int),
    "revs_set_option": ([_P, C.c_char_p, C.c_double], C.c_int),
    "revs_screen_contract": ([C.c_int, C.c_int, C.c_int, C.c_int, _D, _D, _D, C.c_int], C.c_int),
    "revs_get_stats": ([_P, C.POINTER(Stats)], C.c_int),
    "revs_zone_arrays": ([C.c_int, C.POINTER(C.c_int32), _D, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32), _D, _D, _D,
                          C.POINTER(C.c_int32), _D, C.POINTER(C.c_int32), _D, C.POINTER(C.c_int32)], C.c_int),
    "revs_comm_export": ([_P, C.c_void_p], C.c_int),
    "revs_comm_attach": ([_P, C.c_int, C.c_int, C.c_void_p], C.c_int),
    "revs_comm_detach": ([_P], C.c_int),
    "revs_gather_export": ([_P, C.c_int64, C.c_void_p], C.c_int),
    "revs_gather_attach": ([_P, C.c_int, C.c_int, C.c_void_p], C.c_int),
    "revs_reliability_sharded": ([_P, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int32), _D, C.c_double, _D, _D], C.c_int),
}

_lib = None


def load():
    """dlopen the in-tree library and type every entry point."""
    global _lib
    if _lib is None:
        path = os.environ.get("REVS_LIB") or LIB_PATH        # REVS_LIB: another build of this library (A/B measurements)
        if not os.path.exists(path):
            raise RevsError(-1, f"{path} is missing -- run `python revs-admm_b200/_build.py` "
                                "(there is no CPU fallback)")
        lib = C.CDLL(path)
        for name, (args, res) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.argtypes, fn.restype = args, res
        _lib = lib
    return _lib


def _dp(a):
    return None if a is None else a.ctypes.data_as(_D)


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None and a.shape != tuple(shape):
        raise ValueError(f"expected shape {tuple(shape)}, got {a.shape}")
    return a


def _check(rc):
    if rc != 0:
        raise RevsError(rc, load().revs_last_error().decode())


def device_count():
    n = C.c_int(0)
    _check(load().revs_device_count(C.byref(n)))
    return n.value


def contract(A, B, device=0):
    """C = A @ B through the FP64 tensor-core contraction kernel (host in/out)."""
    A, B = _f64(A), _f64(B)
    M, K = A.shape
    K2, T = B.shape
    assert K == K2
    out = np.empty((M, T))
    _check(load().revs_contract(device, M, K, T, _dp(A), _dp(B), _dp(out)))
    return out


def screen_contract(A, B, impl=0, device=0):
    """C ~ A @ B through the BF16 screening kernel (impl 0: mma.sync, 1: tcgen05/TMA)."""
    A, B = _f64(A), _f64(B)
    M, K = A.shape
    K2, T = B.shape
    assert K == K2
    out = np.empty((M, T))
    _check(load().revs_screen_contract(device, M, K, T, _dp(A), _dp(B), _dp(out), impl))
    return out


def expand_schedule(mask, T, has_ev, rating, capacity, initial):
    """P_ev [H,T] and SOC [H,T+1] (the S and C of lpsolver.py:289-290) from the charging bit masks of
    Solver.schedule_compact and the per-home inputs: P_ev = rating * bit, SOC[t+1] = SOC[t] + P_ev[t] / capacity
    (same operation order as the device kernel, so bit-identical with Solver.results)."""
    mask = np.asarray(mask, dtype=np.uint64)
    H = mask.shape[0]
    t = np.arange(T)
    bits = ((mask[:, t // 64] >> (t % 64).astype(np.uint64)) & np.uint64(1)).astype(bool)
    ev = np.asarray(has_ev).astype(bool)
    p_ev = np.where(bits & ev[:, None], np.asarray(rating, dtype=np.float64)[:, None], 0.0)
    soc = np.zeros((H, T + 1))
    soc[:, 0] = np.where(ev, initial, 0.0)
    inc = p_ev / np.where(ev, capacity, 1.0)[:, None]
    for k in range(T):
        soc[:, k + 1] = soc[:, k] + inc[:, k]
    return p_ev, soc


def zone_arrays(parent, r, res_node):
    """Static arrays of the tree-structured operator kernel for one radial zone (host only, no GPU):
    dict(perm, c, d, e, lo, hi, w) with the Cartesian-tree nodes in lo-order, plus the hi-order copy and cnt."""
    parent = np.ascontiguousarray(parent, dtype=np.int32)
    r = _f64(r, (len(parent),))
    res_node = np.ascontiguousarray(res_node, dtype=np.int32)
    n = len(res_node)
    i32 = C.POINTER(C.c_int32)
    I = lambda: np.zeros(n, dtype=np.int32)
    perm, nlo, nhi, cnt = I(), I(), I(), I()
    c, d, e, wlo, whi = (np.zeros(n) for _ in range(5))
    _check(load().revs_zone_arrays(len(parent), parent.ctypes.data_as(i32), _dp(r), n, res_node.ctypes.data_as(i32),
                                   perm.ctypes.data_as(i32), _dp(c), _dp(d), _dp(e), nlo.ctypes.data_as(i32), _dp(wlo),
                                   nhi.ctypes.data_as(i32), _dp(whi), cnt.ctypes.data_as(i32)))
    m = max(n - 1, 0)
    return dict(perm=perm, c=c[:m], d=d, e=e, lo=nlo[:m] & 0xffff, hi=nlo[:m] >> 16, w=wlo[:m],
                lo_b=nhi[:m] & 0xffff, hi_b=nhi[:m] >> 16, w_b=whi[:m], cnt_lo=cnt & 0xffff, cnt_hi=cnt >> 16)


class Solver:
    """A batch of feeders resident on one GPU (wraps revs_solver*)."""

    def __init__(self, feeder_sizes, T, device=0):
        self.lib = load()
        self.sizes = [int(n) for n in feeder_sizes]
        self.off = np.concatenate([[0], np.cumsum(self.sizes)]).astype(np.int64)
        self.H, self.T, self.nf = int(self.off[-1]), int(T), len(self.sizes)
        self._h = _P()
        _check(self.lib.revs_create(C.byref(self._h), device, self.nf,
                                    self.off.ctypes.data_as(C.POINTER(C.c_int64)), self.T))

    def close(self):
        if getattr(self, "_h", None):
            self.lib.revs_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- inputs
    def set_sensitivity(self, feeder, R):
        n = self.sizes[feeder]
        R = _f64(R, (n, n))
        _check(self.lib.revs_set_sensitivity(self._h, feeder, _dp(R)))

    def set_feeder_tree(self, feeder, parent, r, res_node):
        parent = np.ascontiguousarray(parent, dtype=np.int32)
        r = _f64(r, (len(parent),))
        res_node = np.ascontiguousarray(res_node, dtype=np.int32)
        assert len(res_node) == self.sizes[feeder]
        i32 = C.POINTER(C.c_int32)
        _check(self.lib.revs_set_feeder_tree(self._h, feeder, len(parent), parent.ctypes.data_as(i32),
                                             _dp(r), res_node.ctypes.data_as(i32)))

    def pack_trees(self, trees):
        """The concatenated arrays revs_set_feeder_trees takes, from feeder.FeederTree-like objects (parent, r, res_node)."""
        assert len(trees) == self.nf
        node_off = np.concatenate([[0], np.cumsum([len(t.parent) for t in trees])]).astype(np.int64)
        parent = np.ascontiguousarray(np.concatenate([t.parent for t in trees]), dtype=np.int32)
        r = _f64(np.concatenate([t.r for t in trees]))
        res = np.ascontiguousarray(np.concatenate([t.res_node for t in trees]), dtype=np.int32)
        assert len(res) == self.H
        return node_off, parent, r, res

    def set_feeder_trees(self, trees, packed=None):
        """All feeders in one call; `trees` are feeder.FeederTree-like (parent, r, res_node), or `packed` = pack_trees(trees)."""
        node_off, parent, r, res = packed if packed is not None else self.pack_trees(trees)
        i32 = C.POINTER(C.c_int32)
        _check(self.lib.revs_set_feeder_trees(self._h, node_off.ctypes.data_as(C.POINTER(C.c_int64)),
                                              parent.ctypes.data_as(i32), _dp(r), res.ctypes.data_as(i32)))

    def set_homes(self, load, has_ev, rating, capacity, initial, start, end):
        H, T = self.H, self.T
        load = _f64(load, (H, T))
        has_ev = np.ascontiguousarray(has_ev, dtype=np.uint8)
        rating, capacity, initial = _f64(rating, (H,)), _f64(capacity, (H,)), _f64(initial, (H,))
        start = np.ascontiguousarray(start, dtype=np.int32)
        end = np.ascontiguousarray(end, dtype=np.int32)
        i32 = C.POINTER(C.c_int32)
        _check(self.lib.revs_set_homes(self._h, _dp(load), has_ev.ctypes.data_as(C.POINTER(C.c_uint8)),
                                       _dp(rating), _dp(capacity), _dp(initial),
                                       start.ctypes.data_as(i32), end.ctypes.data_as(i32)))

    def set_tariff(self, cost):
        cost = _f64(cost, (self.T,))
        _check(self.lib.revs_set_tariff(self._h, _dp(cost)))

    # ---- solves
    def solve_admm(self, kappa=5.0, iter_max=15, vset=1.0, vlow=0.95, vhigh=1.05, tol=0.0):
        done = C.c_int(0)
        _check(self.lib.revs_solve_admm(self._h, kappa, iter_max, vset, vlow, vhigh, tol, C.byref(done)))
        return done.value

    def admm_begin(self, kappa=5.0, iter_max=15, vset=1.0, vlow=0.95, vhigh=1.05):
        _check(self.lib.revs_admm_begin(self._h, kappa, iter_max, vset, vlow, vhigh))

    def admm_step(self):
        sums = np.zeros(3)
        _check(self.lib.revs_admm_step(self._h, _dp(sums)))
        return sums

    def home_step(self, p_est, p_sch, gamma, kappa=5.0):
        H, T = self.H, self.T
        g, p = np.empty((H, T)), np.empty((H, T))
        _check(self.lib.revs_home_step(self._h, kappa, _dp(_f64(p_est, (H, T))), _dp(_f64(p_sch, (H, T))),
                                       _dp(_f64(gamma, (H, T))), _dp(g), _dp(p)))
        return g, p

    def utility_step(self, p_est, p_sch, gamma, kappa=5.0, vset=1.0, vlow=0.95, vhigh=1.05, lam0=None):
        H, T = self.H, self.T
        g, lam = np.empty((H, T)), np.empty((H, T))
        l0 = None if lam0 is None else _f64(lam0, (H, T))
        _check(self.lib.revs_utility_step(self._h, kappa, vset, vlow, vhigh, _dp(_f64(p_est, (H, T))),
                                          _dp(_f64(p_sch, (H, T))), _dp(_f64(gamma, (H, T))), _dp(l0),
                                          _dp(g), _dp(lam)))
        return g, lam

    def results(self, iters=None, want_diff=True, out=None):
        """`out`: optional dict of preallocated (e.g. page-locked) arrays P_sch, P_ev, SOC, diff."""
        H, T = self.H, self.T
        done = self.stats()["admm_iterations"]
        iters = done if iters is None else iters
        if out is not None:
            P, E, S, D = out["P_sch"], out["P_ev"], out["SOC"], out.get("diff")
            for a, shape in ((P, (H, T)), (E, (H, T)), (S, (H, T + 1))):
                if a.shape != shape or a.dtype != np.float64 or not a.flags.c_contiguous:
                    raise ValueError(f"out arrays must be C-contiguous float64 of shape {shape}")
            if D is not None and (D.ndim != 2 or D.shape[1] != H or D.dtype != np.float64 or not D.flags.c_contiguous):
                raise ValueError("out['diff'] must be a C-contiguous float64 [rows, H] array")
        else:
            P, E, S = np.empty((H, T)), np.empty((H, T)), np.empty((H, T + 1))
            D = np.empty((max(iters, done), H)) if want_diff else None
        # the library refuses a diff buffer with fewer rows than iterations that ran
        _check(self.lib.revs_get_results(self._h, _dp(P), _dp(E), _dp(S), _dp(D), 0 if D is None else D.shape[0]))
        return dict(P_sch=P, P_ev=E, SOC=S, diff=None if D is None else D[:done] if out is None else D)

    def schedule_compact(self, want_diff=True, out=None):
        """P_sch, the charging decisions as bit masks [H, ceil(T/64)] (uint64) and diff -- a third of the
        bytes of results(); expand_schedule() rebuilds P_ev and SOC from the mask on the host."""
        H, T = self.H, self.T
        W = (T + 63) // 64
        done = self.stats()["admm_iterations"]
        if out is None:
            out = dict(P_sch=np.empty((H, T)), mask=np.empty((H, W), dtype=np.uint64),
                       diff=np.empty((done, H)) if want_diff else None)
        P, M, D = out["P_sch"], out["mask"], out.get("diff")
        if P.shape != (H, T) or P.dtype != np.float64 or not P.flags.c_contiguous:
            raise ValueError("out['P_sch'] must be a C-contiguous float64 [H, T] array")
        if M.shape != (H, W) or M.dtype != np.uint64 or not M.flags.c_contiguous:
            raise ValueError("out['mask'] must be a C-contiguous uint64 [H, ceil(T/64)] array")
        # diff may be a column block of a wider [rows, all homes] array (rows strided, elements contiguous): the
        # library writes it in place (revs_get_schedule_ld)
        ld = H
        if D is not None:
            if D.ndim != 2 or D.shape[1] != H or D.dtype != np.float64 or (H > 1 and D.strides[1] != 8) or \
                    (D.shape[0] > 1 and (D.strides[0] % 8 or D.strides[0] < 8 * H)):
                raise ValueError("out['diff'] must be a float64 [rows, H] array with contiguous rows")
            ld = D.strides[0] // 8 if D.shape[0] > 1 else H
        dptr = _dp(D)
        _check(self.lib.revs_get_schedule_ld(self._h, _dp(P), M.ctypes.data_as(C.POINTER(C.c_uint64)), W, dptr,
                                             0 if D is None else D.shape[0], ld))
        return dict(P_sch=P, mask=M, diff=D)

    def estimate(self):
        H, T = self.H, self.T
        P, G = np.empty((H, T)), np.empty((H, T))
        _check(self.lib.revs_get_estimate(self._h, _dp(P), _dp(G)))
        return P, G

    def solve_individual(self):
        H, T = self.H, self.T
        P, E, S = np.empty((H, T)), np.empty((H, T)), np.empty((H, T + 1))
        _check(self.lib.revs_solve_individual(self._h, _dp(P), _dp(E), _dp(S)))
        return dict(P_res=P, P_ev=E, SOC=S)

    def reliability(self, feeder, kind, rows, vset=1.0, scale=None, P=None):
        rows = np.ascontiguousarray(rows, dtype=np.int32)
        out = np.empty((len(rows), self.T))
        sc = None if scale is None else _f64(scale, (len(rows),))
        Pp = None if P is None else _f64(P, (self.sizes[feeder], self.T))
        _check(self.lib.revs_reliability(self._h, feeder, kind, len(rows),
                                         rows.ctypes.data_as(C.POINTER(C.c_int32)), _dp(sc), vset,
                                         _dp(Pp), _dp(out)))
        return out

    # ---- the same check with the rows partitioned over the GPUs of the box; the all-gather of the result is fused
    # into the epilogue of the contraction kernel (peer-memory stores, include/revs_admm.h: revs_gather_*)
    def gather_export(self, capacity_doubles):
        buf = C.create_string_buffer(64)
        _check(self.lib.revs_gather_export(self._h, int(capacity_doubles), buf))
        return buf.raw

    def gather_attach(self, world, rank, handles):
        handles = bytes(handles)
        assert len(handles) == 64 * world
        _check(self.lib.revs_gather_attach(self._h, world, rank, handles))

    def reliability_sharded(self, feeder, kind, rows, P, vset=1.0, scale=None):
        rows = np.ascontiguousarray(rows, dtype=np.int32)
        out = np.empty((len(rows), self.T))
        sc = None if scale is None else _f64(scale, (len(rows),))
        Pp = _f64(P, (self.sizes[feeder], self.T))
        _check(self.lib.revs_reliability_sharded(self._h, feeder, kind, len(rows),
                                                 rows.ctypes.data_as(C.POINTER(C.c_int32)), _dp(sc), vset,
                                                 _dp(Pp), _dp(out)))
        return out

    # ---- residual all-reduce over the GPUs of the box (peer-memory mailboxes, see include/revs_admm.h)
    def comm_export(self):
        buf = C.create_string_buffer(64)
        _check(self.lib.revs_comm_export(self._h, buf))
        return buf.raw

    def comm_attach(self, world, rank, handles):
        handles = bytes(handles)
        assert len(handles) == 64 * world
        _check(self.lib.revs_comm_attach(self._h, world, rank, handles))

    def comm_detach(self):
        _check(self.lib.revs_comm_detach(self._h))

    def set_option(self, name, value):
        _check(self.lib.revs_set_option(self._h, name.encode(), float(value)))

    def stats(self):
        st = Stats()
        _check(self.lib.revs_get_stats(self._h, C.byref(st)))
        return st.as_dict()
