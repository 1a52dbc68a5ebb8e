"""A stand-in for libb200gp.so built from the ORACLE -- TEST INFRASTRUCTURE ONLY.

Purpose: exercise the Python host layer of ``tinygp_b200`` (argument marshalling, shapes, which entry point is called
with what, noise placement, error behaviour) on a machine without a GPU, against the reference-generated goldens
(tests/test_host_layer_golden.py).  Nothing here can be reached from the product: it is installed by a pytest fixture
that swaps the process-wide ``_cabi`` context for one whose ``lib`` is a ``MockLib``, and removed afterwards.  It
re-states the C-ABI contract of include/b200gp.h in NumPy: every function takes exactly the ctypes objects the host
layer passes to the real library (``c_void_p`` addresses, ``byref`` out-parameters) and fills the same buffers.

The kernel-program interpreter (``eval_program``) is an independent NumPy restatement of ``kprog_eval`` /
``parse_prog`` (tinygp_b200/csrc/dense.cu) and therefore also checks the *lowering* (Kernel.lower_for, tau_program,
transforms) against the reference's kernel values.
"""

import ctypes

import numpy as np
import scipy.linalg as sla

from oracle import tinygp_np as o

OP_CONST, OP_EXP, OP_EXPSQ, OP_M32, OP_M52, OP_COS, OP_ESS, OP_RQ, OP_EXPCOS, OP_EXPSIN = range(10)
OP_ADD, OP_MUL, OP_METRIC = 16, 17, 32
QS_EXP, QS_MATERN32, QS_MATERN52, QS_SHO, QS_CELERITE, QS_COSINE, QS_CARMA2 = range(7)


def _addr(p):
    if p is None:
        return 0
    if isinstance(p, ctypes.c_void_p):
        return p.value or 0
    if isinstance(p, int):
        return p
    raise TypeError(type(p))


def arr(p, shape, dtype=np.float64):
    """view of the caller's buffer"""
    n = int(np.prod(shape))
    ct = {np.float64: ctypes.c_double, np.int64: ctypes.c_int64}[dtype]
    buf = (ct * n).from_address(_addr(p))
    return np.frombuffer(buf, dtype=dtype).reshape(shape)


def out(ref):
    """the object behind a byref()"""
    return ref._obj


# ---- kernel programs (include/b200gp.h; dense.cu:28-167) ---------------------------------------
def parse_program(p, n_rows, ndim):
    rows = arr(p, (n_rows, 4)).copy()
    metrics, r = [], 0
    while r < n_rows and int(rows[r, 0]) == OP_METRIC:
        mid, mr, mc = int(rows[r, 1]), int(rows[r, 2]), int(rows[r, 3])
        assert mid == len(metrics) + 1 and mc == ndim, "kernel program: bad metric header"
        nd = -(-mr * mc // 4)
        metrics.append(rows[r + 1:r + 1 + nd].ravel()[:mr * mc].reshape(mr, mc))
        r += 1 + nd
    return metrics, rows[r:]


def eval_program(metrics, instr, D):
    """D: (..., ndim) coordinate differences -> kernel values (...)"""
    st = []
    for op, dcode, p0, p1 in instr:
        op, dcode = int(op), int(dcode)
        if op == OP_ADD:
            b = st.pop(); st[-1] = st[-1] + b
            continue
        if op == OP_MUL:
            b = st.pop(); st[-1] = st[-1] * b
            continue
        if op == OP_CONST:
            st.append(np.full(D.shape[:-1], p0))
            continue
        Z = D if (dcode >> 1) == 0 else D @ metrics[(dcode >> 1) - 1].T
        l1, l2sq = np.sum(np.abs(Z), axis=-1), np.sum(Z * Z, axis=-1)
        l2 = bool(dcode & 1)
        if op in (OP_EXPSQ, OP_RQ):
            r2 = (l2sq if l2 else l1 * l1) / (p0 * p0)
            st.append(np.exp(-0.5 * r2) if op == OP_EXPSQ else (1.0 + 0.5 * r2 / p1) ** (-p1))
            continue
        dist = np.where(l2sq == 0.0, l1, np.sqrt(l2sq)) if l2 else l1
        if op == OP_EXPCOS:
            st.append(np.exp(-p0 * dist) * np.cos(p1 * dist)); continue
        if op == OP_EXPSIN:
            st.append(np.exp(-p0 * dist) * np.sin(p1 * dist)); continue
        r = dist / p0
        if op == OP_EXP:
            v = np.exp(-r)
        elif op == OP_M32:
            v = (1.0 + np.sqrt(3.0) * r) * np.exp(-np.sqrt(3.0) * r)
        elif op == OP_M52:
            a = np.sqrt(5.0) * r
            v = (1.0 + a + a * a / 3.0) * np.exp(-a)
        elif op == OP_COS:
            v = np.cos(2 * np.pi * r)
        elif op == OP_ESS:
            v = np.exp(-p1 * np.sin(np.pi * r) ** 2)
        else:
            raise ValueError(f"kernel program: unknown opcode {op}")
        st.append(v)
    assert len(st) == 1, "kernel program: malformed expression"
    return st[0]


def kmat(prog, X1, X2):
    metrics, instr = prog
    return eval_program(metrics, instr, X1[:, None, :] - X2[None, :, :])


def chol_or_nan(K):
    try:
        return sla.cholesky(K, lower=True, check_finite=False), 0
    except sla.LinAlgError:
        return np.full_like(K, np.nan), 1


# ---- quasiseparable components -> oracle kernels (tinygp_b200/kernels/quasisep.py components()) --------------
class _CarmaPair(o.qs.Quasisep):
    """B200GP_QS_CARMA2 (include/b200gp.h): one complex root pair of a CARMA process as its own state-space block"""

    def __init__(self, c, d, h1, h2, s):
        self.c, self.d, self.h, self.s = c, d, np.array([h1, h2]), s

    def stationary_covariance(self):
        r = self.c / self.d
        return np.array([[self.s, -r], [-r, self.s + 2 * r * r]])

    def observation_model(self, X):
        return self.h

    def transition_matrix(self, X1, X2):
        dt = X2 - X1
        e, cs, sn = np.exp(-self.c * dt), np.cos(self.d * dt), np.sin(self.d * dt)
        return e * np.array([[cs, sn], [-sn, cs]])


class _RateExp(o.qs.Exp):
    """B200GP_QS_EXP with the decay rate given directly (a real CARMA root): exp(-c dt), no division"""

    def __init__(self, rate, sigma):
        super().__init__(1.0 / rate, sigma)
        self.rate = rate

    def transition_matrix(self, X1, X2):
        return np.array([[np.exp(-self.rate * (X2 - X1))]])


def qs_kernel(comps):
    total, term = None, None
    for kind, scale, p0, p1, p2, p3, mul_next, p4 in comps[:, :8]:
        kind = int(kind)
        k = {QS_EXP: lambda: (_RateExp(p2, p1) if p2 != 0.0 else o.qs.Exp(p0, p1)), QS_MATERN32: lambda: o.qs.Matern32(p0, p1),
             QS_MATERN52: lambda: o.qs.Matern52(p0, p1), QS_SHO: lambda: o.qs.SHO(p0, p1, p2),
             QS_CELERITE: lambda: o.qs.Celerite(p0, p1, p2, p3), QS_COSINE: lambda: o.qs.Cosine(p0, p1),
             QS_CARMA2: lambda: _CarmaPair(p0, p1, p2, p3, p4)}[kind]()
        if scale != 1.0:
            k = o.qs.Scale(k, scale)
        term = k if term is None else o.qs.Product(term, k)     # mul_next chains a Product term (quasisep.py:298-331)
        if not mul_next:
            total = term if total is None else total + term
            term = None
    return total


class _Dense:
    pass


class MockLib:
    def __init__(self):
        self.objects, self.next_id, self.calls, self.err = {}, 100, [], b""

    def _new(self, obj, href):
        self.next_id += 1
        self.objects[self.next_id] = obj
        out(href).value = self.next_id
        return obj

    def _get(self, h):
        return self.objects[_addr(h)]

    def __getattr__(self, name):        # any entry point not re-stated here is a test failure, not a silent no-op
        if name.startswith("b200gp_qsm_"):
            # the QSM algebra is not re-stated: it runs from the HOST BUILD of the device source (tests/qsmhost.py)
            fn = getattr(self._qsm_host().lib, name)

            def call(*args):
                self.calls.append(name[len("b200gp_"):])
                return fn(*args)
            return call
        raise AttributeError(f"hostmock: {name} is not mocked")

    def _qsm_host(self):
        if "_qsmhost" not in self.__dict__:
            from qsmhost import HostBackend
            self.__dict__["_qsmhost"] = HostBackend()
        return self.__dict__["_qsmhost"]

    def b200gp_qsm_create(self, ctx, *args):
        self.calls.append("qsm_create")
        h = self._qsm_host()
        return h.lib.b200gp_qsm_create(h.handle, *args)

    def _qsm_from_arrays(self, kind, d, p, q, a, href):
        h = self._qsm_host()
        d, p, q, a = (np.ascontiguousarray(v, dtype=np.float64) for v in (d, p, q, a))
        P = lambda v: ctypes.c_void_p(v.ctypes.data)
        return h.lib.b200gp_qsm_create(h.handle, d.shape[0], kind, p.shape[1], 0, P(d), P(p), P(q), P(a), None, None, None, href)

    def b200gp_qs_kernel_qsm(self, ctx, comps, ncomp, t, n, href):      # kernels/quasisep.py:102-116 from the oracle
        self.calls.append("qs_kernel_qsm")
        k = qs_kernel(arr(comps, (ncomp, 8)).copy())
        d, p, q, a = o.qs_generators_fast(k, arr(t, (n,)).copy())
        return self._qsm_from_arrays(6, d, p, q, a, href)

    def b200gp_qs_factor_qsm(self, h, href):                            # solver.py:82 from the oracle's factor
        self.calls.append("qs_factor_qsm")
        s = self._get(h)
        _, p, _, a = o.qs_generators_fast(s.kernel, s.X)
        return self._qsm_from_arrays(3, s.c, p, s.w, a, href)

    def b200gp_last_error(self, ctx):
        h = self.__dict__.get("_qsmhost")
        if h is not None:
            msg = h.lib.b200gp_last_error(h.handle)
            if msg:
                return msg
        return self.err

    # -- context ------------------------------------------------------------------------------
    def b200gp_set_option(self, ctx, key, value):
        return 0

    # -- kernels ------------------------------------------------------------------------------
    def b200gp_kernel_matrix(self, ctx, prog, n_instr, X1, n1, X2, n2, ndim, outp):
        self.calls.append("kernel_matrix")
        P = parse_program(prog, n_instr, ndim)
        arr(outp, (n1, n2))[:] = kmat(P, arr(X1, (n1, ndim)), arr(X2, (n2, ndim)))
        return 0

    def b200gp_kernel_diag(self, ctx, prog, n_instr, X, n, ndim, outp):
        self.calls.append("kernel_diag")
        metrics, instr = parse_program(prog, n_instr, ndim)
        arr(outp, (n,))[:] = eval_program(metrics, instr, np.zeros((n, ndim)))
        return 0

    def b200gp_kernel_matvec(self, ctx, prog, n_instr, X1, n1, X2, n2, ndim, y, outp):
        self.calls.append("kernel_matvec")
        P = parse_program(prog, n_instr, ndim)
        arr(outp, (n1,))[:] = kmat(P, arr(X1, (n1, ndim)), arr(X2, (n2, ndim))) @ arr(y, (n2,))
        return 0

    # -- dense --------------------------------------------------------------------------------
    def b200gp_dense_create(self, ctx, prog, n_instr, X, n, ndim, diag, href, info):
        self.calls.append("dense_create")
        d = _Dense()
        d.prog, d.X, d.n = parse_program(prog, n_instr, ndim), arr(X, (n, ndim)).copy(), n
        d.K = kmat(d.prog, d.X, d.X) + np.diag(arr(diag, (n,)))
        d.L, bad = chol_or_nan(d.K)
        out(info).value = bad
        self._new(d, href)
        return 0

    def b200gp_dense_create_with_resid(self, ctx, prog, n_instr, X, n, ndim, diag, resid, href, info, sumsq):
        rc = self.b200gp_dense_create(ctx, prog, n_instr, X, n, ndim, diag, href, info)
        self.calls[-1] = "dense_create_with_resid"
        d = self.objects[out(href).value]
        r = arr(resid, (n,)).copy()
        with np.errstate(all="ignore"):
            if np.all(np.isfinite(d.L)):
                a = sla.solve_triangular(d.L, r, lower=True, check_finite=False)
                out(sumsq).value = float(np.sum(a * a))
            else:
                out(sumsq).value = float("nan")
        return rc

    def b200gp_dense_create_from_cov(self, ctx, cov, n, href, info):
        self.calls.append("dense_create_from_cov")
        d = _Dense()
        d.prog, d.X, d.n, d.K = None, None, n, arr(cov, (n, n)).copy()
        d.L, bad = chol_or_nan(d.K)
        out(info).value = bad
        self._new(d, href)
        return 0

    def b200gp_dense_free(self, h):
        self.objects.pop(_addr(h), None)
        return 0

    def b200gp_dense_logdet_half(self, h, ref):
        with np.errstate(all="ignore"):
            out(ref).value = float(np.sum(np.log(np.diag(self._get(h).L))))
        return 0

    def b200gp_dense_solve_triangular(self, h, Y, nrhs, transpose):
        d = self._get(h)
        y = arr(Y, (d.n, nrhs))
        if not np.all(np.isfinite(d.L)):
            y[:] = np.nan
        else:
            y[:] = sla.solve_triangular(d.L, y, lower=True, trans=1 if transpose else 0, check_finite=False)
        return 0

    def b200gp_dense_dot_triangular(self, h, Y, nrhs):
        d = self._get(h)
        y = arr(Y, (d.n, nrhs))
        y[:] = d.L @ y
        return 0

    def b200gp_dense_condition(self, h, prog, n_instr, Xtest, m, diag, outp):
        self.calls.append("dense_condition")
        d = self._get(h)
        ndim = d.X.shape[1]
        P = parse_program(prog, n_instr, ndim)
        Xt = d.X if _addr(Xtest) == 0 else arr(Xtest, (m, ndim))
        m = Xt.shape[0]
        A = sla.solve_triangular(d.L, kmat(P, d.X, Xt), lower=True, check_finite=False)
        arr(outp, (m, m))[:] = kmat(P, Xt, Xt) + np.diag(arr(diag, (m,))) - A.T @ A       # direct.py:88-95
        return 0

    def b200gp_gram_downdate(self, ctx, At, m, k, C):
        self.calls.append("gram_downdate")
        a = arr(At, (m, k))
        arr(C, (m, m))[:] -= a @ a.T
        return 0

    def b200gp_dense_covariance(self, h, outp):
        d = self._get(h)
        arr(outp, (d.n, d.n))[:] = d.K
        return 0

    def b200gp_dense_get_factor(self, h, outp):
        d = self._get(h)
        arr(outp, (d.n, d.n))[:] = d.L
        return 0

    # -- quasisep -----------------------------------------------------------------------------
    def b200gp_qs_check_sorted(self, ctx, t, n, unsorted):
        out(unsorted).value = int(np.any(np.diff(arr(t, (n,))) < 0.0))      # solver.py:142-146
        return 0

    def b200gp_qs_create(self, ctx, comps, ncomp, t, n, diag, assume_sorted, href, unsorted, info):
        self.calls.append("qs_create")
        tt = arr(t, (n,)).copy()
        out(unsorted).value = 0
        if not assume_sorted and np.any(np.diff(tt) < 0.0):
            out(unsorted).value = 1
            return 0
        k = qs_kernel(arr(comps, (ncomp, 8)).copy())
        s = o.QuasisepSolver(k, tt, o.Diagonal(arr(diag, (n,)).copy()), assume_sorted=True)
        out(info).value = 0 if np.all(np.isfinite(s.c)) else 1
        self._new(s, href)
        return 0

    def b200gp_qs_free(self, h):
        self.objects.pop(_addr(h), None)
        return 0

    def b200gp_qs_state_dim(self, h, ref):
        out(ref).value = self._get(h).w.shape[1]
        return 0

    def b200gp_qs_logdet_half(self, h, ref):
        out(ref).value = float(np.sum(np.log(self._get(h).c)))
        return 0

    def b200gp_qs_variance(self, h, outp):
        s = self._get(h)
        arr(outp, s.d.shape)[:] = s.d
        return 0

    def b200gp_qs_get_factor(self, h, c, w):
        s = self._get(h)
        arr(c, s.c.shape)[:] = s.c
        arr(w, s.w.shape)[:] = s.w
        return 0

    def b200gp_qs_solve_triangular(self, h, Y, nrhs, transpose):
        s = self._get(h)
        y = arr(Y, (s.d.shape[0], nrhs))
        y[:] = s.solve_triangular(y.copy(), transpose=bool(transpose))
        return 0

    def b200gp_qs_solve_sumsq(self, h, Y, outp):
        s = self._get(h)
        y = arr(Y, (s.d.shape[0],))
        outp._obj.value = float(np.sum(s.solve_triangular(y.copy()) ** 2))
        return 0

    def b200gp_qs_dot_triangular(self, h, Y, nrhs):
        s = self._get(h)
        y = arr(Y, (s.d.shape[0], nrhs))
        y[:] = s.dot_triangular(y.copy())
        return 0

    def b200gp_qs_matmul(self, h, Y, nrhs):
        s = self._get(h)
        y = arr(Y, (s.d.shape[0], nrhs))
        y[:] = s.covariance() @ y
        return 0

    def b200gp_qs_inverse_diagonal(self, h, outp):
        self.calls.append("qs_inverse_diagonal")
        s = self._get(h)
        arr(outp, s.d.shape)[:] = np.diag(np.linalg.inv(s.covariance()))
        return 0

    def b200gp_qs_conditioned_variance(self, h, noise_pred, outp):
        self.calls.append("qs_conditioned_variance")
        s = self._get(h)
        n = s.d.shape[0]
        K = s.kernel(s.X, s.X)
        A = s.solve_triangular(K)
        arr(outp, (n,))[:] = arr(noise_pred, (n,)) + np.diag(K - A.T @ A)      # solver.py:124-129 read by :84-85
        return 0

    def b200gp_qs_condition(self, h, prog, n_instr, t_test, m, diag, outp):
        self.calls.append("qs_condition")
        s = self._get(h)
        P = parse_program(prog, n_instr, 1)
        X = s.X[:, None]
        Xt = X if _addr(t_test) == 0 else arr(t_test, (m, 1))
        m = Xt.shape[0]
        A = s.solve_triangular(kmat(P, X, Xt))
        res = kmat(P, Xt, Xt) - A.T @ A
        if _addr(diag) != 0:
            res = res + np.diag(arr(diag, (m,)))
        arr(outp, (m, m))[:] = res
        return 0

    def b200gp_searchsorted_right_m1(self, ctx, sorted_, n, query, m, outp):      # kernels/quasisep.py:121
        arr(outp, (m,), np.int64)[:] = np.searchsorted(arr(sorted_, (n,)), arr(query, (m,)), side="right") - 1
        return 0

    def b200gp_qs_kernel_matmul(self, ctx, comps, ncomp, t_test, m, t_train, n, Y, nrhs, outp):
        self.calls.append("qs_kernel_matmul")
        k = qs_kernel(arr(comps, (ncomp, 8)).copy())
        arr(outp, (m, nrhs))[:] = k(arr(t_test, (m,)), arr(t_train, (n,))) @ arr(Y, (n, nrhs))
        return 0
