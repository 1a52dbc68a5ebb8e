"""TEST INFRASTRUCTURE -- CPU restatement of the reference's quasiseparable-matrix algebra.

NumPy restatement, point by point, of src/tinygp/solvers/quasisep/core.py and ops.py: the seven QSM types, their
element-wise sum / product, transpose, scaling, the QSM x QSM product (`qsm_mul`), the triangular and symmetric
inverses, `gram`, the Cholesky factorisation, the triangular solves and the dense products.  Only tests/, bench.py's
cpu_baseline leg and __graft_entry__.smoke() may import this module; the product (tinygp_b200/) never does.

Pinned against the unmodified reference (executed over tests/golden/jaxshim) by tests/golden/qsm_vectors.json
(generator: tests/golden/make_golden_qsm.py) in tests/test_oracle_qsm.py.

Representation: a `QSM` holds `d` (n,) or None, `lower` = (p, q, a) or None, `upper` = (p, q, a) or None and a `symm`
flag (SymmQSM: `upper` is None and means lower^T).  Strictly lower part (core.py:168-236):
    M[i, j] = p[i] . a[i-1] ... a[j+1] . q[j]   (i > j);
the strictly upper part with the same (p, q, a) is its transpose (core.py:239-292).
"""

from __future__ import annotations

import numpy as np


class QSM:
    def __init__(self, d=None, lower=None, upper=None, symm=False):
        self.d = None if d is None else np.asarray(d, dtype=np.float64)
        self.lower = None if lower is None else tuple(np.asarray(v, dtype=np.float64) for v in lower)
        self.upper = None if upper is None else tuple(np.asarray(v, dtype=np.float64) for v in upper)
        self.symm = bool(symm)
        if self.symm:
            assert self.d is not None and self.lower is not None and self.upper is None

    # ---- structure -----------------------------------------------------------------------------------------------
    @property
    def n(self):
        if self.d is not None:
            return self.d.shape[0]
        part = self.lower if self.lower is not None else self.upper
        return part[0].shape[0]

    def parts(self):
        """ops.py `deconstruct` (:217-229): diag, lower, upper with a SymmQSM's upper made explicit"""
        upper = self.lower if self.symm else self.upper
        return self.d, self.lower, upper

    @property
    def kind(self):
        d, lo, up = self.d, self.lower, self.upper
        if self.symm:
            return "symm"
        if lo is None and up is None:
            return "diag"
        if d is None:
            return "strict_lower" if up is None else ("strict_upper" if lo is None else "none")
        if up is None:
            return "lower"
        if lo is None:
            return "upper"
        return "square"

    def transpose(self):
        if self.symm:
            return self
        return QSM(self.d, self.upper, self.lower)              # core.py:186-187, 262-263, 295-296, 351-352, 403-408

    def __neg__(self):                                             # core.py: every __neg__ flips d and the p's
        neg = lambda part: None if part is None else (-part[0], part[1], part[2])
        return QSM(None if self.d is None else -self.d, neg(self.lower), neg(self.upper), self.symm)

    def scale(self, c):
        """core.py:155-156 (d * c), :196-197 (lower: p * c), :272-273 (upper: q * c)"""
        c = np.asarray(c, dtype=np.float64)
        cc = c[:, None] if c.ndim == 1 else c
        lo = None if self.lower is None else (self.lower[0] * cc, self.lower[1], self.lower[2])
        up = None if self.upper is None else (self.upper[0], self.upper[1] * cc, self.upper[2])
        return QSM(None if self.d is None else self.d * c, lo, up, self.symm)

    # ---- dense products (ops.py:308-349) -------------------------------------------------------------------------
    def matmul(self, x):
        x = np.asarray(x, dtype=np.float64)
        shape = x.shape
        x2 = x.reshape(shape[0], -1)                              # core.py:35-44
        d, lo, up = self.parts()
        out = np.zeros_like(x2)
        if d is not None:
            out += d[:, None] * x2
        if lo is not None:
            out += lower_matmul(*lo, x2)
        if up is not None:
            out += upper_matmul(*up, x2)
        return out.reshape(shape)

    def to_dense(self):                                            # core.py:84-90
        return self.matmul(np.eye(self.n))

    def __add__(self, other):
        return elementwise_add(self, other)

    def __sub__(self, other):
        return elementwise_add(self, -other)

    def __matmul__(self, other):
        if isinstance(other, QSM):
            return qsm_mul(self, other)
        return self.matmul(other)

    def __mul__(self, other):
        if isinstance(other, QSM):
            return elementwise_mul(self, other)
        return self.scale(other)

    # ---- factorisations ------------------------------------------------------------------------------------------
    def inv(self):
        k = self.kind
        if k == "diag":
            return QSM(1.0 / self.d)
        if k == "lower":
            return lower_inv(self)
        if k == "upper":
            return lower_inv(self.transpose()).transpose()       # core.py:362-363
        if k == "symm":
            lam, t, s, ell = symm_inv(self.d, *self.lower)
            return QSM(lam, (t, s, ell), symm=True)               # core.py:507-520
        if k == "square":
            return square_inv(self)
        raise ValueError("inv of a strictly triangular QSM")

    def gram(self):                                                # core.py:424-434
        M = qsm_mul(self.transpose(), self)
        return QSM(M.d, M.lower, symm=True)

    def cholesky(self):                                            # core.py:522-537
        assert self.symm
        c, w = cholesky(self.d, *self.lower)
        p, _, a = self.lower
        return QSM(c, (p, w, a))

    def solve(self, y):
        y = np.asarray(y, dtype=np.float64)
        shape = y.shape
        y2 = y.reshape(shape[0], -1)
        if self.kind == "lower":
            out = lower_solve(self.d, *self.lower, y2)
        elif self.kind == "upper":
            out = upper_solve(self.d, *self.upper, y2)
        else:
            raise ValueError("solve needs a triangular QSM")
        return out.reshape(shape)


# ---- scans over the points (ops.py:308-349, 352-365, 463-512) ------------------------------------------------------
def lower_matmul(p, q, a, x):                                     # ops.py:308-316
    n, m = q.shape
    f = np.zeros((m, x.shape[1]))
    out = np.empty_like(x)
    for k in range(n):
        out[k] = p[k] @ f
        f = a[k] @ f + np.outer(q[k], x[k])
    return out


def upper_matmul(p, q, a, x):                                     # ops.py:330-338
    n, m = p.shape
    f = np.zeros((m, x.shape[1]))
    out = np.empty_like(x)
    for k in range(n - 1, -1, -1):
        out[k] = q[k] @ f
        f = a[k].T @ f + np.outer(p[k], x[k])
    return out


def cholesky(d, p, q, a):                                         # ops.py:352-365
    n, m = q.shape
    f = np.zeros((m, m))
    c, w = np.empty(n), np.empty((n, m))
    with np.errstate(invalid="ignore"):
        for k in range(n):
            c[k] = np.sqrt(d[k] - p[k] @ f @ p[k])
            tmp = f @ a[k].T
            w[k] = (q[k] - p[k] @ tmp) / c[k]
            f = a[k] @ tmp + np.outer(w[k], w[k])
    return c, w


def lower_solve(d, p, q, a, x):                                   # ops.py:463-472
    n, m = q.shape
    f = np.zeros((m, x.shape[1]))
    out = np.empty_like(x)
    for k in range(n):
        out[k] = (x[k] - p[k] @ f) / d[k]
        f = a[k] @ f + np.outer(q[k], out[k])
    return out


def upper_solve(d, p, q, a, x):                                   # ops.py:489-498
    n, m = p.shape
    f = np.zeros((m, x.shape[1]))
    out = np.empty_like(x)
    for k in range(n - 1, -1, -1):
        out[k] = (x[k] - q[k] @ f) / d[k]
        f = a[k].T @ f + np.outer(p[k], out[k])
    return out


# ---- element-wise sum and product (ops.py:24-49, 232-296; core.py:158-165, 199-236, 275-284) -------------------------
def _construct(d, lower, upper, symm):                            # ops.py:232-268
    if lower is None and upper is None:
        return QSM(d)
    if symm:
        return QSM(d, lower, symm=True)
    return QSM(d, lower, upper)


def _block_diag(a1, a2):
    n, m1, m2 = a1.shape[0], a1.shape[1], a2.shape[1]
    out = np.zeros((n, m1 + m2, m1 + m2))
    out[:, :m1, :m1] = a1
    out[:, m1:, m1:] = a2
    return out


def _add_tri(x, y):                                               # core.py:199-216 (upper: via transposes, :275-277)
    if x is None:
        return y
    if y is None:
        return x
    return (np.concatenate((x[0], y[0]), axis=1), np.concatenate((x[1], y[1]), axis=1), _block_diag(x[2], y[2]))


def _mul_tri(x, y):                                               # core.py:218-233
    if x is None or y is None:
        return None
    m1, m2 = x[0].shape[1], y[0].shape[1]
    i, j = np.meshgrid(np.arange(m1), np.arange(m2))
    i, j = i.flatten(), j.flatten()
    return (x[0][:, i] * y[0][:, j], x[1][:, i] * y[1][:, j],
            x[2][:, i[:, None], i[None, :]] * y[2][:, j[:, None], j[None, :]])


def elementwise_add(a, b):                                        # ops.py:24-35
    da, la, ua = a.parts()
    db, lb, ub = b.parts()
    d = da if db is None else (db if da is None else da + db)
    symm = (a.symm or a.kind == "diag") and (b.symm or b.kind == "diag")
    return _construct(d, _add_tri(la, lb), _add_tri(ua, ub), symm)


def elementwise_mul(a, b):                                        # ops.py:38-49
    da, la, ua = a.parts()
    db, lb, ub = b.parts()
    d = None if (da is None or db is None) else da * db
    symm = (a.symm or a.kind == "diag") and (b.symm or b.kind == "diag")
    return _construct(d, _mul_tri(la, lb), _mul_tri(ua, ub), symm)


# ---- QSM x QSM (ops.py:52-214) -------------------------------------------------------------------------------------
def qsm_mul(A, B):
    da, la, ua = A.parts()
    db, lb, ub = B.parts()
    if la is None and ua is None and lb is None and ub is None:   # ops.py:57-60
        return QSM(da * db)
    n = A.n
    phi = psi = None
    if la is not None and ub is not None:                         # ops.py:62-72: phi_{k+1} = a phi b^T + q g^T
        f = np.zeros((la[1].shape[1], ub[1].shape[1]))
        phi = np.empty((n,) + f.shape)
        for k in range(n):
            phi[k] = f
            f = la[2][k] @ f @ ub[2][k].T + np.outer(la[1][k], ub[1][k])
    if ua is not None and lb is not None:                         # ops.py:77-87 (reverse): psi = a^T psi b + q g^T
        f = np.zeros((ua[0].shape[1], lb[0].shape[1]))
        psi = np.empty((n,) + f.shape)
        for k in range(n - 1, -1, -1):
            psi[k] = f
            f = ua[2][k].T @ f @ lb[2][k] + np.outer(ua[0][k], lb[0][k])

    lam = None
    out_lower = out_upper = None
    s_l, t_l, v_l, u_l, ell_l, del_l = [], [], [], [], [], []
    for k in range(n):                                            # ops.py:92-203, one point at a time
        alpha = beta = theta = eta = lamk = None
        if db is not None and la is not None:
            alpha = la[1][k] * db[k]
        if da is not None and lb is not None:
            beta = da[k] * lb[0][k]
        if da is not None and ub is not None:
            theta = da[k] * ub[1][k]
        if db is not None and ua is not None:
            eta = ua[0][k] * db[k]
        if da is not None and db is not None:
            lamk = da[k] * db[k]
        add = lambda x, y: y if x is None else x + y
        if la is not None and phi is not None and ub is not None:
            alpha = add(alpha, la[2][k] @ phi[k] @ ub[0][k])
            theta = add(theta, la[0][k] @ phi[k] @ ub[2][k].T)
            lamk = add(lamk, la[0][k] @ phi[k] @ ub[0][k])
        if ua is not None and psi is not None and lb is not None:
            beta = add(beta, ua[1][k] @ psi[k] @ lb[2][k])
            eta = add(eta, ua[2][k].T @ psi[k] @ lb[1][k])
            lamk = add(lamk, ua[1][k] @ psi[k] @ lb[1][k])
        s = ([alpha] if alpha is not None else []) + ([lb[1][k]] if lb is not None else [])
        t = ([la[0][k]] if la is not None else []) + ([beta] if beta is not None else [])
        v = ([ua[1][k]] if ua is not None else []) + ([theta] if theta is not None else [])
        u = ([eta] if eta is not None else []) + ([ub[0][k]] if ub is not None else [])
        if la is not None and lb is not None:
            m1, m2 = la[2].shape[1], lb[2].shape[1]
            ell = np.zeros((m1 + m2, m1 + m2))
            ell[:m1, :m1] = la[2][k]
            ell[:m1, m1:] = np.outer(la[1][k], lb[0][k])
            ell[m1:, m1:] = lb[2][k]
        else:
            ell = la[2][k] if la is not None else (lb[2][k] if lb is not None else None)
        if ua is not None and ub is not None:
            m1, m2 = ua[2].shape[1], ub[2].shape[1]
            delta = np.zeros((m1 + m2, m1 + m2))
            delta[:m1, :m1] = ua[2][k]
            delta[m1:, :m1] = np.outer(ub[1][k], ua[0][k])
            delta[m1:, m1:] = ub[2][k]
        else:
            delta = ua[2][k] if ua is not None else (ub[2][k] if ub is not None else None)
        lam = lam if lamk is None else (np.empty(n) if lam is None else lam)
        if lamk is not None:
            lam[k] = lamk
        if t and s and ell is not None:
            t_l.append(np.concatenate(t)); s_l.append(np.concatenate(s)); ell_l.append(ell)
        if u and v and delta is not None:
            u_l.append(np.concatenate(u)); v_l.append(np.concatenate(v)); del_l.append(delta)
    if t_l:
        out_lower = (np.array(t_l), np.array(s_l), np.array(ell_l))
    if u_l:
        out_upper = (np.array(u_l), np.array(v_l), np.array(del_l))
    symm = (A.symm or A.kind == "diag") and (B.symm or B.kind == "diag")
    return _construct(lam, out_lower, out_upper, symm)


# ---- inverses ------------------------------------------------------------------------------------------------------
def lower_inv(L):                                                 # core.py:310-317
    d = L.d
    p, q, a = L.lower
    g = 1.0 / d
    u = -g[:, None] * p
    v = g[:, None] * q
    b = a - v[:, :, None] * p[:, None, :]
    return QSM(g, (u, v, b))


def symm_inv(d, p, q, a):                                         # ops.py:403-429
    n, m = q.shape
    f = np.zeros((m, m))
    ig, s, ell = np.empty(n), np.empty((n, m)), np.empty((n, m, m))
    for k in range(n):
        fpk = f @ p[k]
        left = q[k] - a[k] @ fpk
        ig[k] = 1.0 / (d[k] - p[k] @ fpk)
        s[k] = ig[k] * left
        ell[k] = a[k] - np.outer(s[k], p[k])
        f = a[k] @ f @ a[k].T + ig[k] * np.outer(left, left)
    z = np.zeros((m, m))
    lam, t = np.empty(n), np.empty((n, m))
    for k in range(n - 1, -1, -1):
        zak = z @ a[k]
        skzak = s[k] @ zak
        lam[k] = ig[k] + s[k] @ z @ s[k]
        t[k] = skzak - lam[k] * p[k]
        z = a[k].T @ zak - np.outer(skzak, p[k]) - np.outer(p[k], t[k])
    return lam, t, s, ell


def square_inv(M):                                                # core.py:436-478
    d = M.d
    p, q, a = M.lower
    h, g, b = M.upper
    n = d.shape[0]
    f = np.zeros((q.shape[1], g.shape[1]))
    ig, s, ell, v, del_ = [], [], [], [], []
    for k in range(n):
        fhk = f @ h[k]
        fbk = f @ b[k].T
        left = q[k] - a[k] @ fhk
        right = g[k] - p[k] @ fbk
        igk = 1.0 / (d[k] - p[k] @ fhk)
        sk = igk * left
        vk = igk * right
        ig.append(igk); s.append(sk); ell.append(a[k] - np.outer(sk, p[k]))
        v.append(vk); del_.append(b[k] - np.outer(vk, h[k]))
        f = a[k] @ fbk + igk * np.outer(left, right)
    z = np.zeros((h.shape[1], p.shape[1]))
    lam, t, u = np.empty(n), np.empty((n, p.shape[1])), np.empty((n, h.shape[1]))
    for k in range(n - 1, -1, -1):
        zsk = z @ s[k]
        zak = z @ a[k]
        lk = ig[k] + v[k] @ zsk
        tk = v[k] @ zak - lk * p[k]
        uk = b[k].T @ zsk - lk * h[k]
        z = b[k].T @ zak - np.outer(uk + lk * h[k], p[k]) - np.outer(h[k], tk)
        lam[k], t[k], u[k] = lk, tk, uk
    return QSM(lam, (t, np.array(s), np.array(ell)), (u, np.array(v), np.array(del_)))


# ---- the conditioned covariance at the inputs (solvers/quasisep/solver.py:124-129) -----------------------------------
def condition_qsm(factor, M, noise_diag):
    """`factor`: LowerTriQSM of the training covariance; `M`: SymmQSM of the predictive kernel at the inputs"""
    delta = qsm_mul(factor.inv(), M).gram()
    Mn = elementwise_add(M, QSM(np.asarray(noise_diag, dtype=np.float64)))   # M += noise.to_qsm()  (noise.py:92-95)
    return Mn - delta
