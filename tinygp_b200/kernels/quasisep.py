"""Quasiseparable (celerite-style) kernels (reference: src/tinygp/kernels/quasisep.py).

Each kernel is a parameter holder that lowers to a list of state-space *components*
(include/b200gp.h, B200GP_QS_*): a ``Sum`` is block diagonal (quasisep.py:241-295) so it is the
concatenation of its terms' components, a ``Scale`` multiplies the stationary covariance
(quasisep.py:334-340).  The per-point generators d, p, q, a (``to_symm_qsm``, quasisep.py:102-116)
are computed on the device in registers from t[k] - t[k-1]; they are never materialised unless
``QuasisepSolver.generators()`` asks for them.
"""

from __future__ import annotations

__all__ = ["Quasisep", "Wrapper", "Sum", "Product", "Scale", "Celerite", "SHO", "Exp", "Matern32", "Matern52",
           "Cosine", "CARMA", "carma_roots", "carma_quads2poly", "carma_poly2quads", "carma_acvf"]

import numpy as np

from tinygp_b200.kernels.base import Kernel

# dense kernel-program opcodes (include/b200gp.h)
OP_CONST, OP_EXP, OP_MATERN32, OP_MATERN52, OP_COSINE, OP_EXPCOS, OP_EXPSIN, OP_ADD, OP_MUL = 0, 1, 3, 4, 5, 8, 9, 16, 17

QS_EXP, QS_MATERN32, QS_MATERN52, QS_SHO, QS_CELERITE, QS_COSINE, QS_CARMA2 = range(7)
STATE_DIM = {QS_EXP: 1, QS_MATERN32: 2, QS_MATERN52: 3, QS_SHO: 2, QS_CELERITE: 2, QS_COSINE: 2, QS_CARMA2: 2}


class Quasisep(Kernel):
    """Base class (quasisep.py:47-215)."""

    def components(self) -> list[tuple]:
        """Rows ``(kind, pinf_scale, p0, p1, p2, p3, mul_next)`` for the device.  A row with ``mul_next = 1`` is
        multiplied (Kronecker-structured state, quasisep.py:298-331) with the row after it; a *term* is a maximal run of
        such rows, and the model is the sum (block-diagonal state) of its terms."""
        raise NotImplementedError(
            f"{type(self).__name__} is unsupported by the B200 quasiseparable solver backend")

    def component_array(self) -> np.ndarray:
        comps = self.components()
        if len(comps) > 8 or self.state_dim() > 8:      # B200GP_QS_MAX_COMP / B200GP_QS_MAX_J (include/b200gp.h)
            raise NotImplementedError(f"a quasiseparable model with {len(comps)} component rows and {self.state_dim()} states "
                                      "(more than 8 of either) is unsupported by the B200 quasiseparable solver backend")
        out = np.zeros((len(comps), 8))
        for i, c in enumerate(comps):
            out[i, : len(c)] = c
        return out

    def state_dim(self) -> int:
        total, term = 0, 1
        for c in self.components():
            term *= STATE_DIM[int(c[0])]
            if not (len(c) > 6 and c[6]):
                total += term
                term = 1
        return total

    def coord_to_sortable(self, X):
        """quasisep.py:88-100: the sortable (time) coordinate of ONE input point; the identity unless a subclass overrides it"""
        return X

    def _sortable(self, X):
        """coord_to_sortable over all points of X (the reference vmaps it): the array itself for the default identity -- also
        through Sum / Product / Scale, which forward to their first kernel --, a per-point call of a user's override otherwise"""
        fn = type(self).coord_to_sortable
        if fn is Quasisep.coord_to_sortable:
            return X
        if fn in _FORWARDED_COORDS:
            return (self.kernel1 if hasattr(self, "kernel1") else self.kernel)._sortable(X)
        return np.asarray([self.coord_to_sortable(x) for x in np.asarray(X)], dtype=np.float64)

    def _on_device(self) -> bool:
        """does the state-space model lower to the device's component rows (a built-in model with at most 8 states)?  If not --
        a user-defined subclass that writes design_matrix / stationary_covariance / observation_model / transition_matrix in
        Python (quasisep.py:60-100), or a model with more states -- the generators are evaluated point by point on the host
        and uploaded, and the device works on them as generator arrays of any order (solvers/quasisep/core.py, csrc/qsm.cu)."""
        try:
            self.component_array()
        except NotImplementedError:
            return False
        return True

    def _has_closed_form(self) -> bool:
        try:
            self.tau_program(0)
        except NotImplementedError:
            return False
        return True

    def _host_symm_qsm(self, X):
        """quasisep.py:102-116 as written, with the model's own Python methods per point (X sorted by coord_to_sortable)"""
        from tinygp_b200.solvers.quasisep import core
        from tinygp_b200.solvers.quasisep.block import ensure_dense
        X = np.asarray(X, dtype=np.float64)
        n = X.shape[0]
        try:
            self.components()
        except NotImplementedError:
            pass
        else:
            # a built-in model (just more states than the model-specialised kernels compile): the same formulas for all points
            # at once -- the component rows' closed-form transition matrices over the array of time steps
            t = np.asarray(self._sortable(X), dtype=np.float64)
            a = np.transpose(self._assemble("T", np.diff(t, prepend=t[:1])), (2, 1, 0))     # a_k = T(t_k - t_{k-1})^T, a_0 = I
            h = self._assemble("h")
            hP = h @ self._assemble("P")
            return core.SymmQSM(diag=core.DiagQSM(d=np.full(n, hP @ h)),
                                lower=core.StrictLowerTriQSM(p=np.einsum("i,nij->nj", h, a), q=np.tile(hP, (n, 1)),
                                                             a=np.ascontiguousarray(a)))
        Pinf = np.asarray(ensure_dense(self.stationary_covariance()), dtype=np.float64)
        h = np.stack([np.asarray(self.observation_model(X[k]), dtype=np.float64) for k in range(n)])
        a = np.stack([np.asarray(ensure_dense(self.transition_matrix(X[max(k - 1, 0)], X[k])), dtype=np.float64).T
                      for k in range(n)])                                        # a_0 = T(x_0, x_0)^T
        hP = h @ Pinf
        return core.SymmQSM(diag=core.DiagQSM(d=np.sum(hP * h, axis=1)),
                            lower=core.StrictLowerTriQSM(p=np.einsum("ni,nij->nj", h, a), q=hP, a=a))

    def _host_joint(self, X1, X2):
        """the SymmQSM of the union of two coordinate sets in sorted order + where each set sits in it: cross-covariances and
        their products with vectors (quasisep.py:118-163) for models without a device lowering, by the symmetric device algebra"""
        X1, X2 = np.asarray(X1, dtype=np.float64), np.asarray(X2, dtype=np.float64)
        joint = np.concatenate((X1, X2), axis=0)
        order = np.argsort(np.asarray(self._sortable(joint), dtype=np.float64), kind="stable")
        rank = np.empty_like(order)
        rank[order] = np.arange(order.size)
        return self._host_symm_qsm(joint[order]), order, rank[: X1.shape[0]], rank[X1.shape[0]:]

    def _host_dense(self, X1, X2=None):
        from tinygp_b200.solvers.quasisep.block import ensure_dense
        X1 = np.asarray(X1, dtype=np.float64)
        if X2 is None:                                                           # evaluate_diag, quasisep.py:212-215
            Pinf = np.asarray(ensure_dense(self.stationary_covariance()), dtype=np.float64)
            h = np.stack([np.asarray(self.observation_model(x), dtype=np.float64) for x in X1])
            return np.sum((h @ Pinf) * h, axis=1)
        S, _, r1, r2 = self._host_joint(X1, X2)
        return S.to_dense()[np.ix_(r1, r2)]

    # ---- the state-space model on the host (quasisep.py:60-100: design_matrix, stationary_covariance, observation_model,
    # transition_matrix).  Small J x J NumPy matrices assembled from the SAME component rows the device lowers
    # (`_leaf_state_space` restates qs_core.cuh's build_model / qs_leaf_transition); nothing on the solver path calls them.
    def _terms(self):
        """[[leaf rows of one term], ...]: a term is a run of rows chained by mul_next (a Product); terms are summed"""
        terms, cur = [], []
        for r in self.components():
            r = tuple(r) + (0.0,) * (8 - len(r))
            cur.append(r)
            if not r[6]:
                terms.append(cur)
                cur = []
        return terms

    def _assemble(self, what, *args):
        blocks = []
        for term in self._terms():
            mats = [_leaf_state_space(r, what, *args) for r in term]
            m = mats[0]
            for nxt in mats[1:]:
                if what == "F":      # F = F1 (x) I + I (x) F2 in the product's state order (quasisep.py:307-311)
                    m = _prod(m, np.eye(nxt.shape[0])) + _prod(np.eye(m.shape[0]), nxt)
                else:
                    m = _prod(m, nxt)
            blocks.append(m)
        if blocks[0].ndim == 1:
            return np.concatenate(blocks)
        n = sum(b.shape[0] for b in blocks)
        out, o = np.zeros((n, n) + blocks[0].shape[2:]), 0      # trailing axis: "T" for an array of time steps
        for b in blocks:
            out[o:o + b.shape[0], o:o + b.shape[0]] = b
            o += b.shape[0]
        return out

    def design_matrix(self):
        return self._assemble("F")

    def stationary_covariance(self):
        return self._assemble("P")

    def observation_model(self, X):
        return self._assemble("h")

    def transition_matrix(self, X1, X2):
        """the ADJOINT transition matrix between two coordinates, as in the reference (quasisep.py:88-95)"""
        return self._assemble("T", float(X2) - float(X1))

    def to_symm_qsm(self, X):
        """quasisep.py:102-116: the SymmQSM of this kernel at the sorted coordinates X, generated on the device (d, p, q,
        a never touch the host) -> ``tinygp_b200.solvers.quasisep.core.SymmQSM``."""
        from ctypes import byref, c_void_p

        from tinygp_b200 import _cabi
        from tinygp_b200.solvers.quasisep import core
        if not self._on_device():
            return self._host_symm_qsm(X)
        t = _cabi.f64(np.asarray(self._sortable(X), dtype=np.float64))
        if t.ndim != 1:
            raise ValueError("quasiseparable kernels take 1-D sortable coordinates")
        comps = self.component_array()
        ctx = core._backend()
        out = c_void_p()
        ctx.check(ctx.lib.b200gp_qs_kernel_qsm(ctx.handle, _cabi.ptr(comps), comps.shape[0], _cabi.ptr(t), t.shape[0],
                                               byref(out)))
        return core.QSM._wrap(ctx, out)

    # Dense evaluation (quasisep.py:118-145 ``evaluate``; used by ``condition`` at test points,
    # solver.py:131-139): the closed form k(tau), tau = |t1 - t2|, lowered to the same device kernel program as
    # the stationary kernels and evaluated by the CUDA build kernel (Kernel.__call__).
    def tau_program(self, dist: int) -> list[tuple]:
        raise NotImplementedError(
            f"{type(self).__name__} has no dense closed form on the B200 backend")

    def lower(self, lc):
        return self.tau_program(lc.metric_code())   # distance code 0 = L1 = |t1 - t2| in one dimension

    def __call__(self, X1, X2=None):
        if not self._has_closed_form():
            return self._host_dense(X1, X2)
        return self._dense(self._sortable(X1), None if X2 is None else self._sortable(X2))

    def _dense(self, t1, t2=None):
        """k(t1, t2) for coordinates that are already sortable (coord_to_sortable applied): the CUDA build kernel"""
        t1 = np.asarray(t1, dtype=np.float64)
        if t1.ndim != 1 or (t2 is not None and np.ndim(t2) != 1):
            raise ValueError("quasiseparable kernels take 1-D sortable coordinates")
        return super().__call__(t1, t2)

    def to_general_qsm(self, X1, X2):
        """quasisep.py:118-145: the rectangular cross-covariance as a handle whose ``@ y`` runs on the device"""
        from tinygp_b200.solvers.quasisep.general import GeneralQSM
        return GeneralQSM(self, X1, X2)

    def evaluate(self, X1, X2):
        """Scalar evaluation k(t1, t2) (kernels/quasisep.py:118-145): 1-element 1-D coordinates, not the (1, 1) arrays the
        stationary base class builds."""
        if not self._has_closed_form():                                          # quasisep.py:201-210 as written
            from tinygp_b200.solvers.quasisep.block import ensure_dense
            if self.coord_to_sortable(X1) >= self.coord_to_sortable(X2):
                X1, X2 = X2, X1
            h1, h2 = self.observation_model(X1), self.observation_model(X2)
            return float(h2 @ ensure_dense(self.transition_matrix(X1, X2)).T @ ensure_dense(self.stationary_covariance()) @ h1)
        t1 = np.atleast_1d(np.asarray(X1, dtype=np.float64)).reshape(-1)[:1]
        t2 = np.atleast_1d(np.asarray(X2, dtype=np.float64)).reshape(-1)[:1]
        return self(t1, t2)[0, 0]

    def evaluate_diag(self, X):
        return self.evaluate(X, X)

    def matmul(self, X1, X2=None, y=None):
        """quasisep.py:147-163: ``to_general_qsm(X1, X2) @ y`` in O((n + m) J^2) on the device (two state scans
        over the sorted X2, a searchsorted and two transition matrices per row of X1).  With X2 omitted the
        general form is used with X1 for both (reference: ``to_symm_qsm(X1) @ y``, the same matrix)."""
        from tinygp_b200 import _cabi
        if y is None:
            y, X2 = X2, None
        if X2 is None:
            X2 = X1
        if not self._on_device():
            # no device model: K(X1, X2) y = rows X1 of  S [y at X2; 0 at X1]  with S the SymmQSM of the sorted union --
            # O((n + m) J^2) in the device's symmetric QSM product, nothing densified
            y = np.asarray(y, dtype=np.float64)
            S, order, r1, r2 = self._host_joint(X1, X2)
            if y.shape[0] != r2.shape[0]:
                raise ValueError("dimension mismatch")
            v = np.zeros((order.size,) + y.shape[1:])
            v[r2] = y
            return (S @ v)[r1]
        t1 = _cabi.f64(np.asarray(self._sortable(X1), dtype=np.float64))
        t2 = _cabi.f64(np.asarray(self._sortable(X2), dtype=np.float64))
        y = np.asarray(y, dtype=np.float64)
        if t1.ndim != 1 or t2.ndim != 1:
            raise ValueError("quasiseparable kernels take 1-D sortable coordinates")
        if y.shape[0] != t2.shape[0]:
            raise ValueError("dimension mismatch")
        comps = self.component_array()
        yy = np.ascontiguousarray(y.reshape(t2.shape[0], -1))
        out = np.empty((t1.shape[0], yy.shape[1]))
        ctx = _cabi.get_context()
        ctx.check(ctx.lib.b200gp_qs_kernel_matmul(ctx.handle, _cabi.ptr(comps), comps.shape[0], _cabi.ptr(t1),
                                                  t1.shape[0], _cabi.ptr(t2), t2.shape[0], _cabi.ptr(yy),
                                                  yy.shape[1], _cabi.ptr(out)))
        return out.reshape((t1.shape[0],) + y.shape[1:])

    # algebra (quasisep.py:165-199)
    def __add__(self, other):
        if not isinstance(other, Quasisep):
            raise ValueError("Quasisep kernels can only be added to other Quasisep kernels")
        return Sum(self, other)

    def __radd__(self, other):
        if not isinstance(other, Kernel) and np.ndim(other) == 0 and other == 0:
            return self
        if not isinstance(other, Quasisep):
            raise ValueError("Quasisep kernels can only be added to other Quasisep kernels")
        return Sum(other, self)

    def __mul__(self, other):
        if isinstance(other, Quasisep):
            return Product(self, other)
        if isinstance(other, Kernel) or np.ndim(other) != 0:
            raise ValueError("Quasisep kernels can only be multiplied by scalars and other Quasisep kernels")
        return Scale(kernel=self, scale=other)

    def __rmul__(self, other):
        if isinstance(other, Quasisep):
            return Product(other, self)
        if isinstance(other, Kernel) or np.ndim(other) != 0:
            raise ValueError("Quasisep kernels can only be multiplied by scalars and other Quasisep kernels")
        return Scale(kernel=self, scale=other)


def _prod(a1, a2):
    """`_prod_helper` (quasisep.py:676-687): Kronecker structure with the FIRST factor's index running fastest"""
    a1, a2 = np.asarray(a1, dtype=np.float64), np.asarray(a2, dtype=np.float64)
    i, j = np.meshgrid(np.arange(a1.shape[0]), np.arange(a2.shape[0]))
    i, j = i.flatten(), j.flatten()
    if a1.ndim == 1:
        return a1[i] * a2[j]
    return a1[i[:, None], i[None, :]] * a2[j[:, None], j[None, :]]


def _leaf_state_space(row, what, dt=None):
    """F (design matrix), P (stationary covariance incl. the row's Pinf scale), h (observation model) or T(dt) (adjoint
    transition matrix) of ONE component row -- the host twin of build_model / qs_leaf_transition in csrc/qs_core.cuh"""
    kind, ps, p0, p1, p2, p3, _, p4 = row[:8]
    kind = int(kind)
    if kind == QS_EXP:                                          # quasisep.py:491-525 (p2: a CARMA real root's rate)
        rate = p2 if p2 != 0.0 else 1.0 / p0
        return {"F": lambda: np.array([[-rate]]), "P": lambda: ps * np.ones((1, 1)), "h": lambda: np.array([p1]),
                "T": lambda: np.array([[np.exp(-rate * dt)]])}[what]()
    if kind == QS_MATERN32:                                     # quasisep.py:528-569
        f = np.sqrt(3.0) / p0
        return {"F": lambda: np.array([[0.0, 1.0], [-f * f, -2 * f]]), "P": lambda: ps * np.diag([1.0, 3.0 / (p0 * p0)]),
                "h": lambda: np.array([p1, 0.0]),
                "T": lambda: np.exp(-f * dt) * np.array([[1 + f * dt, -f * f * dt], [dt, 1 - f * dt]])}[what]()
    if kind == QS_MATERN52:                                     # quasisep.py:572-633
        f = np.sqrt(5.0) / p0
        f2 = f * f
        if what == "F":
            return np.array([[0.0, 1.0, 0.0], [0.0, 0.0, 1.0], [-f2 * f, -3 * f2, -3 * f]])
        if what == "P":
            return ps * np.array([[1.0, 0.0, -f2 / 3], [0.0, f2 / 3, 0.0], [-f2 / 3, 0.0, f2 * f2]])
        if what == "h":
            return np.array([p1, 0.0, 0.0])
        d2 = dt * dt
        return np.exp(-f * dt) * np.array([
            [0.5 * f2 * d2 + f * dt + 1, -0.5 * f * f2 * d2, 0.5 * f2 * f * dt * (f * dt - 2)],
            [dt * (f * dt + 1), -f2 * d2 + f * dt + 1, f2 * dt * (f * dt - 3)],
            [0.5 * d2, 0.5 * dt * (2 - f * dt), 0.5 * f2 * d2 - 2 * f * dt + 1]])
    if kind == QS_SHO:                                          # quasisep.py:404-488: p0 = omega, p1 = quality, p2 = sigma
        w, q = p0, p1
        if what == "F":
            return np.array([[0.0, 1.0], [-w * w, -w / q]])
        if what == "P":
            return ps * np.diag([1.0, w * w])
        if what == "h":
            return np.array([p2, 0.0])
        if np.allclose(q, 0.5):
            return np.exp(-w * dt) * np.array([[1 + w * dt, -w * w * dt], [dt, 1 - w * dt]])
        under = q > 0.5
        f = np.sqrt(max(4 * q * q - 1, 0.0)) if under else np.sqrt(max(1 - 4 * q * q, 0.0))
        arg = 0.5 * f * w * dt / q
        sn, cs = (np.sin(arg), np.cos(arg)) if under else (np.sinh(arg), np.cosh(arg))
        return np.exp(-0.5 * w * dt / q) * np.array([[cs + sn / f, -2 * q * w * sn / f], [2 * q * sn / (w * f), cs - sn / f]])
    if kind in (QS_CELERITE, QS_CARMA2):                        # quasisep.py:343-401 / one complex CARMA pair (:770-900)
        if kind == QS_CELERITE:
            a, b, c, d = p0, p1, p2, p3
            c2, d2 = c * c, d * d
            h2_2 = d2 * (a * c - b * d) / (2 * c * (c2 + d2))
            h2 = np.sqrt(h2_2)
            h = np.array([(c * h2 - np.sqrt(a * d2 - (c2 + d2) * h2_2)) / d, h2])
            sgn = 1.0
        else:
            c, d, h, sgn = p0, p1, np.array([p2, p3]), p4
        if what == "F":
            return np.array([[-c, -d], [d, -c]])
        if what == "P":
            return ps * np.array([[sgn, -c / d], [-c / d, sgn + 2 * (c / d) ** 2]])
        if what == "h":
            return h
        cs, sn = np.cos(d * dt), np.sin(d * dt)
        return np.exp(-c * dt) * np.array([[cs, sn], [-sn, cs]])
    if kind == QS_COSINE:                                       # quasisep.py:636-673
        f = 2 * np.pi / p0
        if what == "F":
            return np.array([[0.0, -f], [f, 0.0]])
        if what == "P":
            return ps * np.eye(2)
        if what == "h":
            return np.array([p1, 0.0])
        cs, sn = np.cos(f * dt), np.sin(f * dt)
        return np.array([[cs, sn], [-sn, cs]])
    raise ValueError(f"unknown quasiseparable component kind {kind}")


class Sum(Quasisep):
    """quasisep.py:241-295.  ``use_block`` (:247-257): the host-side state-space matrices of a sum come back as
    ``solvers.quasisep.block.Block`` (default) or dense; the device always works on the block structure."""

    def __init__(self, kernel1: Quasisep, kernel2: Quasisep, use_block: bool = True):
        self.kernel1, self.kernel2, self.use_block = kernel1, kernel2, use_block

    def _blocked(self, m1, m2):      # quasisep.py:262-270 (Block objects are not nested)
        from tinygp_b200.solvers.quasisep.block import Block, ensure_dense
        if not self.use_block:
            return Block(ensure_dense(m1), ensure_dense(m2)).to_dense()
        return Block(*(m1.blocks if isinstance(m1, Block) else (m1,)), *(m2.blocks if isinstance(m2, Block) else (m2,)))

    def design_matrix(self):
        return self._blocked(self.kernel1.design_matrix(), self.kernel2.design_matrix())

    def stationary_covariance(self):
        return self._blocked(self.kernel1.stationary_covariance(), self.kernel2.stationary_covariance())

    def transition_matrix(self, X1, X2):
        return self._blocked(self.kernel1.transition_matrix(X1, X2), self.kernel2.transition_matrix(X1, X2))

    def observation_model(self, X):  # quasisep.py:283-289
        return np.concatenate((self.kernel1.observation_model(X), self.kernel2.observation_model(X)))

    def coord_to_sortable(self, X):  # quasisep.py:259-260
        return self.kernel1.coord_to_sortable(X)

    def components(self):
        return self.kernel1.components() + self.kernel2.components()

    def tau_program(self, dist):
        return self.kernel1.tau_program(dist) + self.kernel2.tau_program(dist) + [(OP_ADD, 0, 0.0, 0.0)]


class Wrapper(Quasisep):
    """quasisep.py:218-238: a base class for kernels that wrap another quasiseparable kernel.  Everything is forwarded to
    ``self.kernel``; a subclass may override ``coord_to_sortable`` (e.g. to pick the time column of structured inputs) and keep
    the device model.  A subclass that overrides a state-space method -- e.g. a coordinate-dependent ``observation_model``, the
    reference's multiband tutorial -- has no device rows (their h is a constant of the kernel): its generators are evaluated by
    its Python methods and the device works on them as generator arrays (Quasisep._on_device)."""

    def __init__(self, kernel: Quasisep):
        self.kernel = kernel

    def coord_to_sortable(self, X):
        return self.kernel.coord_to_sortable(X)

    def _plain(self):
        """does the wrapper leave the wrapped model as it is (only the coordinate mapping may differ)?  A subclass that
        overrides one of the state-space methods has no device rows and no closed form: Quasisep._on_device() is then False
        and the generators come from its Python methods"""
        for m in ("design_matrix", "stationary_covariance", "observation_model", "transition_matrix"):
            if getattr(type(self), m) is not getattr(Wrapper, m):
                raise NotImplementedError(f"{type(self).__name__} overrides {m}: no device lowering for a user-defined "
                                          "state-space method")

    def components(self):
        self._plain()
        return self.kernel.components()

    def tau_program(self, dist):
        self._plain()
        return self.kernel.tau_program(dist)

    def design_matrix(self):
        return self.kernel.design_matrix()

    def stationary_covariance(self):
        return self.kernel.stationary_covariance()

    def observation_model(self, X):
        return self.kernel.observation_model(self.coord_to_sortable(X))

    def transition_matrix(self, X1, X2):
        return self.kernel.transition_matrix(self.coord_to_sortable(X1), self.coord_to_sortable(X2))


class Scale(Wrapper):
    """quasisep.py:334-340"""

    def __init__(self, kernel: Quasisep, scale):
        self.kernel, self.scale = kernel, scale

    def _plain(self):
        pass

    def stationary_covariance(self):
        return self.scale * self.kernel.stationary_covariance()

    def components(self):
        s = float(self.scale)
        out, term_start = [], True
        for c in self.kernel.components():      # the stationary covariance of every TERM is scaled once (quasisep.py:339-340)
            out.append(((c[0], c[1] * s) + tuple(c[2:])) if term_start else tuple(c))
            term_start = not (len(c) > 6 and c[6])
        return out

    def tau_program(self, dist):
        return self.kernel.tau_program(dist) + [(OP_CONST, 0, float(self.scale), 0.0), (OP_MUL, 0, 0.0, 0.0)]


class Product(Quasisep):
    """quasisep.py:298-331: the state of the product is Kronecker-structured (``_prod_helper``, :676-687, first kernel's
    index fastest).  The device model chains the leaves of ONE term; a Sum inside a Product is multiplied out on the host,
    (a + b) * c == a * c + b * c, which gives the same kernel with the reference's state vector permuted (block of a * c,
    then block of b * c, instead of the interleaved Kronecker order) -- every solver result is unchanged by that."""

    def __init__(self, kernel1, kernel2):
        self.kernel1, self.kernel2 = kernel1, kernel2

    def components(self):
        def terms(rows):
            out, cur = [], []
            for r in rows:
                r = tuple(r) + (0.0,) * (7 - len(r))
                cur.append(r)
                if not r[6]:
                    out.append(cur)
                    cur = []
            return out
        rows = []
        for t1 in terms(self.kernel1.components()):
            for t2 in terms(self.kernel2.components()):
                term = list(t1) + list(t2)
                term[len(t1) - 1] = term[len(t1) - 1][:6] + (1.0,) + term[len(t1) - 1][7:]      # mul_next; slot 7 (CARMA2's sign) kept
                dim = 1
                for r in term:
                    dim *= STATE_DIM[int(r[0])]
                if dim > 6 or len(term) > 3:
                    raise NotImplementedError("a Product with more than 3 factors or a state dimension above 6 is "
                                              "unsupported by the B200 quasiseparable solver backend")
                rows += term
        return rows

    def tau_program(self, dist):
        return self.kernel1.tau_program(dist) + self.kernel2.tau_program(dist) + [(OP_MUL, 0, 0.0, 0.0)]

    # the product's state-space model in the reference's (interleaved Kronecker) state order, quasisep.py:304-331
    def coord_to_sortable(self, X):
        return self.kernel1.coord_to_sortable(X)

    def design_matrix(self):
        from tinygp_b200.solvers.quasisep.block import ensure_dense
        F1, F2 = ensure_dense(self.kernel1.design_matrix()), ensure_dense(self.kernel2.design_matrix())
        return _prod(F1, np.eye(F2.shape[0])) + _prod(np.eye(F1.shape[0]), F2)

    def stationary_covariance(self):
        from tinygp_b200.solvers.quasisep.block import ensure_dense
        return _prod(ensure_dense(self.kernel1.stationary_covariance()), ensure_dense(self.kernel2.stationary_covariance()))

    def observation_model(self, X):
        return _prod(self.kernel1.observation_model(X), self.kernel2.observation_model(X))

    def transition_matrix(self, X1, X2):
        from tinygp_b200.solvers.quasisep.block import ensure_dense
        return _prod(ensure_dense(self.kernel1.transition_matrix(X1, X2)), ensure_dense(self.kernel2.transition_matrix(X1, X2)))


class Celerite(Quasisep):
    """exp(-c tau) [a cos(d tau) + b sin(d tau)], quasisep.py:343-401"""

    def __init__(self, a, b, c, d):
        self.a, self.b, self.c, self.d = a, b, c, d

    def components(self):
        return [(QS_CELERITE, 1.0, float(self.a), float(self.b), float(self.c), float(self.d))]

    def tau_program(self, dist):
        a, b, c, d = float(self.a), float(self.b), float(self.c), float(self.d)
        return [(OP_EXPCOS, dist, c, d), (OP_CONST, 0, a, 0.0), (OP_MUL, 0, 0.0, 0.0),
                (OP_EXPSIN, dist, c, d), (OP_CONST, 0, b, 0.0), (OP_MUL, 0, 0.0, 0.0), (OP_ADD, 0, 0.0, 0.0)]


class SHO(Quasisep):
    """quasisep.py:404-488"""

    def __init__(self, omega, quality, sigma=1.0):
        self.omega, self.quality, self.sigma = omega, quality, sigma

    def components(self):
        return [(QS_SHO, 1.0, float(self.omega), float(self.quality), float(self.sigma), 0.0)]

    def tau_program(self, dist):
        """sigma^2 exp(-w tau / 2Q) [cos + sin / g] (Q > 1/2), [cosh + sinh / f] (Q < 1/2, written as two
        exponentials), (1 + w tau) exp(-w tau) (Q = 1/2): the regimes of quasisep.py:449-488."""
        w, q, s2 = float(self.omega), float(self.quality), float(self.sigma) ** 2
        beta = 0.5 * w / q
        scale = [(OP_CONST, 0, s2, 0.0), (OP_MUL, 0, 0.0, 0.0)]
        if np.allclose(q, 0.5):
            return [(OP_MATERN32, dist, np.sqrt(3.0) / w, 0.0)] + scale
        if q > 0.5:
            g = np.sqrt(4 * q * q - 1)
            return [(OP_EXPCOS, dist, beta, g * beta), (OP_EXPSIN, dist, beta, g * beta),
                    (OP_CONST, 0, 1.0 / g, 0.0), (OP_MUL, 0, 0.0, 0.0), (OP_ADD, 0, 0.0, 0.0)] + scale
        f = np.sqrt(1 - 4 * q * q)
        return [(OP_EXPCOS, dist, beta * (1 - f), 0.0), (OP_CONST, 0, 0.5 * (1 + 1 / f), 0.0), (OP_MUL, 0, 0.0, 0.0),
                (OP_EXPCOS, dist, beta * (1 + f), 0.0), (OP_CONST, 0, 0.5 * (1 - 1 / f), 0.0), (OP_MUL, 0, 0.0, 0.0),
                (OP_ADD, 0, 0.0, 0.0)] + scale


class Exp(Quasisep):
    """sigma^2 exp(-tau / scale), quasisep.py:491-525"""

    def __init__(self, scale, sigma=1.0):
        self.scale, self.sigma = scale, sigma

    def components(self):
        return [(QS_EXP, 1.0, float(self.scale), float(self.sigma), 0.0, 0.0)]

    def tau_program(self, dist):
        return [(OP_EXP, dist, float(self.scale), 0.0), (OP_CONST, 0, float(self.sigma) ** 2, 0.0), (OP_MUL, 0, 0.0, 0.0)]


class Matern32(Quasisep):
    """quasisep.py:528-569"""

    def __init__(self, scale, sigma=1.0):
        self.scale, self.sigma = scale, sigma

    def components(self):
        return [(QS_MATERN32, 1.0, float(self.scale), float(self.sigma), 0.0, 0.0)]

    def tau_program(self, dist):
        return [(OP_MATERN32, dist, float(self.scale), 0.0), (OP_CONST, 0, float(self.sigma) ** 2, 0.0),
                (OP_MUL, 0, 0.0, 0.0)]


class Matern52(Quasisep):
    """quasisep.py:572-633"""

    def __init__(self, scale, sigma=1.0):
        self.scale, self.sigma = scale, sigma

    def components(self):
        return [(QS_MATERN52, 1.0, float(self.scale), float(self.sigma), 0.0, 0.0)]

    def tau_program(self, dist):
        return [(OP_MATERN52, dist, float(self.scale), 0.0), (OP_CONST, 0, float(self.sigma) ** 2, 0.0),
                (OP_MUL, 0, 0.0, 0.0)]


class Cosine(Quasisep):
    """quasisep.py:636-673"""

    def __init__(self, scale, sigma=1.0):
        self.scale, self.sigma = scale, sigma

    def components(self):
        return [(QS_COSINE, 1.0, float(self.scale), float(self.sigma), 0.0, 0.0)]

    def tau_program(self, dist):
        return [(OP_COSINE, dist, float(self.scale), 0.0), (OP_CONST, 0, float(self.sigma) ** 2, 0.0),
                (OP_MUL, 0, 0.0, 0.0)]


class CARMA(Quasisep):
    """CARMA(p, q) process (quasisep.py:690-900).  The host finds the roots of the autoregressive polynomial and the
    autocovariance coefficients once (O(p^2), `carma_roots` :903-906, `carma_acvf` :989-1029) and lowers the process to
    the device's block-diagonal state-space model: one ``Exp`` component per real root (stationary variance +-1 = the
    sign of Re(acf), :870) and one ``B200GP_QS_CARMA2`` component per complex pair (Celerite-type transition :886-900,
    Pinf of :866-882, observation model of :770-792 incl. its ``ravel(om_complex)[::2]`` selection).  Per-point work
    then runs in the model-specialised CUDA scans like every other quasiseparable kernel."""

    def __init__(self, alpha, beta):
        sigma = 1.0
        alpha = np.atleast_1d(np.asarray(alpha, dtype=np.float64))
        beta = np.atleast_1d(np.asarray(beta, dtype=np.float64))
        assert alpha.ndim == 1
        assert beta.ndim == 1
        p = alpha.shape[0]
        assert beta.shape[0] <= p
        arroots = carma_roots(np.append(alpha, 1.0))
        acf = carma_acvf(arroots, alpha, beta * sigma)
        real_mask = np.abs(arroots.imag) < 10 * np.finfo(np.float64).eps        # :758
        complex_mask = ~real_mask
        complex_idx = np.cumsum(complex_mask) * complex_mask
        complex_select = complex_mask * complex_idx % 2
        with np.errstate(all="ignore"):
            om_real = np.sqrt(np.abs(acf.real))
            a, b, c, d = 2 * acf.real, 2 * acf.imag, -arroots.real, -arroots.imag
            c2, d2 = np.square(c), np.square(d)
            s2 = c2 + d2
            denom = np.where(real_mask, 1.0, 2 * c * s2)
            h2_2 = d2 * (a * c - b * d) / denom
            h2 = np.sqrt(h2_2)
            denom = np.where(real_mask, 1.0, d)
            h1 = (c * h2 - np.sqrt(a * d2 - s2 * h2_2)) / denom
        om_complex = np.array([h1, h2])
        self.obsmodel = np.where(real_mask, om_real, np.ravel(om_complex)[::2])  # :792
        self.alpha, self.beta, self.sigma, self.arroots, self.acf = alpha, beta, sigma, arroots, acf
        self._real_mask, self._complex_mask, self._complex_select = real_mask, complex_mask, complex_select

    @classmethod
    def init(cls, alpha, beta):
        return cls(alpha, beta)

    @classmethod
    def from_quads(cls, alpha_quads, beta_quads, beta_mult):
        """quasisep.py:806-838"""
        alpha_quads, beta_quads, beta_mult = (np.atleast_1d(np.asarray(v, dtype=np.float64))
                                              for v in (alpha_quads, beta_quads, beta_mult))
        alpha = carma_quads2poly(np.append(alpha_quads, 1.0))[:-1]
        beta = carma_quads2poly(np.append(beta_quads, beta_mult))
        return cls(alpha, beta)

    def _blocks(self):
        """[(kind, params...)] in state order: real roots and complex pairs as the roots are sorted (:903-906)"""
        out, i, p = [], 0, self.arroots.shape[0]
        sgn = np.where(self.acf.real > 0, 1.0, -1.0)
        while i < p:
            if self._real_mask[i]:
                out.append(("real", -self.arroots.real[i], self.obsmodel[i], sgn[i]))
                i += 1
                continue
            if not (self._complex_select[i] == 1 and i + 1 < p and self._complex_mask[i + 1] and self._complex_select[i + 1] == 0):
                raise ValueError("CARMA: the complex roots do not come in adjacent conjugate pairs")
            if sgn[i + 1] != sgn[i]:
                raise ValueError("CARMA: Re(acf) changes sign inside a conjugate pair")
            out.append(("pair", -self.arroots.real[i], -self.arroots.imag[i], self.obsmodel[i], self.obsmodel[i + 1], sgn[i]))
            i += 2
        return out

    def components(self):
        comps = []
        for blk in self._blocks():
            if blk[0] == "real":
                _, c, h, s = blk
                comps.append((QS_EXP, float(s), float(1.0 / c), float(h), float(c), 0.0, 0.0, 0.0))   # p2 = the rate itself
            else:
                _, c, d, h1, h2, s = blk
                comps.append((QS_CARMA2, 1.0, float(c), float(d), float(h1), float(h2), 0.0, float(s)))
        return comps

    def tau_program(self, dist):
        """k(tau) = h^T T(tau)^T Pinf h block by block (quasisep.py:201-210): s h^2 exp(-c tau) per real root and
        exp(-c tau) [A cos(d tau) + B sin(d tau)], A = h^T P h, B = h^T J P h with J = [[0, -1], [1, 0]], per pair."""
        prog, first = [], True
        for blk in self._blocks():
            if blk[0] == "real":
                _, c, h, s = blk
                term = [(OP_EXP, dist, float(1.0 / c), 0.0), (OP_CONST, 0, float(s * h * h), 0.0), (OP_MUL, 0, 0.0, 0.0)]
            else:
                _, c, d, h1, h2, s = blk
                P = np.array([[s, -c / d], [-c / d, s + 2.0 * (c / d) ** 2]])
                h = np.array([h1, h2])
                A = float(h @ P @ h)
                B = float(h @ np.array([[0.0, -1.0], [1.0, 0.0]]) @ P @ h)
                term = [(OP_EXPCOS, dist, float(c), float(d)), (OP_CONST, 0, A, 0.0), (OP_MUL, 0, 0.0, 0.0),
                        (OP_EXPSIN, dist, float(c), float(d)), (OP_CONST, 0, B, 0.0), (OP_MUL, 0, 0.0, 0.0), (OP_ADD, 0, 0.0, 0.0)]
            prog += term
            if not first:
                prog.append((OP_ADD, 0, 0.0, 0.0))
            first = False
        return prog


def carma_roots(poly_coeffs):
    """quasisep.py:903-906"""
    roots = np.roots(np.asarray(poly_coeffs, dtype=np.float64)[::-1]).astype(np.complex128)
    return roots[np.argsort(roots.real, kind="stable")]


def carma_quads2poly(quads_coeffs):
    """quasisep.py:909-947"""
    quads_coeffs = np.asarray(quads_coeffs, dtype=np.float64)
    size = quads_coeffs.shape[0] - 1
    remain, n_pair = size % 2, size // 2
    mult_f = quads_coeffs[-1:]
    poly = np.array([1.0, quads_coeffs[-2]]) if remain == 1 else np.array([0.0, 1.0])
    poly = poly[-remain + 1:]
    for p in range(n_pair):
        poly = np.convolve(poly, np.append(np.array([quads_coeffs[p * 2], quads_coeffs[p * 2 + 1]]), np.ones(1))[::-1])
    return poly[::-1] * mult_f


def carma_poly2quads(poly_coeffs):
    """quasisep.py:950-986"""
    poly_coeffs = np.asarray(poly_coeffs, dtype=np.float64)
    quads = np.empty(0)
    mult_f = poly_coeffs[-1]
    roots = carma_roots(poly_coeffs / mult_f)
    odd = bool(len(roots) & 0x1)
    roots_comp = roots[roots.imag != 0]
    roots_real = roots[roots.imag == 0]
    for i in range(len(roots_comp) // 2):
        r1, r2 = roots_comp[i], roots_comp[i + 1]
        quads = np.append(quads, [(r1 * r2).real, -(r1.real + r2.real)])
    for i in range(len(roots_real) // 2):
        r1, r2 = roots_real[i], roots_real[i + 1]
        quads = np.append(quads, [(r1 * r2).real, -(r1.real + r2.real)])
    if odd:
        quads = np.append(quads, -roots_real[-1].real)
    return np.append(quads, mult_f)


def carma_acvf(arroots, arparam, maparam):
    """quasisep.py:989-1029: the autocovariance coefficient of every root (Kelly et al. 2014, eq. 4)"""
    arparam, maparam = np.atleast_1d(arparam), np.atleast_1d(maparam)
    p, q = arparam.shape[0], maparam.shape[0] - 1
    sigma = maparam[0]
    maparam = maparam / sigma
    num_left = np.zeros(p, dtype=np.complex128)
    num_right = np.zeros(p, dtype=np.complex128)
    denom = -2 * arroots.real + np.zeros_like(arroots) * 1j
    for k in range(q + 1):
        num_left = num_left + maparam[k] * np.power(arroots, k)
        num_right = num_right + maparam[k] * np.power(np.negative(arroots), k)
    root_idx = np.arange(p)
    for j in range(1, p):
        root_k = arroots[np.roll(root_idx, j)]
        denom = denom * ((root_k - arroots) * (np.conj(root_k) + arroots))
    return sigma**2 * num_left * num_right / denom


# coord_to_sortable implementations that only forward to a wrapped kernel (Quasisep._sortable sees through them)
_FORWARDED_COORDS = (Sum.coord_to_sortable, Product.coord_to_sortable, Wrapper.coord_to_sortable)
