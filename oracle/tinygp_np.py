"""NumPy/SciPy restatement of the tinygp hot path (TEST ORACLE -- see oracle/__init__.py).

Every function cites the reference file:line (relative to /root/reference/src/tinygp)
it follows.  Operation order follows the reference so that differences to XLA are
limited to libm/LAPACK rounding.  Pinned against reference-generated goldens
(tests/golden/reference_vectors.json, <= 1e-12): see the package docstring.
"""

from __future__ import annotations

import numpy as np
import scipy.linalg as sla

__all__ = [
    "L1Distance", "L2Distance",
    "Constant", "Sum", "Product",
    "Exp", "ExpSquared", "Matern32", "Matern52", "Cosine", "ExpSineSquared",
    "RationalQuadratic",
    "Transform", "Linear", "Cholesky", "Subspace",
    "Diagonal", "DirectSolver", "GaussianProcess",
    "qs", "QuasisepSolver", "KalmanLogp",
]


# ----------------------------------------------------------------------------
# kernels/distance.py
# ----------------------------------------------------------------------------
def _diff(X1, X2):
    """Pairwise explicit differences, (N1, N2, D). kernels/base.py:94-96 (double vmap)."""
    X1 = np.asarray(X1, dtype=np.float64)
    X2 = np.asarray(X2, dtype=np.float64)
    if X1.ndim == 1:
        X1 = X1[:, None]
    if X2.ndim == 1:
        X2 = X2[:, None]
    return X1[:, None, :] - X2[None, :, :]


class L1Distance:
    """kernels/distance.py:41-45"""

    def distance(self, d):
        return np.sum(np.abs(d), axis=-1)

    def squared_distance(self, d):
        # Distance.squared_distance default: square of distance (distance.py:30-38)
        return np.square(self.distance(d))


class L2Distance:
    """kernels/distance.py:48-59"""

    def distance(self, d):
        r1 = np.sum(np.abs(d), axis=-1)
        r2 = self.squared_distance(d)
        zeros = r2 == 0
        r2 = np.where(zeros, 1.0, r2)
        return np.where(zeros, r1, np.sqrt(r2))

    def squared_distance(self, d):
        return np.sum(np.square(d), axis=-1)


# ----------------------------------------------------------------------------
# kernels/base.py
# ----------------------------------------------------------------------------
class Kernel:
    def evaluate_diff(self, d):  # d: (..., D) explicit differences
        raise NotImplementedError

    def evaluate_pts(self, A, B):
        """kernel between broadcastable point arrays (..., D); stationary kernels only need A - B, the
        transforms of transforms.py map the points first."""
        return self.evaluate_diff(A - B)

    def __call__(self, X1, X2=None, chunk=2048):
        """kernels/base.py:84-103"""
        if X2 is None:
            # evaluate_diag(X) = evaluate(X, X)  (base.py:59-66,85-86)
            X1 = np.asarray(X1, dtype=np.float64).reshape(np.shape(X1)[0], -1)
            return self.evaluate_pts(X1, X1)
        X1 = np.asarray(X1, dtype=np.float64)
        X2 = np.asarray(X2, dtype=np.float64)
        if X1.ndim == 1:
            X1 = X1[:, None]
        if X2.ndim == 1:
            X2 = X2[:, None]
        out = np.empty((X1.shape[0], X2.shape[0]))
        for s in range(0, X1.shape[0], chunk):
            out[s:s + chunk] = self.evaluate_pts(X1[s:s + chunk, None, :], X2[None, :, :])
        return out

    def matmul(self, X1, X2=None, y=None):
        """kernels/base.py:68-82"""
        if y is None:
            y = X2
            X2 = None
        if X2 is None:
            X2 = X1
        return self(X1, X2) @ y

    def __add__(self, other):
        return Sum(self, other if isinstance(other, Kernel) else Constant(other))

    def __radd__(self, other):
        if not isinstance(other, Kernel) and other == 0:
            return self
        return Sum(other if isinstance(other, Kernel) else Constant(other), self)

    def __mul__(self, other):
        return Product(self, other if isinstance(other, Kernel) else Constant(other))

    def __rmul__(self, other):
        return Product(other if isinstance(other, Kernel) else Constant(other), self)


class Sum(Kernel):
    """kernels/base.py:170-177"""

    def __init__(self, kernel1, kernel2):
        self.kernel1, self.kernel2 = kernel1, kernel2

    def evaluate_diff(self, d):
        return self.kernel1.evaluate_diff(d) + self.kernel2.evaluate_diff(d)

    def evaluate_pts(self, A, B):
        return self.kernel1.evaluate_pts(A, B) + self.kernel2.evaluate_pts(A, B)


class Product(Kernel):
    """kernels/base.py:180-187"""

    def __init__(self, kernel1, kernel2):
        self.kernel1, self.kernel2 = kernel1, kernel2

    def evaluate_diff(self, d):
        return self.kernel1.evaluate_diff(d) * self.kernel2.evaluate_diff(d)

    def evaluate_pts(self, A, B):
        return self.kernel1.evaluate_pts(A, B) * self.kernel2.evaluate_pts(A, B)


class Constant(Kernel):
    """kernels/base.py:190-209"""

    def __init__(self, value):
        if np.ndim(value) != 0:
            raise ValueError("The value of a constant kernel must be a scalar")
        self.value = float(value)

    def evaluate_diff(self, d):
        return np.full(d.shape[:-1], self.value)


# ----------------------------------------------------------------------------
# kernels/stationary.py
# ----------------------------------------------------------------------------
class Stationary(Kernel):
    default_distance = L1Distance

    def __init__(self, scale=1.0, distance=None):
        if np.ndim(scale):
            raise ValueError("Only scalar scales are permitted for stationary kernels")
        self.scale = float(scale)
        self.distance = self.default_distance() if distance is None else distance


class Exp(Stationary):
    """kernels/stationary.py:76-82"""

    def evaluate_diff(self, d):
        return np.exp(-self.distance.distance(d) / self.scale)


class ExpSquared(Stationary):
    """kernels/stationary.py:104-106 (default L2Distance, :102)"""
    default_distance = L2Distance

    def evaluate_diff(self, d):
        r2 = self.distance.squared_distance(d) / np.square(self.scale)
        return np.exp(-0.5 * r2)


class Matern32(Stationary):
    """kernels/stationary.py:126-129"""

    def evaluate_diff(self, d):
        r = self.distance.distance(d) / self.scale
        arg = np.sqrt(3) * r
        return (1 + arg) * np.exp(-arg)


class Matern52(Stationary):
    """kernels/stationary.py:150-153"""

    def evaluate_diff(self, d):
        r = self.distance.distance(d) / self.scale
        arg = np.sqrt(5) * r
        return (1 + arg + np.square(arg) / 3) * np.exp(-arg)


class Cosine(Stationary):
    """kernels/stationary.py:173-175"""

    def evaluate_diff(self, d):
        r = self.distance.distance(d) / self.scale
        return np.cos(2 * np.pi * r)


class ExpSineSquared(Stationary):
    """kernels/stationary.py:202-205"""

    def __init__(self, scale=1.0, distance=None, *, gamma=None):
        super().__init__(scale, distance)
        if gamma is None:
            raise ValueError("Missing required argument 'gamma'")
        self.gamma = float(gamma)

    def evaluate_diff(self, d):
        r = self.distance.distance(d) / self.scale
        return np.exp(-self.gamma * np.square(np.sin(np.pi * r)))


class RationalQuadratic(Stationary):
    """kernels/stationary.py:232-235 -- NOTE default distance is L1 (stationary.py:56),
    so r2 = (sum|d|)^2 / scale^2 unless distance=L2Distance() is passed."""

    def __init__(self, scale=1.0, distance=None, *, alpha=None):
        super().__init__(scale, distance)
        if alpha is None:
            raise ValueError("Missing required argument 'alpha'")
        self.alpha = float(alpha)

    def evaluate_diff(self, d):
        r2 = self.distance.squared_distance(d) / np.square(self.scale)
        return (1.0 + 0.5 * r2 / self.alpha) ** -self.alpha


# ----------------------------------------------------------------------------
# transforms.py
# ----------------------------------------------------------------------------
class _PointTransform(Kernel):
    """kernel.evaluate(t(X1), t(X2)) with t applied to every point (last axis)."""

    def _map(self, A):
        raise NotImplementedError

    def evaluate_pts(self, A, B):
        return self.kernel.evaluate_pts(self._map(np.asarray(A, dtype=np.float64)),
                                        self._map(np.asarray(B, dtype=np.float64)))

    def evaluate_diff(self, d):
        raise TypeError("a transformed kernel is not a function of the raw coordinate difference")


class Transform(_PointTransform):
    """transforms.py:23-37"""

    def __init__(self, transform, kernel):
        self.transform, self.kernel = transform, kernel

    def _map(self, A):
        flat = A.reshape(-1, A.shape[-1])
        out = np.asarray([np.atleast_1d(self.transform(p[0] if p.size == 1 else p)) for p in flat], dtype=np.float64)
        return out.reshape(A.shape[:-1] + (-1,))


class Linear(_PointTransform):
    """transforms.py:40-74"""

    def __init__(self, scale, kernel):
        self.scale, self.kernel = np.asarray(scale, dtype=np.float64), kernel

    def _map(self, A):
        if self.scale.ndim < 2:
            return self.scale * A                       # jnp.multiply(scale, X)
        if self.scale.ndim == 2:
            return A @ self.scale.T                     # jnp.dot(scale, X) for every point X
        raise ValueError("'scale' must be 0-, 1-, or 2-dimensional")


class Cholesky(_PointTransform):
    """transforms.py:77-136"""

    def __init__(self, factor, kernel):
        self.factor, self.kernel = np.asarray(factor, dtype=np.float64), kernel

    def _map(self, A):
        if self.factor.ndim < 2:
            return (1.0 / self.factor) * A
        if self.factor.ndim == 2:
            flat = A.reshape(-1, A.shape[-1])
            out = sla.solve_triangular(self.factor, flat.T, lower=True).T
            return out.reshape(A.shape)
        raise ValueError("'scale' must be 0-, 1-, or 2-dimensional")

    @classmethod
    def from_parameters(cls, diagonal, off_diagonal, kernel):
        diagonal, off_diagonal = np.asarray(diagonal, float), np.asarray(off_diagonal, float)
        ndim = diagonal.size
        if off_diagonal.size != ((ndim - 1) * ndim) // 2:
            raise ValueError("Dimension mismatch")
        factor = np.zeros((ndim, ndim))
        factor[np.diag_indices(ndim)] += diagonal
        factor[np.tril_indices(ndim, -1)] += off_diagonal
        return cls(factor, kernel)


class Subspace(_PointTransform):
    """transforms.py:139-161"""

    def __init__(self, axis, kernel):
        self.axis, self.kernel = axis, kernel

    def _map(self, A):
        out = A[..., self.axis] if np.ndim(self.axis) == 0 else A[..., list(self.axis)]
        return out[..., None] if np.ndim(self.axis) == 0 else out


# ----------------------------------------------------------------------------
# noise.py
# ----------------------------------------------------------------------------
class Diagonal:
    """noise.py:55-95"""

    def __init__(self, diag):
        diag = np.asarray(diag, dtype=np.float64)
        if diag.ndim != 1:
            raise ValueError("The diagonal for the noise model be the same shape as the data")
        self.diag = diag

    def diagonal(self):
        return self.diag

    def add_to(self, other):  # _add, noise.py:77-78
        out = np.array(other, dtype=np.float64, copy=True)
        idx = np.arange(out.shape[0])
        out[idx, idx] += self.diag
        return out

    def matmul(self, other):  # noise.py:86-90
        other = np.asarray(other)
        return self.diag * other if other.ndim == 1 else self.diag[:, None] * other


class Dense:
    """noise.py:98-123: a full-rank N x N observation model"""

    def __init__(self, value):
        self.value = np.asarray(value, dtype=np.float64)

    def diagonal(self):
        return np.diag(self.value)

    def add_to(self, other):  # noise.py:109-113
        return self.value + np.asarray(other, dtype=np.float64)

    def matmul(self, other):  # noise.py:115-116
        return self.value @ np.asarray(other, dtype=np.float64)


class Banded:
    """noise.py:126-240: diag (N,) and J symmetric off-diagonals, off_diags[n, j] = N[n, n + j + 1]"""

    def __init__(self, diag, off_diags):
        self.diag = np.asarray(diag, dtype=np.float64)
        self.off_diags = np.asarray(off_diags, dtype=np.float64)

    def diagonal(self):
        return self.diag

    def add_to(self, other):  # noise.py:196-215
        out = np.array(other, dtype=np.float64, copy=True)
        N, J = self.off_diags.shape
        for n in range(N):
            out[n, n] += self.diag[n]
            for j in range(J):
                if n + j + 1 < N:
                    out[n, n + j + 1] += self.off_diags[n, j]
                    out[n + j + 1, n] += self.off_diags[n, j]
        return out

    def matmul(self, other):  # noise.py:223-224 (to_qsm() @ other: the same banded matrix)
        N = self.diag.shape[0]
        return self.add_to(np.zeros((N, N))) @ np.asarray(other, dtype=np.float64)

    def to_qsm_arrays(self):  # noise.py:226-240: (d, p, q, a) with p = e_1, q = off_diags, a = the upper shift
        N, J = self.off_diags.shape
        p = np.zeros((N, J))
        p[:, 0] = 1.0
        a = np.zeros((N, J, J))
        for j in range(J - 1):
            a[:, j, j + 1] = 1.0
        return self.diag, p, self.off_diags, a


# ----------------------------------------------------------------------------
# solvers/direct.py
# ----------------------------------------------------------------------------
class DirectSolver:
    """solvers/direct.py:17-95.  linalg.cholesky -> LAPACK dpotrf (same routine XLA:CPU calls)."""

    def __init__(self, kernel, X, noise, *, covariance=None):
        self.X = np.asarray(X, dtype=np.float64)
        self.variance_value = kernel(self.X) + noise.diagonal()          # direct.py:49
        if covariance is None:
            covariance = noise.add_to(kernel(self.X, self.X))            # direct.py:51
        self.covariance_value = covariance
        # direct.py:53 ; JAX returns NaNs (no exception) for non-PD input
        try:
            self.scale_tril = sla.cholesky(covariance, lower=True, check_finite=False)
        except sla.LinAlgError:
            self.scale_tril = np.full_like(covariance, np.nan)

    def variance(self):
        return self.variance_value

    def covariance(self):
        return self.covariance_value

    def normalization(self):  # direct.py:61-64
        return np.sum(np.log(np.diag(self.scale_tril))) + 0.5 * self.scale_tril.shape[0] * np.log(2 * np.pi)

    def solve_triangular(self, y, *, transpose=False):  # direct.py:66-70
        return sla.solve_triangular(self.scale_tril, y, lower=True, trans=1 if transpose else 0,
                                    check_finite=False)

    def dot_triangular(self, y):  # direct.py:72-73
        return np.einsum("ij,j...->i...", self.scale_tril, y)

    def condition(self, kernel, X_test, noise):  # direct.py:75-95
        if X_test is None:
            Ks = kernel(self.X, self.X)
            Kss = noise.add_to(Ks)
        else:
            Ks = kernel(self.X, X_test)
            Kss = noise.add_to(kernel(X_test, X_test))
        A = self.solve_triangular(Ks)
        return Kss - A.T @ A


# ----------------------------------------------------------------------------
# solvers/quasisep: generators + recursions, dense `a` (J x J) throughout
# ----------------------------------------------------------------------------
class qs:
    """Namespace for the quasiseparable kernels (kernels/quasisep.py)."""

    class Quasisep:
        def design_matrix(self):
            raise NotImplementedError

        def stationary_covariance(self):
            raise NotImplementedError

        def observation_model(self, X):
            raise NotImplementedError

        def transition_matrix(self, X1, X2):
            raise NotImplementedError

        def __add__(self, other):
            return qs.Sum(self, other)

        def __radd__(self, other):
            if not isinstance(other, qs.Quasisep) and other == 0:
                return self
            return qs.Sum(other, self)

        def __mul__(self, other):
            if isinstance(other, qs.Quasisep):
                return qs.Product(self, other)             # kernels/quasisep.py:181-190
            return qs.Scale(self, other)

        def __rmul__(self, other):
            if isinstance(other, qs.Quasisep):
                return qs.Product(other, self)
            return qs.Scale(self, other)

        def to_symm_qsm(self, X):
            """kernels/quasisep.py:102-116 -> (d, p, q, a)"""
            X = np.asarray(X, dtype=np.float64)
            Pinf = self.stationary_covariance()
            Xs = np.append(X[0], X[:-1])                                   # :105-107  (a_0 = T(t0,t0))
            a = np.stack([self.transition_matrix(x1, x2) for x1, x2 in zip(Xs, X)])
            h = np.stack([self.observation_model(x) for x in X])
            hP = h @ Pinf
            d = np.sum(hP * h, axis=1)
            a = np.swapaxes(a, -1, -2)
            p = np.einsum("nj,njk->nk", h, a)
            q = hP
            return d, p, q, a

        def evaluate(self, X1, X2):
            """kernels/quasisep.py:201-210 (scalar)"""
            Pinf = self.stationary_covariance()
            h1 = self.observation_model(X1)
            h2 = self.observation_model(X2)
            if X1 < X2:
                return h2 @ self.transition_matrix(X1, X2).T @ Pinf @ h1
            return h1 @ self.transition_matrix(X2, X1).T @ Pinf @ h2

        def matmul(self, X1, X2=None, y=None):
            """quasisep.py:147-163 (to_general_qsm @ y); evaluated densely here -- same values."""
            if y is None:
                y, X2 = X2, None
            if X2 is None:
                X2 = X1
            return self(X1, X2) @ np.asarray(y, dtype=np.float64)

        def __call__(self, X1, X2=None):
            X1 = np.asarray(X1, dtype=np.float64)
            if X2 is None:
                Pinf = self.stationary_covariance()
                return np.array([self.observation_model(x) @ Pinf @ self.observation_model(x) for x in X1])
            X2 = np.asarray(X2, dtype=np.float64)
            return np.array([[self.evaluate(x1, x2) for x2 in X2] for x1 in X1])

    class Sum(Quasisep):
        """kernels/quasisep.py:241-295 (Block == block_diag when densified)"""

        def __init__(self, kernel1, kernel2):
            self.kernel1, self.kernel2 = kernel1, kernel2

        def design_matrix(self):
            return sla.block_diag(self.kernel1.design_matrix(), self.kernel2.design_matrix())

        def stationary_covariance(self):
            return sla.block_diag(self.kernel1.stationary_covariance(), self.kernel2.stationary_covariance())

        def observation_model(self, X):
            return np.concatenate((self.kernel1.observation_model(X), self.kernel2.observation_model(X)))

        def transition_matrix(self, X1, X2):
            return sla.block_diag(self.kernel1.transition_matrix(X1, X2), self.kernel2.transition_matrix(X1, X2))

    class Product(Quasisep):
        """kernels/quasisep.py:298-331 with `_prod_helper` (:676-687): Kronecker-structured state, the FIRST kernel's state
        index running fastest (np.meshgrid's default "xy" indexing)."""

        def __init__(self, kernel1, kernel2):
            self.kernel1, self.kernel2 = kernel1, kernel2

        @staticmethod
        def _prod_helper(a1, a2):
            a1, a2 = np.asarray(a1, dtype=np.float64), np.asarray(a2, dtype=np.float64)
            i, j = np.meshgrid(np.arange(a1.shape[0]), np.arange(a2.shape[0]))
            i, j = i.flatten(), j.flatten()
            if a1.ndim == 1:
                return a1[i] * a2[j]
            return a1[i[:, None], i[None, :]] * a2[j[:, None], j[None, :]]

        def design_matrix(self):
            F1, F2 = self.kernel1.design_matrix(), self.kernel2.design_matrix()
            return self._prod_helper(F1, np.eye(F2.shape[0])) + self._prod_helper(np.eye(F1.shape[0]), F2)

        def stationary_covariance(self):
            return self._prod_helper(self.kernel1.stationary_covariance(), self.kernel2.stationary_covariance())

        def observation_model(self, X):
            return self._prod_helper(self.kernel1.observation_model(X), self.kernel2.observation_model(X))

        def transition_matrix(self, X1, X2):
            return self._prod_helper(self.kernel1.transition_matrix(X1, X2), self.kernel2.transition_matrix(X1, X2))

    class Scale(Quasisep):
        """kernels/quasisep.py:334-340"""

        def __init__(self, kernel, scale):
            self.kernel, self.scale = kernel, float(scale)

        def design_matrix(self):
            return self.kernel.design_matrix()

        def stationary_covariance(self):
            return self.scale * self.kernel.stationary_covariance()

        def observation_model(self, X):
            return self.kernel.observation_model(X)

        def transition_matrix(self, X1, X2):
            return self.kernel.transition_matrix(X1, X2)

    class Celerite(Quasisep):
        """kernels/quasisep.py:343-401"""

        def __init__(self, a, b, c, d):
            self.a, self.b, self.c, self.d = map(float, (a, b, c, d))

        def design_matrix(self):
            return np.array([[-self.c, -self.d], [self.d, -self.c]])

        def stationary_covariance(self):
            c, d = self.c, self.d
            return np.array([[1, -c / d], [-c / d, 1 + 2 * np.square(c) / np.square(d)]])

        def observation_model(self, X):
            a, b, c, d = self.a, self.b, self.c, self.d
            c2, d2 = np.square(c), np.square(d)
            s2 = c2 + d2
            h2_2 = d2 * (a * c - b * d) / (2 * c * s2)
            h2 = np.sqrt(h2_2)
            h1 = (c * h2 - np.sqrt(a * d2 - s2 * h2_2)) / d
            return np.array([h1, h2])

        def transition_matrix(self, X1, X2):
            dt = X2 - X1
            cos, sin = np.cos(self.d * dt), np.sin(self.d * dt)
            return np.exp(-self.c * dt) * np.array([[cos, -sin], [sin, cos]]).T

    class SHO(Quasisep):
        """kernels/quasisep.py:404-488"""

        def __init__(self, omega, quality, sigma=1.0):
            self.omega, self.quality, self.sigma = float(omega), float(quality), float(sigma)

        def design_matrix(self):
            return np.array([[0, 1], [-np.square(self.omega), -self.omega / self.quality]])

        def stationary_covariance(self):
            return np.diag(np.array([1, np.square(self.omega)]))

        def observation_model(self, X):
            return np.array([self.sigma, 0])

        def transition_matrix(self, X1, X2):
            dt = X2 - X1
            w, q = self.omega, self.quality
            if np.allclose(q, 0.5):                                       # :483-488 lax.cond
                return np.exp(-w * dt) * np.array([[1 + w * dt, -np.square(w) * dt], [dt, 1 - w * dt]])
            if q > 0.5:
                f = np.sqrt(np.maximum(4 * np.square(q) - 1, 0))
                arg = 0.5 * f * w * dt / q
                sin, cos = np.sin(arg), np.cos(arg)
                return np.exp(-0.5 * w * dt / q) * np.array(
                    [[cos + sin / f, -2 * q * w * sin / f], [2 * q * sin / (w * f), cos - sin / f]])
            f = np.sqrt(np.maximum(1 - 4 * np.square(q), 0))
            arg = 0.5 * f * w * dt / q
            sinh, cosh = np.sinh(arg), np.cosh(arg)
            return np.exp(-0.5 * w * dt / q) * np.array(
                [[cosh + sinh / f, -2 * q * w * sinh / f], [2 * q * sinh / (w * f), cosh - sinh / f]])

    class Exp(Quasisep):
        """kernels/quasisep.py:491-525"""

        def __init__(self, scale, sigma=1.0):
            self.scale, self.sigma = float(scale), float(sigma)

        def design_matrix(self):
            return np.array([[-1 / self.scale]])

        def stationary_covariance(self):
            return np.ones((1, 1))

        def observation_model(self, X):
            return np.array([self.sigma])

        def transition_matrix(self, X1, X2):
            dt = X2 - X1
            return np.exp(-np.array([[dt]]) / self.scale)

    class Matern32(Quasisep):
        """kernels/quasisep.py:528-569"""

        def __init__(self, scale, sigma=1.0):
            self.scale, self.sigma = float(scale), float(sigma)

        def design_matrix(self):
            f = np.sqrt(3) / self.scale
            return np.array([[0, 1], [-np.square(f), -2 * f]])

        def stationary_covariance(self):
            return np.diag(np.array([1, 3 / np.square(self.scale)]))

        def observation_model(self, X):
            return np.array([self.sigma, 0])

        def transition_matrix(self, X1, X2):
            dt = X2 - X1
            f = np.sqrt(3) / self.scale
            return np.exp(-f * dt) * np.array([[1 + f * dt, -np.square(f) * dt], [dt, 1 - f * dt]])

    class Matern52(Quasisep):
        """kernels/quasisep.py:572-633"""

        def __init__(self, scale, sigma=1.0):
            self.scale, self.sigma = float(scale), float(sigma)

        def design_matrix(self):
            f = np.sqrt(5) / self.scale
            f2 = np.square(f)
            return np.array([[0, 1, 0], [0, 0, 1], [-f2 * f, -3 * f2, -3 * f]])

        def stationary_covariance(self):
            f = np.sqrt(5) / self.scale
            f2 = np.square(f)
            f2o3 = f2 / 3
            return np.array([[1, 0, -f2o3], [0, f2o3, 0], [-f2o3, 0, np.square(f2)]])

        def observation_model(self, X):
            return np.array([self.sigma, 0, 0])

        def transition_matrix(self, X1, X2):
            dt = X2 - X1
            f = np.sqrt(5) / self.scale
            f2 = np.square(f)
            d2 = np.square(dt)
            return np.exp(-f * dt) * np.array([
                [0.5 * f2 * d2 + f * dt + 1, -0.5 * f * f2 * d2, 0.5 * f2 * f * dt * (f * dt - 2)],
                [dt * (f * dt + 1), -f2 * d2 + f * dt + 1, f2 * dt * (f * dt - 3)],
                [0.5 * d2, 0.5 * dt * (2 - f * dt), 0.5 * f2 * d2 - 2 * f * dt + 1],
            ])

    class Cosine(Quasisep):
        """kernels/quasisep.py:636-673"""

        def __init__(self, scale, sigma=1.0):
            self.scale, self.sigma = float(scale), float(sigma)

        def design_matrix(self):
            f = 2 * np.pi / self.scale
            return np.array([[0, -f], [f, 0]])

        def stationary_covariance(self):
            return np.eye(2)

        def observation_model(self, X):
            return np.array([self.sigma, 0])

        def transition_matrix(self, X1, X2):
            dt = X2 - X1
            f = 2 * np.pi / self.scale
            cos, sin = np.cos(f * dt), np.sin(f * dt)
            return np.array([[cos, sin], [-sin, cos]])

    class CARMA(Quasisep):
        """kernels/quasisep.py:690-900 (CARMA(p, q) as a sum of real / complex exponential state blocks) with its helpers
        carma_roots (:903-906), carma_quads2poly (:909-947) and carma_acvf (:989-1029), operation for operation -- incl.
        the `ravel(om_complex)[::2]` selection of the complex observation model (:792)."""

        def __init__(self, alpha, beta):
            sigma = 1.0
            alpha = np.atleast_1d(np.asarray(alpha, dtype=np.float64))
            beta = np.atleast_1d(np.asarray(beta, dtype=np.float64))
            assert alpha.ndim == 1 and beta.ndim == 1 and beta.shape[0] <= alpha.shape[0]
            arroots = qs.carma_roots(np.append(alpha, 1.0))
            acf = qs.carma_acvf(arroots, alpha, beta * sigma)
            real_mask = np.abs(arroots.imag) < 10 * np.finfo(np.float64).eps
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
            self.obsmodel = np.where(real_mask, om_real, np.ravel(om_complex)[::2])
            self.alpha, self.beta, self.sigma, self.arroots, self.acf = alpha, beta, sigma, arroots, acf
            self._real_mask, self._complex_mask, self._complex_select = real_mask, complex_mask, complex_select

        @classmethod
        def init(cls, alpha, beta):
            return cls(alpha, beta)

        @classmethod
        def from_quads(cls, alpha_quads, beta_quads, beta_mult):
            alpha_quads, beta_quads, beta_mult = (np.atleast_1d(np.asarray(v, dtype=np.float64))
                                                  for v in (alpha_quads, beta_quads, beta_mult))
            alpha = qs.carma_quads2poly(np.append(alpha_quads, 1.0))[:-1]
            beta = qs.carma_quads2poly(np.append(beta_quads, beta_mult))
            return cls(alpha, beta)

        def design_matrix(self):
            dm_real = np.diag(self.arroots.real * self._real_mask)
            dm_complex_diag = np.diag(self.arroots.real * self._complex_mask)
            dm_complex_u = np.diag((self.arroots.imag * self._complex_select)[:-1], k=1)
            return dm_real + dm_complex_diag + -dm_complex_u.T + dm_complex_u

        def stationary_covariance(self):
            p = self.acf.shape[0]
            diag = np.diag(np.where(self.acf.real > 0, np.ones(p), -np.ones(p)))
            denom = np.where(self._real_mask, 1.0, self.arroots.imag)
            diag_complex = np.diag(2 * np.square(self.arroots.real / denom * np.roll(self._complex_select, 1)
                                                 * self._complex_mask))
            c_over_d = self.arroots.real / denom
            sc_complex_u = np.diag((-c_over_d * self._complex_select)[:-1], k=1)
            return diag + diag_complex + sc_complex_u + sc_complex_u.T

        def observation_model(self, X):
            return self.obsmodel

        def transition_matrix(self, X1, X2):
            dt = X2 - X1
            c, d = -self.arroots.real, -self.arroots.imag
            decay = np.exp(-c * dt)
            sin = np.sin(d * dt)
            tm_real = np.diag(decay * self._real_mask)
            tm_complex_diag = np.diag(decay * np.cos(d * dt) * self._complex_mask)
            tm_complex_u = np.diag((decay * sin * self._complex_select)[:-1], k=1)
            return tm_real + tm_complex_diag + -tm_complex_u.T + tm_complex_u

    @staticmethod
    def carma_roots(poly_coeffs):                                           # kernels/quasisep.py:903-906
        roots = np.roots(np.asarray(poly_coeffs)[::-1]).astype(np.complex128)
        return roots[np.argsort(roots.real, kind="stable")]

    @staticmethod
    def carma_quads2poly(quads_coeffs):                                     # kernels/quasisep.py:909-947
        quads_coeffs = np.asarray(quads_coeffs, dtype=np.float64)
        size = quads_coeffs.shape[0] - 1
        remain, n_pair = size % 2, size // 2
        mult_f = quads_coeffs[-1:]
        poly = np.array([1.0, quads_coeffs[-2]]) if remain == 1 else np.array([0.0, 1.0])
        poly = poly[-remain + 1:]
        for p in range(n_pair):
            poly = np.convolve(poly, np.append(np.array([quads_coeffs[p * 2], quads_coeffs[p * 2 + 1]]), np.ones(1))[::-1])
        return poly[::-1] * mult_f

    @staticmethod
    def carma_acvf(arroots, arparam, maparam):                              # kernels/quasisep.py:989-1029
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


def qs_generators_fast(kernel, X):
    """Vectorised to_symm_qsm for the bench CPU baseline (same formulas, no Python loop per
    point for the closed-form leaves).  Falls back to the scalar loop for unknown kernels."""
    X = np.asarray(X, dtype=np.float64)
    Xs = np.append(X[0], X[:-1])
    dt = X - Xs

    def leaf(k):
        if isinstance(k, qs.Scale):
            Pinf, h, T = leaf(k.kernel)
            return k.scale * Pinf, h, T
        if isinstance(k, qs.Sum):
            P1, h1, T1 = leaf(k.kernel1)
            P2, h2, T2 = leaf(k.kernel2)
            J1, J2 = P1.shape[0], P2.shape[0]
            T = np.zeros((X.shape[0], J1 + J2, J1 + J2))
            T[:, :J1, :J1] = T1
            T[:, J1:, J1:] = T2
            return sla.block_diag(P1, P2), np.concatenate((h1, h2)), T
        if isinstance(k, qs.Matern32):
            f = np.sqrt(3) / k.scale
            e = np.exp(-f * dt)
            T = e[:, None, None] * np.stack([np.stack([1 + f * dt, -np.square(f) * dt], -1),
                                            np.stack([dt, 1 - f * dt], -1)], -2)
            return k.stationary_covariance(), k.observation_model(0.0), T
        if isinstance(k, qs.SHO) and k.quality > 0.5 and not np.allclose(k.quality, 0.5):
            w, q = k.omega, k.quality
            f = np.sqrt(np.maximum(4 * np.square(q) - 1, 0))
            arg = 0.5 * f * w * dt / q
            sin, cos = np.sin(arg), np.cos(arg)
            e = np.exp(-0.5 * w * dt / q)
            T = e[:, None, None] * np.stack([np.stack([cos + sin / f, -2 * q * w * sin / f], -1),
                                            np.stack([2 * q * sin / (w * f), cos - sin / f], -1)], -2)
            return k.stationary_covariance(), k.observation_model(0.0), T
        T = np.stack([k.transition_matrix(x1, x2) for x1, x2 in zip(Xs, X)])
        return k.stationary_covariance(), k.observation_model(0.0), T

    Pinf, h, T = leaf(kernel)
    hP = h @ Pinf
    a = np.swapaxes(T, -1, -2)
    N = X.shape[0]
    d = np.full(N, np.sum(hP * h))
    p = np.einsum("j,njk->nk", h, a)
    q = np.broadcast_to(hP, (N, hP.shape[0])).copy()
    return d, p, q, a


# solvers/quasisep/ops.py ------------------------------------------------------
def qs_cholesky(d, p, q, a):
    """ops.py:352-365"""
    N, J = p.shape
    f = np.zeros((J, J))
    c = np.empty(N)
    w = np.empty((N, J))
    for k in range(N):
        ck = np.sqrt(d[k] - p[k] @ f @ p[k])
        tmp = f @ a[k].T
        wk = (q[k] - p[k] @ tmp) / ck
        f = a[k] @ tmp + np.outer(wk, wk)
        c[k], w[k] = ck, wk
    return c, w


def qs_lower_solve(d, p, q, a, x):
    """ops.py:463-472 ; x (N, K)"""
    N, J = p.shape
    f = np.zeros((J, x.shape[1]))
    out = np.empty_like(x)
    for k in range(N):
        y = (x[k] - p[k] @ f) / d[k]
        f = a[k] @ f + np.outer(q[k], y)
        out[k] = y
    return out


def qs_upper_solve(d, p, q, a, x):
    """ops.py:489-498 (reverse scan)"""
    N, J = p.shape
    f = np.zeros((J, x.shape[1]))
    out = np.empty_like(x)
    for k in range(N - 1, -1, -1):
        y = (x[k] - q[k] @ f) / d[k]
        f = a[k].T @ f + np.outer(p[k], y)
        out[k] = y
    return out


def qs_lower_matmul(p, q, a, x):
    """ops.py:308-316 -- the scan emits the carry *before* the update (exclusive), so
    out_k = p_k . f_k with f_k = a_{k-1} f_{k-1} + q_{k-1} (x) x_{k-1}, f_0 = 0."""
    N, J = p.shape
    f = np.zeros((J, x.shape[1]))
    out = np.empty_like(x)
    for k in range(N):
        out[k] = p[k] @ f
        f = a[k] @ f + np.outer(q[k], x[k])
    return out


def qs_upper_matmul(p, q, a, x):
    """ops.py:330-338 (reverse scan, exclusive carry)"""
    N, J = p.shape
    f = np.zeros((J, x.shape[1]))
    out = np.empty_like(x)
    for k in range(N - 1, -1, -1):
        out[k] = q[k] @ f
        f = a[k].T @ f + np.outer(p[k], x[k])
    return out


def _as2d(x):
    x = np.asarray(x, dtype=np.float64)
    return x.reshape(x.shape[0], -1), x.shape           # core.py:35-44 handle_matvec_shapes


class QuasisepSolver:
    """solvers/quasisep/solver.py:19-139 (dense-fallback condition only)."""

    def __init__(self, kernel, X, noise, *, covariance=None, assume_sorted=False, parallel=False):
        X = np.asarray(X, dtype=np.float64)
        if not assume_sorted and np.any(np.diff(X) < 0.0):              # solver.py:142-146
            raise ValueError("Input coordinates must be sorted in order to use the QuasisepSolver")
        d, p, q, a = kernel.to_symm_qsm(X)                               # solver.py:73
        if hasattr(noise, "to_qsm_arrays"):
            # noise.Banded: SymmQSM + SymmQSM (core.py:161-163 -> StrictLowerTriQSM.self_add, core.py:208-214: p and q are
            # concatenated, a is block-diagonal)
            dn, pn, qn, an = noise.to_qsm_arrays()
            J1, J2 = p.shape[1], pn.shape[1]
            aa = np.zeros((X.shape[0], J1 + J2, J1 + J2))
            aa[:, :J1, :J1], aa[:, J1:, J1:] = a, an
            d, p, q, a = d + dn, np.concatenate((p, pn), axis=1), np.concatenate((q, qn), axis=1), aa
        else:
            d = d + noise.diagonal()                                     # solver.py:74 ; core.py:161-163
        self.X, self.kernel = X, kernel
        self.d, self.p, self.q, self.a = d, p, q, a
        self.c, self.w = qs_cholesky(d, p, q, a)                         # solver.py:82 ; core.py:524-539

    def variance(self):
        return self.d

    def covariance(self):  # solver.py:87-88 -> to_dense (core.py:84-90)
        N = self.d.shape[0]
        eye = np.eye(N)
        return (self.d[:, None] * eye + qs_lower_matmul(self.p, self.q, self.a, eye)
                + qs_upper_matmul(self.p, self.q, self.a, eye))

    def normalization(self):  # solver.py:90-93
        return np.sum(np.log(self.c)) + 0.5 * self.c.shape[0] * np.log(2 * np.pi)

    def solve_triangular(self, y, *, transpose=False):
        y2, shape = _as2d(y)
        if transpose:                                                    # core.py:366-383
            return qs_upper_solve(self.c, self.p, self.w, self.a, y2).reshape(shape)
        return qs_lower_solve(self.c, self.p, self.w, self.a, y2).reshape(shape)  # core.py:319-336

    def dot_triangular(self, y):  # core.py:303-305
        y2, shape = _as2d(y)
        return (self.c[:, None] * y2 + qs_lower_matmul(self.p, self.w, self.a, y2)).reshape(shape)

    def condition(self, kernel, X_test, noise):
        """solver.py:104-139.  The QSM branch (:124-129, X_test None and a quasiseparable kernel) is
        ``M + noise - (factor.inv() @ M).gram()``; restated densely -- (L^-1 M)^T (L^-1 M) is the same matrix -- so this
        branch returns the same VALUES as the reference's SymmQSM, noise included; the dense branch (:131-139) has no
        noise term.  `last_condition_has_noise` tells GaussianProcess.condition which one ran."""
        self.last_condition_has_noise = False
        if X_test is None and isinstance(kernel, qs.Quasisep):
            M = kernel(self.X, self.X)
            A = self.solve_triangular(M)
            self.last_condition_has_noise = True
            return M + np.diag(noise.diagonal()) - A.T @ A
        if X_test is None:
            Kss = Ks = kernel(self.X, self.X)
        else:
            Kss = kernel(X_test, X_test)
            Ks = kernel(self.X, X_test)
        A = self.solve_triangular(Ks)
        return Kss - A.T @ A


def KalmanLogp(kernel, X, y, diag):
    """solvers/kalman.py:87-122 -- third independent formula for the log-probability."""
    X = np.asarray(X, dtype=np.float64)
    Pinf = kernel.stationary_covariance()
    J = Pinf.shape[0]
    m = np.zeros(J)
    P = Pinf.copy()
    ll = 0.0
    Xs = np.append(X[0], X[:-1])
    for k in range(X.shape[0]):
        A = kernel.transition_matrix(Xs[k], X[k]).T      # m2 = F m1 with F = T^T (quasisep.py:77-84)
        h = kernel.observation_model(X[k])
        m = A @ m
        P = A @ P @ A.T + (Pinf - A @ Pinf @ A.T)
        v = y[k] - h @ m
        S = h @ P @ h + diag[k]
        Kg = P @ h / S
        m = m + Kg * v
        P = P - np.outer(Kg, Kg) * S
        ll += -0.5 * (v * v / S + np.log(2 * np.pi * S))
    return ll


# ----------------------------------------------------------------------------
# gp.py
# ----------------------------------------------------------------------------
def _default_diag(reference):  # gp.py:388-393
    return np.sqrt(np.finfo(np.asarray(reference).dtype).eps)


class GaussianProcess:
    """gp.py:30-361 (constant / zero mean only)."""

    def __init__(self, kernel, X, *, diag=None, noise=None, mean=None, solver=None,
                 mean_value=None, covariance_value=None, **solver_kwargs):
        self.kernel = kernel
        self.X = np.asarray(X, dtype=np.float64)
        self.mean_const = 0.0 if mean is None else mean
        if mean_value is None:
            if callable(self.mean_const):
                mean_value = np.array([self.mean_const(x) for x in self.X], dtype=np.float64)
            else:
                mean_value = np.full(self.X.shape[0], float(self.mean_const))
        self.mean = mean_value
        self.num_data = mean_value.shape[0]
        if noise is None:
            diag = _default_diag(self.mean) if diag is None else diag
            noise = Diagonal(np.broadcast_to(np.asarray(diag, dtype=np.float64), self.mean.shape).copy())
        self.noise = noise
        if solver is None:                                               # gp.py:101-105
            solver = QuasisepSolver if isinstance(kernel, qs.Quasisep) else DirectSolver
        self.solver = solver(kernel, self.X, self.noise, covariance=covariance_value, **solver_kwargs)

    @property
    def loc(self):
        return self.mean

    @property
    def variance(self):
        return self.solver.variance()

    @property
    def covariance(self):
        return self.solver.covariance()

    def _get_alpha(self, y):  # gp.py:318-320
        return self.solver.solve_triangular(np.asarray(y, dtype=np.float64) - self.loc)

    def _compute_log_prob(self, alpha):  # gp.py:313-316
        with np.errstate(all="ignore"):
            loglike = -0.5 * np.sum(np.square(alpha)) - self.solver.normalization()
        return loglike if np.isfinite(loglike) else -np.inf

    def log_probability(self, y):  # gp.py:126-138
        return self._compute_log_prob(self._get_alpha(y))

    def condition(self, y, X_test=None, *, diag=None, noise=None, include_mean=True, kernel=None):
        """gp.py:140-223,322-361 ; returns (log_prob, conditioned GaussianProcess)."""
        y = np.asarray(y, dtype=np.float64)
        alpha = self._get_alpha(y)
        log_prob = self._compute_log_prob(alpha)
        alpha = self.solver.solve_triangular(alpha, transpose=True)      # gp.py:334
        if X_test is None:
            if kernel is None:
                mean_value = y - self.noise.matmul(alpha)                # gp.py:342-346
                if not include_mean:
                    mean_value = mean_value - self.loc
            else:
                mean_value = kernel.matmul(self.X, y=alpha)
                if include_mean:
                    mean_value = mean_value + self.loc
        else:
            X_test = np.asarray(X_test, dtype=np.float64)
            k = self.kernel if kernel is None else kernel
            mean_value = k.matmul(X_test, self.X, alpha)                 # gp.py:357
            if include_mean:
                if callable(self.mean_const):
                    mean_value = mean_value + np.array([self.mean_const(x) for x in X_test])
                else:
                    mean_value = mean_value + float(self.mean_const)
        if kernel is None:
            kernel = self.kernel
        if noise is None:
            diag = _default_diag(mean_value) if diag is None else diag
            noise = Diagonal(np.broadcast_to(np.asarray(diag, dtype=np.float64), mean_value.shape).copy())
        covariance_value = self.solver.condition(kernel, X_test, noise)  # gp.py:201
        if X_test is None:
            X_test = self.X
        gp = GaussianProcess(kernel, X_test, noise=noise, mean_value=mean_value,
                             covariance_value=covariance_value, solver=_PrecomputedDirect,
                             noise_in_covariance=getattr(self.solver, "last_condition_has_noise", True))
        return log_prob, gp

    def predict(self, y, X_test=None, *, kernel=None, include_mean=True, return_var=False, return_cov=False):
        _, cond = self.condition(y, X_test, kernel=kernel, include_mean=include_mean)  # gp.py:267
        if return_var:
            return cond.loc, cond.variance
        if return_cov:
            return cond.loc, cond.covariance
        return cond.loc

    def sample_from_normal(self, normal_samples):
        """gp.py:298-311 with the N(0,1) draws supplied (JAX threefry is not reproducible here)."""
        return self.mean + np.moveaxis(self.solver.dot_triangular(normal_samples), 0, -1)


class _PrecomputedDirect(DirectSolver):
    """DirectSolver given covariance_value (gp.py:208-221 -> direct.py:49-52).  The variance of the conditioned GP is
    kernels.Conditioned.evaluate_diag + noise.diagonal() (direct.py:49 with base.py:149-153), and
    Conditioned.evaluate_diag == diag(Kss - A^T A).  DirectSolver.condition adds the noise to Kss (direct.py:88-92),
    QuasisepSolver.condition's dense branch does not (solvers/quasisep/solver.py:131-139) -- so for a quasiseparable
    parent the conditioned GP's *variance* carries the noise while its *covariance* does not.  Pinned by
    tests/golden/reference_vectors.json (pred_var vs pred_cov_row0)."""

    def __init__(self, kernel, X, noise, *, covariance=None, noise_in_covariance=True):
        self.X = np.asarray(X, dtype=np.float64)
        self.variance_value = np.diag(covariance).copy()
        if not noise_in_covariance:
            self.variance_value = self.variance_value + noise.diagonal()
        self.covariance_value = covariance
        try:
            self.scale_tril = sla.cholesky(covariance, lower=True, check_finite=False)
        except sla.LinAlgError:
            self.scale_tril = np.full_like(covariance, np.nan)
