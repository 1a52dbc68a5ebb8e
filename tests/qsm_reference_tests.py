"""The reference's own tests of the QSM classes -- tests/test_solvers/test_quasisep/test_core.py (:120-374) -- restated
against tinygp_b200.solvers.quasisep.core: same matrices (get_matrices :61-117), same assertions, the reference's
tolerance (assert_allclose rtol = atol = 5e-7 for float64, src/tinygp/test_utils.py:9-26).  Backend-agnostic: run on the
CPU over the host build of the device source (test_qsm_reference_tests_host.py) and on the GPU (test_qsm_gpu.py)."""

from itertools import combinations

import numpy as np

from tinygp_b200.solvers.quasisep.core import (DiagQSM, LowerTriQSM, SquareQSM, StrictLowerTriQSM, StrictUpperTriQSM,
                                               SymmQSM)


def assert_allclose(a, b, rtol=5e-7, atol=5e-7):
    np.testing.assert_allclose(np.asarray(a), np.asarray(b), rtol=rtol, atol=atol)


def get_matrices(name):                                           # test_core.py:61-117
    N = 100
    random = np.random.default_rng(1234)
    diag = np.exp(random.normal(size=N))
    if name == "random":
        J = 5
        p = random.normal(size=(N, J))
        q = random.normal(size=(N, J))
        a = np.repeat(np.eye(J)[None, :, :], N, axis=0)
        l = np.tril(p @ q.T, -1)
        u = np.triu(q @ p.T, 1)
        diag = diag + np.sum(p * q, axis=1)
    elif name == "celerite":
        t = np.sort(random.uniform(0, 10, N))
        a = np.array([1.0, 2.5]); b = np.array([0.5, 1.5]); c = np.array([1.2, 0.5]); d = np.array([0.5, 0.1])
        tau = np.abs(t[:, None] - t[None, :])[:, :, None]
        K = np.sum(np.exp(-c[None, None] * tau) * (a[None, None] * np.cos(d[None, None] * tau)
                                                   + b[None, None] * np.sin(d[None, None] * tau)), axis=-1)
        K += np.diag(diag)
        diag = np.diag(K)
        l = np.tril(K, -1)
        u = np.triu(K, 1)
        cos = np.cos(d[None] * t[:, None])
        sin = np.sin(d[None] * t[:, None])
        p = np.concatenate((a[None] * cos + b[None] * sin, a[None] * sin - b[None] * cos), axis=1)
        q = np.concatenate((cos, sin), axis=1)
        c = np.append(c, c)
        dt = np.append(0, np.diff(t))
        a = np.stack([np.diag(v) for v in np.exp(-c[None] * dt[:, None])], axis=0)
        p = np.einsum("ni,nij->nj", p, a)
    else:
        raise AssertionError()
    v = random.normal(size=N)
    m = random.normal(size=(N, 4))
    return diag, p, q, a, v, m, l, u


def some_nice_matrices():                                         # test_core.py:36-58
    diag1, p1, q1, a1, _, _, _, _ = get_matrices("celerite")
    diag2, p2, q2, a2, _, _, _, _ = get_matrices("random")
    mat1 = LowerTriQSM(diag=DiagQSM(diag1), lower=StrictLowerTriQSM(p=p1, q=q1, a=a1))
    mat2 = SquareQSM(diag=DiagQSM(diag2), lower=StrictLowerTriQSM(p=p2, q=q2, a=a2), upper=StrictUpperTriQSM(p=p2, q=q2, a=a2))
    mat3 = SquareQSM(diag=DiagQSM(diag1), lower=StrictLowerTriQSM(p=p1, q=q1, a=a1),
                     upper=StrictUpperTriQSM(p=np.zeros_like(p2), q=np.zeros_like(q2), a=a2))
    mat4 = SquareQSM(diag=DiagQSM(diag1), lower=StrictLowerTriQSM(p=p1, q=q1, a=a1), upper=StrictUpperTriQSM(p=p2, q=q2, a=a2))
    return mat1, mat2, mat3, mat4


def check_quasisep_def():                                         # test_core.py:120-154
    random = np.random.default_rng(2022)
    n, m1, m2 = 17, 3, 5
    d = random.normal(size=n)
    p = random.normal(size=(n, m1)); q = random.normal(size=(n, m1)); a = random.normal(size=(n, m1, m1))
    g = random.normal(size=(n, m2)); h = random.normal(size=(n, m2)); b = random.normal(size=(n, m2, m2))
    m = SquareQSM(diag=DiagQSM(d=d), lower=StrictLowerTriQSM(p=p, q=q, a=a), upper=StrictUpperTriQSM(p=g, q=h, a=b)).to_dense()

    def get_value(i, j):
        if i == j:
            return d[i]
        if j < i:
            tmp = np.copy(q[j])
            for k in range(j + 1, i):
                tmp = a[k] @ tmp
            return p[i] @ tmp
        tmp = np.copy(h[i])
        for k in range(i + 1, j):
            tmp = tmp @ b[k].T
        return tmp @ g[j]

    for i in range(n):
        for j in range(n):
            assert_allclose(get_value(i, j), m[i, j])


def check_strict_tri_matmul(name, parallel):                      # :157-172
    _, p, q, a, v, m, l, u = get_matrices(name)
    mat = StrictLowerTriQSM(p=p, q=q, a=a)
    assert_allclose(mat.to_dense(), l)
    assert_allclose(mat.T.to_dense(), u)
    assert_allclose(mat.matmul(v, parallel=parallel), l @ v)
    assert_allclose(mat.T.matmul(v, parallel=parallel), u @ v)
    assert_allclose(mat.matmul(m, parallel=parallel), l @ m)
    assert_allclose(mat.T.matmul(m, parallel=parallel), u @ m)


def check_tri_matmul(name, parallel):                             # :175-190
    diag, p, q, a, v, m, l, _ = get_matrices(name)
    mat = LowerTriQSM(diag=DiagQSM(diag), lower=StrictLowerTriQSM(p=p, q=q, a=a))
    dense = l + np.diag(diag)
    assert_allclose(mat.to_dense(), dense)
    assert_allclose(mat.T.to_dense(), dense.T)
    assert_allclose(mat.matmul(v, parallel=parallel), dense @ v)
    assert_allclose(mat.T.matmul(v, parallel=parallel), dense.T @ v)
    assert_allclose(mat.matmul(m, parallel=parallel), dense @ m)
    assert_allclose(mat.T.matmul(m, parallel=parallel), dense.T @ m)


def _square(symm, diag, p, q, a):
    if symm:
        return SymmQSM(diag=DiagQSM(diag), lower=StrictLowerTriQSM(p=p, q=q, a=a))
    return SquareQSM(diag=DiagQSM(diag), lower=StrictLowerTriQSM(p=p, q=q, a=a), upper=StrictUpperTriQSM(p=p, q=q, a=a))


def check_square_matmul(symm, name, parallel):                    # :193-216
    diag, p, q, a, v, m, l, u = get_matrices(name)
    mat = _square(symm, diag, p, q, a)
    dense = mat.to_dense()
    assert_allclose(np.tril(dense, -1), l)
    assert_allclose(np.triu(dense, 1), u)
    assert_allclose(np.diag(dense), diag)
    assert_allclose(mat.matmul(v, parallel=parallel), dense @ v)
    assert_allclose(mat.matmul(m, parallel=parallel), dense @ m)
    assert_allclose(v.T @ mat, v.T @ dense)
    assert_allclose(m.T @ mat, m.T @ dense)


def check_tri_inv():                                              # :219-226 (celerite)
    diag, p, q, a, _, _, _, _ = get_matrices("celerite")
    mat = LowerTriQSM(diag=DiagQSM(diag), lower=StrictLowerTriQSM(p=p, q=q, a=a))
    dense = mat.to_dense()
    minv = mat.inv()
    assert_allclose(minv.to_dense(), np.linalg.inv(dense))
    assert_allclose(minv.matmul(dense), np.eye(len(diag)))


def check_tri_solve(parallel):                                    # :229-244 (celerite)
    diag, p, q, a, v, m, _, _ = get_matrices("celerite")
    mat = LowerTriQSM(diag=DiagQSM(diag), lower=StrictLowerTriQSM(p=p, q=q, a=a))
    dense = mat.to_dense()
    assert_allclose(mat.solve(v, parallel=parallel), np.linalg.solve(dense, v))
    assert_allclose(mat.solve(m, parallel=parallel), np.linalg.solve(dense, m))
    assert_allclose(mat.T.solve(v, parallel=parallel), np.linalg.solve(dense.T, v))
    assert_allclose(mat.T.solve(m, parallel=parallel), np.linalg.solve(dense.T, m))
    assert_allclose(mat.inv().solve(v, parallel=parallel), dense @ v)
    assert_allclose(mat.inv().solve(m, parallel=parallel), dense @ m)
    assert_allclose(mat.T.inv().solve(v, parallel=parallel), dense.T @ v)
    assert_allclose(mat.T.inv().solve(m, parallel=parallel), dense.T @ m)


def check_square_inv(symm, name):                                 # :247-281
    diag, p, q, a, _, _, l, u = get_matrices(name)
    mat = _square(symm, diag, p, q, a)
    dense = mat.to_dense()
    assert_allclose(np.tril(dense, -1), l)
    assert_allclose(np.triu(dense, 1), u)
    assert_allclose(np.diag(dense), diag)
    minv = mat.inv()
    assert_allclose(minv.to_dense(), np.linalg.inv(dense))
    assert_allclose(minv.matmul(dense), np.eye(len(diag)))
    if not symm:
        assert_allclose(minv.lower.p, minv.upper.p)
        assert_allclose(minv.lower.q, minv.upper.q)
        assert_allclose(minv.lower.a, minv.upper.a)
    mat2 = minv.inv()
    assert_allclose(mat2.to_dense(), dense, rtol=1e-4)


def check_gram(name):                                             # :284-308
    diag, p, q, a, _, _, _, _ = get_matrices(name)
    mat = _square(False, diag, p, q, a)
    dense = mat.to_dense()
    assert_allclose(mat.gram().to_dense(), dense.T @ dense)
    mat = mat.inv()
    dense = mat.to_dense()
    assert_allclose(mat.gram().to_dense(), dense.T @ dense)
    mat = SquareQSM(diag=DiagQSM(diag), lower=StrictLowerTriQSM(p=p, q=q, a=a),
                    upper=StrictUpperTriQSM(p=np.zeros_like(p), q=np.zeros_like(q), a=np.zeros_like(a)))
    dense = mat.to_dense()
    assert_allclose(mat.gram().to_dense(), dense.T @ dense)


def check_cholesky(parallel):                                     # :311-325 (celerite)
    diag, p, q, a, v, m, _, _ = get_matrices("celerite")
    mat = SymmQSM(diag=DiagQSM(diag), lower=StrictLowerTriQSM(p=p, q=q, a=a))
    dense = mat.to_dense()
    chol = mat.cholesky(parallel=parallel)
    assert_allclose(chol.to_dense(), np.linalg.cholesky(dense))
    mat = mat.inv()
    dense = mat.to_dense()
    chol = mat.cholesky(parallel=parallel)
    assert_allclose(chol.to_dense(), np.linalg.cholesky(dense))
    dense = chol.to_dense()
    assert_allclose(chol.solve(v, parallel=parallel), np.linalg.solve(dense, v))
    assert_allclose(chol.solve(m, parallel=parallel), np.linalg.solve(dense, m))


def _check_product(mat1, mat2):
    mat = mat1 @ mat2
    a = mat.to_dense()
    b = mat1.to_dense() @ mat2.to_dense()
    assert_allclose(np.diag(a), np.diag(b))
    assert_allclose(np.tril(a, -1), np.tril(b, -1))
    assert_allclose(np.triu(a, 1), np.triu(b, 1))


def check_tri_qsmul():                                            # :328-346
    mat1, mat2, mat3, mat4 = some_nice_matrices()
    minv = mat1.inv()
    mTinv = mat1.T.inv()
    for m in [mat2, mat3, mat4, mat2.inv()]:
        _check_product(mat1, m)
        _check_product(minv, m)
        _check_product(mat1.T, m)
        _check_product(mTinv, m)


def check_square_qsmul():                                         # :349-361
    mat1, mat2, mat3, mat4 = some_nice_matrices()
    mat1 = mat1 + mat1.lower.transpose()
    for m1, m2 in combinations([mat1, mat2, mat3, mat4, mat1.inv(), mat2.inv()], 2):
        _check_product(m1, m2)


def check_ops():                                                  # :364-378
    mat1, mat2, mat3, mat4 = some_nice_matrices()

    def check(mat1, mat2):
        for m1, m2 in combinations([mat1, mat2, mat1.lower, mat2.lower], 2):
            a = m1.to_dense()
            b = m2.to_dense()
            assert_allclose((-m1).to_dense(), -a)
            assert_allclose((m1 + m2).to_dense(), a + b)
            assert_allclose((m1 - m2).to_dense(), a - b)
            assert_allclose((m1 * m2).to_dense(), a * b)
            assert_allclose((2.5 * m1).to_dense(), 2.5 * a)

    for m1, m2 in combinations([mat1, mat2, mat3, mat4, mat1.inv(), mat2.inv()], 2):
        check(m1, m2)
