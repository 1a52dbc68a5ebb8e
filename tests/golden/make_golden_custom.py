"""Generate tests/golden/custom_kernel_vectors.json: a USER-DEFINED quasiseparable kernel -- the two-state non-reversible
`CausalFilter` of the reference's tests/test_kernels/test_quasisep_nonreversible.py, defined there as a Quasisep subclass with its
own design_matrix / stationary_covariance / observation_model / coord_to_sortable / transition_matrix -- run through the
UNMODIFIED reference (sources under /root/reference/src over tests/golden/jaxshim) with pytree coordinates (time, channel).
The product takes the same model with (N, 2) array coordinates (tests/test_host_layer_reference_tests.py::CausalFilter).
Run from the repo root:   python tests/golden/make_golden_custom.py
"""

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

import refimport  # noqa: E402

TIME = [0.0, 0.3, 0.8, 1.4, 2.2, 3.1]
CHANNEL = [0, 1, 0, 1, 1, 0]
TIME_TEST = [0.1, 0.9, 1.8, 2.8]
CHANNEL_TEST = [1, 0, 1, 0]
Y = [0.2, -0.1, 0.3, 0.15, -0.2, 0.05]
DIAG = 0.05


def main():
    refimport.install()
    import jax.numpy as jnp
    import jax.scipy as jsp
    from tinygp import GaussianProcess
    from tinygp.kernels import quasisep
    from tinygp.solvers import DirectSolver, QuasisepSolver

    class CausalFilter(quasisep.Quasisep):
        def design_matrix(self):
            return jnp.array([[-1.0, 0.0], [0.8, -2.0]])

        def stationary_covariance(self):
            return jnp.array([[0.5, 2.0 / 15.0], [2.0 / 15.0, 91.0 / 300.0]])

        def observation_model(self, X):
            _time, channel = X
            return jnp.eye(2)[channel]

        def coord_to_sortable(self, X):
            time, _channel = X
            return time

        def transition_matrix(self, X1, X2):
            return jsp.linalg.expm(self.design_matrix().T * (X2[0] - X1[0]))

    kernel = CausalFilter()
    X = (jnp.array(TIME), jnp.array(CHANNEL))
    Xt = (jnp.array(TIME_TEST), jnp.array(CHANNEL_TEST))
    y = jnp.array(Y)
    diag = jnp.full(len(TIME), DIAG)
    out = {"generator": "tests/golden/make_golden_custom.py", "time": TIME, "channel": CHANNEL, "time_test": TIME_TEST,
           "channel_test": CHANNEL_TEST, "y": Y, "diag": DIAG}
    out["K"] = np.asarray(kernel(X, X)).tolist()
    out["K_cross"] = np.asarray(kernel(Xt, X)).tolist()
    out["symm_qsm_dense"] = np.asarray(kernel.to_symm_qsm(X).to_dense()).tolist()
    out["cross_matmul"] = np.asarray(kernel.matmul(Xt, X, y)).tolist()
    for name, solver in (("quasisep", QuasisepSolver), ("direct", DirectSolver)):
        gp = GaussianProcess(kernel, X, diag=diag, solver=solver)
        rec = {"log_probability": float(gp.log_probability(y)), "covariance": np.asarray(gp.covariance).tolist()}
        c_in, c_test = gp.condition(y), gp.condition(y, X_test=Xt)
        rec["cond_in_loc"] = np.asarray(c_in.gp.loc).tolist()
        rec["cond_in_cov"] = np.asarray(c_in.gp.covariance).tolist()
        rec["cond_test_loc"] = np.asarray(c_test.gp.loc).tolist()
        rec["cond_test_cov"] = np.asarray(c_test.gp.covariance).tolist()
        out[name] = rec
        print(name, rec["log_probability"])
    # a Wrapper with a coordinate-dependent observation model (quasisep.py:218-238; the multiband construction of the
    # reference's documentation): amplitude per band times the wrapped kernel's observation model
    class Multiband(quasisep.Wrapper):
        amplitudes: jnp.ndarray

        def coord_to_sortable(self, X):
            return X[0]

        def observation_model(self, X):
            return self.amplitudes[X[1]] * self.kernel.observation_model(X[0])

    rng = np.random.default_rng(11)
    t = np.sort(rng.uniform(0, 10, 40))
    band = rng.integers(0, 2, 40)
    tt = rng.uniform(-1, 11, 6)
    band_t = np.array([0, 1, 1, 0, 1, 0])
    ym = np.sin(t)
    base = quasisep.Matern32(1.5) + 0.5 * quasisep.SHO(omega=1.2, quality=2.0)
    mb = Multiband(kernel=base, amplitudes=jnp.array([1.0, 0.6]))
    Xm, Xmt = (jnp.asarray(t), jnp.asarray(band)), (jnp.asarray(tt), jnp.asarray(band_t))
    rec = {"t": t.tolist(), "band": band.tolist(), "t_test": tt.tolist(), "band_test": band_t.tolist(), "y": ym.tolist(),
           "amplitudes": [1.0, 0.6], "K": np.asarray(mb(Xm, Xm)).tolist(), "K_cross": np.asarray(mb(Xmt, Xm)).tolist()}
    gp = GaussianProcess(mb, Xm, diag=0.1)
    rec["log_probability"] = float(gp.log_probability(ym))
    c_in, c_test = gp.condition(ym), gp.condition(ym, X_test=Xmt)
    rec["cond_in_loc"], rec["cond_in_var"] = np.asarray(c_in.gp.loc).tolist(), np.asarray(c_in.gp.variance).tolist()
    rec["cond_test_loc"], rec["cond_test_cov"] = np.asarray(c_test.gp.loc).tolist(), np.asarray(c_test.gp.covariance).tolist()
    out["multiband"] = rec
    print("multiband", rec["log_probability"])
    with open(os.path.join(HERE, "custom_kernel_vectors.json"), "w") as f:
        json.dump(out, f, indent=0)
    print("wrote custom_kernel_vectors.json")


if __name__ == "__main__":
    main()
