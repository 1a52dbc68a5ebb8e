/* Plain-C restatement of the reference's SEQUENTIAL quasiseparable recursions -- TEST ORACLE / CPU
 * BASELINE ONLY (see oracle/__init__.py; nothing under tinygp_b200/ links this).
 *
 *   qs_cholesky      src/tinygp/solvers/quasisep/ops.py:352-365
 *   qs_lower_solve   src/tinygp/solvers/quasisep/ops.py:463-472   (single right-hand side)
 *   log-probability  src/tinygp/gp.py:313-320 with solvers/quasisep/solver.py:90-93
 *
 * Generators d (n), p (n,J), q (n,J), a (n,J,J) are supplied by the caller (NumPy restatement of
 * kernels/quasisep.py:102-116), row-major.  Pinned through the NumPy oracle (tests/test_oracle_quasisep.py compares
 * the two), which is pinned to reference-generated goldens (oracle/__init__.py).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define MAXJ 16

/* returns 0, or k+1 of the first non-positive pivot */
int64_t qs_cholesky(int64_t n, int J, const double* d, const double* p, const double* q, const double* a,
                    double* c, double* w) {
    double f[MAXJ * MAXJ], tmp[MAXJ * MAXJ];
    int64_t bad = 0;
    memset(f, 0, sizeof(f));
    for (int64_t k = 0; k < n; ++k) {
        const double* pk = p + k * J;
        const double* qk = q + k * J;
        const double* ak = a + k * J * J;
        double quad = 0.0;
        for (int j = 0; j < J; ++j) { /* pk @ fp @ pk */
            double s = 0.0;
            for (int i = 0; i < J; ++i) s += pk[i] * f[i * J + j];
            quad += s * pk[j];
        }
        const double c2 = d[k] - quad;
        if (!(c2 > 0.0) && !bad) bad = k + 1;
        const double ck = sqrt(c2);
        for (int i = 0; i < J; ++i) /* tmp = fp @ ak.T */
            for (int j = 0; j < J; ++j) {
                double s = 0.0;
                for (int l = 0; l < J; ++l) s += f[i * J + l] * ak[j * J + l];
                tmp[i * J + j] = s;
            }
        double* wk = w + k * J;
        for (int j = 0; j < J; ++j) { /* wk = (qk - pk @ tmp) / ck */
            double s = 0.0;
            for (int i = 0; i < J; ++i) s += pk[i] * tmp[i * J + j];
            wk[j] = (qk[j] - s) / ck;
        }
        for (int i = 0; i < J; ++i) /* fk = ak @ tmp + outer(wk, wk) */
            for (int j = 0; j < J; ++j) {
                double s = 0.0;
                for (int l = 0; l < J; ++l) s += ak[i * J + l] * tmp[l * J + j];
                f[i * J + j] = s + wk[i] * wk[j];
            }
        c[k] = ck;
    }
    return bad;
}

void qs_lower_solve(int64_t n, int J, const double* c, const double* p, const double* w, const double* a,
                    const double* x, double* y) {
    double f[MAXJ], nf[MAXJ];
    memset(f, 0, sizeof(f));
    for (int64_t k = 0; k < n; ++k) {
        const double* pk = p + k * J;
        const double* wk = w + k * J;
        const double* ak = a + k * J * J;
        double s = 0.0;
        for (int j = 0; j < J; ++j) s += pk[j] * f[j];
        const double yk = (x[k] - s) / c[k];
        for (int i = 0; i < J; ++i) {
            double v = 0.0;
            for (int j = 0; j < J; ++j) v += ak[i * J + j] * f[j];
            nf[i] = v + wk[i] * yk;
        }
        memcpy(f, nf, sizeof(double) * J);
        y[k] = yk;
    }
}

/* -0.5 sum(alpha^2) - sum(log c) - n/2 log(2 pi);  c, w, alpha are caller-provided scratch */
double qs_log_probability(int64_t n, int J, const double* d, const double* p, const double* q, const double* a,
                          const double* resid, double* c, double* w, double* alpha) {
    const int64_t bad = qs_cholesky(n, J, d, p, q, a, c, w);
    qs_lower_solve(n, J, c, p, w, a, resid, alpha);
    double ss = 0.0, ld = 0.0;
    for (int64_t k = 0; k < n; ++k) {
        ss += alpha[k] * alpha[k];
        ld += log(c[k]);
    }
    const double lp = -0.5 * ss - ld - 0.5 * (double)n * log(2.0 * M_PI);
    if (bad || !isfinite(lp)) return -INFINITY;
    return lp;
}
