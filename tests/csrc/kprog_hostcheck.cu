// TEST INFRASTRUCTURE: the kernel-program interpreter of the dense CUDA path (tinygp_b200/csrc/kprog.cuh:
// parse_prog_impl + kprog_eval, the code build_rect_kernel and the GEMM / int8 generator epilogues inline) compiled for
// the CPU, so `-m "not gpu"` tests can compare the SAME source with the reference goldens and the oracle.
// Built on demand by tests/test_device_code_on_host.py; never linked into libb200gp.so.
#include "../../tinygp_b200/csrc/kprog.cuh"

extern "C" {

// out[i * n2 + j] = k(X1_i, X2_j); returns 0, or 2 with the parser's message in err (<= 255 chars)
int hostcheck_kernel_matrix(const double* prog, int n_rows, const double* X1, int64_t n1, const double* X2, int64_t n2,
                            int ndim, double* out, char* err) {
    try {
        const KProg P = parse_prog_impl(prog, n_rows, ndim);
        for (int64_t i = 0; i < n1; ++i)
            for (int64_t j = 0; j < n2; ++j) {
                const double* xa = X1 + i * ndim;
                const double* xb = X2 + j * ndim;
                out[i * n2 + j] = kprog_eval(P, ndim, [&](int d) { return xa[d] - xb[d]; });
            }
        return 0;
    } catch (const std::exception& e) {
        snprintf(err, 256, "%s", e.what());
        return 2;
    }
}

// the same through the sum-of-products normal form (kprog_to_fast + kfast_eval, what build_rect_kernel_t<true> runs);
// returns 4 if the program has no normal form (the product then interprets it)
int hostcheck_kernel_matrix_fast(const double* prog, int n_rows, const double* X1, int64_t n1, const double* X2, int64_t n2,
                                 int ndim, double* out, char* err) {
    try {
        const KProg P = parse_prog_impl(prog, n_rows, ndim);
        KFast F{};
        if (!kprog_to_fast(P, F)) return 4;
        for (int64_t i = 0; i < n1; ++i)
            for (int64_t j = 0; j < n2; ++j) {
                const double* xa = X1 + i * ndim;
                const double* xb = X2 + j * ndim;
                out[i * n2 + j] = kfast_eval(F, ndim, [&](int d) { return xa[d] - xb[d]; });
            }
        return 0;
    } catch (const std::exception& e) {
        snprintf(err, 256, "%s", e.what());
        return 2;
    }
}

// parse -> kprog_encode -> rows; returns the number of rows written (<= max_rows) or -1
int hostcheck_reencode(const double* prog, int n_rows, int ndim, double* out, int max_rows) {
    try {
        const std::vector<double> rows = kprog_encode(parse_prog_impl(prog, n_rows, ndim));
        const int n = (int)(rows.size() / B200GP_PROG_STRIDE);
        if (n > max_rows) return -1;
        for (size_t i = 0; i < rows.size(); ++i) out[i] = rows[i];
        return n;
    } catch (const std::exception&) { return -1; }
}

int hostcheck_kernel_diag(const double* prog, int n_rows, int ndim, double* out, char* err) {
    try {
        const KProg P = parse_prog_impl(prog, n_rows, ndim);
        *out = kprog_eval_zero(P);
        return 0;
    } catch (const std::exception& e) {
        snprintf(err, 256, "%s", e.what());
        return 2;
    }
}

}  // extern "C"
