// The kernel-program interpreter (device side of Kernel.__call__, base.py:84-103): evaluation is __host__ __device__
// so that tests/csrc/kprog_hostcheck.cu can run the SAME source on the CPU against the reference goldens.
// Included by dense.cu (which exports parse_prog for the other translation units).
#pragma once
#include "common.cuh"

#define MAX_NDIM 16
#define SQRT3 1.7320508075688772
#define SQRT5 2.23606797749979
#define PI_D 3.141592653589793

// one stationary leaf given the two distances of its metric (stationary.py:76-235 as cited per branch)
__host__ __device__ __forceinline__ double kprog_leaf(const int op, const bool l2, const double p0, const double p1,
                                                     const double l1, const double l2sq) {
    if (op == B200GP_OP_EXPSQUARED || op == B200GP_OP_RATIONALQUADRATIC) {
        // squared_distance / square(scale)   (stationary.py:105,234 ; distance.py:30-38,58-59)
        const double sq = l2 ? l2sq : l1 * l1;
        const double r2 = sq / (p0 * p0);
        if (op == B200GP_OP_EXPSQUARED) return exp(-0.5 * r2);
        return pow(1.0 + 0.5 * r2 / p1, -p1);
    }
    // distance (distance.py:44-45 / 51-56: sqrt with the r2==0 guard)
    const double dist = l2 ? ((l2sq == 0.0) ? l1 : sqrt(l2sq)) : l1;
    if (op == B200GP_OP_EXPCOS || op == B200GP_OP_EXPSIN) {
        double sn, cs;
        sincos(p1 * dist, &sn, &cs);
        return exp(-p0 * dist) * ((op == B200GP_OP_EXPCOS) ? cs : sn);
    }
    const double r = dist / p0;
    if (op == B200GP_OP_EXP) return exp(-r);
    if (op == B200GP_OP_MATERN32) {
        const double arg = SQRT3 * r;
        return (1.0 + arg) * exp(-arg);
    }
    if (op == B200GP_OP_MATERN52) {
        const double arg = SQRT5 * r;
        return (1.0 + arg + arg * arg / 3.0) * exp(-arg);
    }
    if (op == B200GP_OP_COSINE) return cos(2.0 * PI_D * r);
    const double s = sin(PI_D * r);   // EXPSINESQUARED
    return exp(-p1 * (s * s));
}

// ---- sum-of-products normal form of a kernel program (no interpreter stack: everything stays in registers) -----------
// K = sum_t coef[t] * prod_{i in mask[t]} leaf_i , at most 4 distinct identity-metric leaves and 4 terms; programs that do
// not fit (input transforms, a leaf multiplied by itself, more leaves / terms) keep the generic interpreter below.
#define KFAST_MAX 4
struct KFast {
    int nleaf, nterm;
    int op[KFAST_MAX], l2[KFAST_MAX];
    double p0[KFAST_MAX], p1[KFAST_MAX];
    double coef[KFAST_MAX];
    int mask[KFAST_MAX];
};

template <typename F>
__host__ __device__ __forceinline__ double kfast_eval(const KFast& K, int ndim, F diff) {
    double l1 = 0.0, l2sq = 0.0;
    for (int d = 0; d < ndim; ++d) {
        const double df = diff(d);
        l1 += fabs(df);
        l2sq += df * df;
    }
    double v[KFAST_MAX];
#pragma unroll
    for (int i = 0; i < KFAST_MAX; ++i) {
        v[i] = 1.0;
        if (i < K.nleaf) v[i] = kprog_leaf(K.op[i], K.l2[i] != 0, K.p0[i], K.p1[i], l1, l2sq);
    }
    double acc = 0.0;
#pragma unroll
    for (int t = 0; t < KFAST_MAX; ++t) {
        if (t < K.nterm) {
            double prod = K.coef[t];
#pragma unroll
            for (int i = 0; i < KFAST_MAX; ++i)
                if ((K.mask[t] >> i) & 1) prod *= v[i];
            acc += prod;
        }
    }
    return acc;
}

// host: postfix program -> normal form by expanding the expression as a polynomial in its leaves
static inline bool kprog_to_fast(const KProg& P, KFast& out) {
    struct Poly { int n; double coef[KFAST_MAX]; int mask[KFAST_MAX]; };
    Poly st[8];
    int sp = 0;
    KFast K{};
    for (int i = 0; i < P.n; ++i) {
        const int op = P.op[i];
        if (op == B200GP_OP_ADD || op == B200GP_OP_MUL) {
            if (sp < 2) return false;
            const Poly b = st[--sp], a = st[sp - 1];
            Poly r{};
            if (op == B200GP_OP_ADD) {
                r = a;
                for (int j = 0; j < b.n; ++j) {
                    int hit = -1;
                    for (int k = 0; k < r.n; ++k) if (r.mask[k] == b.mask[j]) hit = k;
                    if (hit >= 0) { r.coef[hit] += b.coef[j]; continue; }
                    if (r.n >= KFAST_MAX) return false;
                    r.coef[r.n] = b.coef[j]; r.mask[r.n] = b.mask[j]; ++r.n;
                }
            } else {
                for (int j = 0; j < a.n; ++j)
                    for (int k = 0; k < b.n; ++k) {
                        if (a.mask[j] & b.mask[k]) return false;      // a leaf squared: not representable
                        const int mk = a.mask[j] | b.mask[k];
                        const double cf = a.coef[j] * b.coef[k];
                        int hit = -1;
                        for (int q = 0; q < r.n; ++q) if (r.mask[q] == mk) hit = q;
                        if (hit >= 0) { r.coef[hit] += cf; continue; }
                        if (r.n >= KFAST_MAX) return false;
                        r.coef[r.n] = cf; r.mask[r.n] = mk; ++r.n;
                    }
            }
            st[sp - 1] = r;
            continue;
        }
        if (sp >= 8) return false;
        Poly r{};
        r.n = 1;
        if (op == B200GP_OP_CONST) {
            r.coef[0] = P.p0[i]; r.mask[0] = 0;
        } else {
            if (P.metric[i] != 0 || K.nleaf >= KFAST_MAX) return false;
            K.op[K.nleaf] = op; K.l2[K.nleaf] = (P.dist[i] == B200GP_DIST_L2) ? 1 : 0;
            K.p0[K.nleaf] = P.p0[i]; K.p1[K.nleaf] = P.p1[i];
            r.coef[0] = 1.0; r.mask[0] = 1 << K.nleaf;
            ++K.nleaf;
        }
        st[sp++] = r;
    }
    if (sp != 1) return false;
    K.nterm = st[0].n;
    for (int t = 0; t < K.nterm; ++t) { K.coef[t] = st[0].coef[t]; K.mask[t] = st[0].mask[t]; }
    out = K;
    return true;
}

// l1 = sum_d |x1_d - x2_d| ; l2sq = sum_d (x1_d - x2_d)^2   (explicit differences, distance.py:45,59)
// `diff(d)` returns x1_d - x2_d; it is re-evaluated (not cached in a register array) so that the DMMA
// GEMM epilogue, which inlines this, keeps its register budget.  Leaves that carry a linear input
// transform (transforms.py:57-161) measure their distance on M (x1 - x2).
template <typename F>
__host__ __device__ __forceinline__ double kprog_eval(const KProg& P, int ndim, F diff) {
    double l1_id = 0.0, l2_id = 0.0;
    for (int d = 0; d < ndim; ++d) {
        const double df = diff(d);
        l1_id += fabs(df);
        l2_id += df * df;
    }
    double st[8];
    int sp = 0;
    for (int i = 0; i < P.n; ++i) {
        const int op = P.op[i];
        if (op == B200GP_OP_ADD) {
            --sp;
            st[sp - 1] = st[sp - 1] + st[sp];
            continue;
        }
        if (op == B200GP_OP_MUL) {
            --sp;
            st[sp - 1] = st[sp - 1] * st[sp];
            continue;
        }
        const double p0 = P.p0[i], p1 = P.p1[i];
        double v;
        if (op == B200GP_OP_CONST) {
            v = p0;
        } else {
            double l1 = l1_id, l2sq = l2_id;
            const int m = P.metric[i];
            if (m > 0) {
                l1 = 0.0;
                l2sq = 0.0;
                const double* Mm = P.M[m - 1];
                for (int r = 0; r < P.mrows[m - 1]; ++r) {
                    double z = 0.0;
                    for (int c = 0; c < ndim; ++c) z += Mm[r * B200GP_METRIC_MAX_DIM + c] * diff(c);
                    l1 += fabs(z);
                    l2sq += z * z;
                }
            }
            v = kprog_leaf(op, P.dist[i] == B200GP_DIST_L2, p0, p1, l1, l2sq);
        }
        st[sp++] = v;
    }
    return st[0];
}

// k(x, x): every difference is zero, with or without a transform
__host__ __device__ __forceinline__ double kprog_eval_zero(const KProg& P) {
    return kprog_eval(P, 0, [](int) { return 0.0; });
}

// ndim < 0: the caller has no coordinates (k(x, x) only) and metrics are accepted for any width
static inline KProg parse_prog_impl(const double* prog, int n_rows, int ndim) {
    if (n_rows <= 0 || n_rows > B200GP_PROG_MAX_ROWS) throw GpError("kernel program: bad length");
    KProg P{};
    int depth = 0, row = 0;
    // metric definitions come first
    while (row < n_rows && (int)prog[(size_t)row * B200GP_PROG_STRIDE] == B200GP_OP_METRIC) {
        const double* q = prog + (size_t)row * B200GP_PROG_STRIDE;
        const int id = (int)q[1], r = (int)q[2], c = (int)q[3];
        if (id != P.nmetric + 1 || id > B200GP_PROG_MAX_METRICS)
            throw GpError("kernel program: metric ids must be 1..3 in order (at most 3 distinct input transforms)");
        if (r < 1 || r > B200GP_METRIC_MAX_DIM || c < 1 || c > B200GP_METRIC_MAX_DIM)
            throw GpError("kernel program: transformed kernels support at most 8 input / output dimensions");
        if (ndim >= 0 && c != ndim) throw GpError("kernel program: transform width does not match ndim");
        if (P.nmetric > 0 && c != P.mcols) throw GpError("kernel program: inconsistent transform widths");
        const int nd_rows = (r * c + B200GP_PROG_STRIDE - 1) / B200GP_PROG_STRIDE;
        if (row + 1 + nd_rows > n_rows) throw GpError("kernel program: truncated metric definition");
        const double* data = q + B200GP_PROG_STRIDE;
        for (int i = 0; i < r; ++i)
            for (int j = 0; j < c; ++j) {
                const double v = data[i * c + j];
                if (!(v == v) || v - v != 0.0) throw GpError("kernel program: non-finite transform matrix");
                P.M[id - 1][i * B200GP_METRIC_MAX_DIM + j] = v;
            }
        P.mrows[id - 1] = r;
        P.mcols = c;
        P.nmetric = id;
        row += 1 + nd_rows;
    }
    const int n_instr = n_rows - row;
    if (n_instr <= 0 || n_instr > B200GP_PROG_MAX_INSTR) throw GpError("kernel program: bad length");
    P.n = n_instr;
    for (int i = 0; i < n_instr; ++i) {
        const double* q = prog + (size_t)(row + i) * B200GP_PROG_STRIDE;
        P.op[i] = (int)q[0];
        const int dcode = (int)q[1];
        P.p0[i] = q[2];
        P.p1[i] = q[3];
        const int op = P.op[i];
        if (op == B200GP_OP_ADD || op == B200GP_OP_MUL) {
            if (depth < 2) throw GpError("kernel program: stack underflow");
            --depth;
        } else if (op >= B200GP_OP_CONST && op <= B200GP_OP_EXPSIN) {
            if (dcode < 0 || (dcode >> 1) > P.nmetric) throw GpError("kernel program: leaf refers to an undefined metric");
            P.dist[i] = dcode & 1;
            P.metric[i] = (op == B200GP_OP_CONST) ? 0 : (dcode >> 1);
            ++depth;
            if (depth > 8) throw GpError("kernel program: expression too deep (max 8)");
        } else {
            throw GpError("kernel program: unknown opcode");
        }
    }
    if (depth != 1) throw GpError("kernel program: malformed expression");
    return P;
}


// inverse of parse_prog_impl: the program rows (metric definitions first, then the instructions with the metric index
// folded back into the distance code) -- used where a parsed program has to cross the C-ABI again
static inline std::vector<double> kprog_encode(const KProg& P) {
    std::vector<double> prog;
    for (int k = 0; k < P.nmetric; ++k) {
        const int r = P.mrows[k], c = P.mcols;
        const int nd_rows = (r * c + B200GP_PROG_STRIDE - 1) / B200GP_PROG_STRIDE;
        prog.push_back((double)B200GP_OP_METRIC); prog.push_back((double)(k + 1));
        prog.push_back((double)r); prog.push_back((double)c);
        std::vector<double> flat((size_t)nd_rows * B200GP_PROG_STRIDE, 0.0);
        for (int i = 0; i < r; ++i)
            for (int j = 0; j < c; ++j) flat[(size_t)i * c + j] = P.M[k][i * B200GP_METRIC_MAX_DIM + j];
        prog.insert(prog.end(), flat.begin(), flat.end());
    }
    for (int i = 0; i < P.n; ++i) {
        prog.push_back((double)P.op[i]); prog.push_back((double)(P.dist[i] + 2 * P.metric[i]));
        prog.push_back(P.p0[i]); prog.push_back(P.p1[i]);
    }
    return prog;
}
