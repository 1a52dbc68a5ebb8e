// TEST INFRASTRUCTURE: runs the __host__ __device__ pieces of tinygp_b200/csrc/qs_core.cuh on the CPU, one "thread"
// at a time, so that `-m "not gpu"` tests can compare the SAME source the CUDA kernels execute (model lowering,
// per-point generators, the GramBack scan monoid, chunk / tree / replay decomposition) with the oracle.
// Built on demand by tests/test_device_code_on_host.py with nvcc; never linked into libb200gp.so.
#include "../../tinygp_b200/csrc/qs_core.cuh"
#include "../../tinygp_b200/csrc/qs_fast.cuh"

#include <vector>

namespace {

int g_tree_mode = 0;   // 0: fan-in-16 thread-sequential tree, 1: warp-shuffle scan (fan-in 32)

// the algorithm of warp_scan_kernel / warp_propagate_kernel / run_tree_warp (quasisep.cu), lane by lane: a "warp" is an
// array of 32 composites, __shfl_up becomes an indexed read of the previous step's array
template <class Op>
void host_tree_warp_top(const std::vector<double>& comp, int64_t n, std::vector<double>& start) {
    typedef typename StateOf<Op>::type State;
    const int64_t nw = (n + 31) / 32;
    std::vector<double> pre((size_t)Op::SIZE * n), totals((size_t)Op::SIZE * nw);
    for (int64_t w = 0; w < nw; ++w) {                               // warp_scan_kernel
        Op cur[32], nxt[32];
        for (int lane = 0; lane < 32; ++lane) {
            const int64_t i = w * 32 + lane;
            if (i < n) cur[lane].load(comp.data(), n, i);
            else cur[lane].identity();
        }
        for (int d = 1; d < 32; d <<= 1) {
            for (int lane = 0; lane < 32; ++lane) {
                if (lane >= d) { Op left = cur[lane - d]; left.combine(cur[lane]); nxt[lane] = left; }
                else nxt[lane] = cur[lane];
            }
            for (int lane = 0; lane < 32; ++lane) cur[lane] = nxt[lane];
        }
        for (int lane = 0; lane < 32; ++lane) {
            const int64_t i = w * 32 + lane;
            Op ex;
            if (lane == 0) ex.identity();
            else ex = cur[lane - 1];
            if (i < n) ex.store(pre.data(), n, i);
        }
        const int64_t last = ((n - w * 32) < 32 ? (n - w * 32) : 32) - 1;
        cur[last].store(totals.data(), nw, w);
    }
    std::vector<double> pstart;
    if (n > 32) {
        pstart.resize((size_t)Op::STATE * nw);
        host_tree_warp_top<Op>(totals, nw, pstart);
    }
    for (int64_t i = 0; i < n; ++i) {                                // warp_propagate_kernel
        State s;
        if (n > 32) state_load(s, pstart.data(), nw, i >> 5);
        else state_zero(s);
        Op e;
        e.load(pre.data(), n, i);
        e.apply(s);
        state_store(s, start.data(), n, i);
    }
}

// the structure of run_tree<Op> (quasisep.cu): up-sweep with fan-in TREE_R, top walk, down-sweep -- with the same
// Op::combine / Op::apply / state_* helpers the tree kernels call
template <class Op>
void host_tree(const std::vector<double>& comp, int64_t n, std::vector<double>& start) {
    typedef typename StateOf<Op>::type State;
    if (g_tree_mode == 1) {
        host_tree_warp_top<Op>(comp, n, start);
        return;
    }
    if (n <= TREE_R) {                                         // tree_top_kernel
        State s;
        state_zero(s);
        Op e;
        for (int64_t i = 0; i < n; ++i) {
            state_store(s, start.data(), n, i);
            e.load(comp.data(), n, i);
            e.apply(s);
        }
        return;
    }
    const int64_t np_ = (n + TREE_R - 1) / TREE_R;
    std::vector<double> parent((size_t)Op::SIZE * np_), pstart((size_t)Op::STATE * np_);
    for (int64_t i = 0; i < np_; ++i) {                        // tree_up_kernel
        Op acc, e;
        const int64_t b = i * TREE_R;
        acc.load(comp.data(), n, b);
        for (int64_t j = b + 1; j < b + TREE_R && j < n; ++j) { e.load(comp.data(), n, j); acc.combine(e); }
        acc.store(parent.data(), np_, i);
    }
    host_tree<Op>(parent, np_, pstart);
    for (int64_t i = 0; i < np_; ++i) {                        // tree_down_kernel
        State s;
        state_load(s, pstart.data(), np_, i);
        Op e;
        const int64_t b = i * TREE_R;
        for (int64_t j = b; j < b + TREE_R && j < n; ++j) {
            state_store(s, start.data(), n, j);
            if (j + 1 < b + TREE_R && j + 1 < n) { e.load(comp.data(), n, j); e.apply(s); }
        }
    }
}

template <int J>
void inv_diag_host(const QsModel& m, const double* t, const double* c, const double* w, int64_t n, double* out) {
    const int64_t nch = (n + m.chunk - 1) / m.chunk;
    std::vector<double> comp((size_t)GramBack<J>::SIZE * nch), start((size_t)J * J * nch);
    for (int64_t ch = 0; ch < nch; ++ch) gram_chunk_body<J>(m, t, c, w, n, comp.data(), nch, ch);
    host_tree<GramBack<J>>(comp, nch, start);
    for (int64_t ch = 0; ch < nch; ++ch) gram_replay_body<J>(m, t, c, w, n, start.data(), nch, out, ch);
}

// QuasisepSolver.__init__ (solver.py:73-82): chunk composites -> tree -> replay; optionally with the fused forward solve
template <int J>
void factor_host(const QsModel& m, const double* t, const double* diag, int64_t n, double* c, double* w,
                 double* logdet_half, int* info, const double* x_fuse, double* alpha) {
    const int64_t nch = (n + m.chunk - 1) / m.chunk;
    std::vector<double> comp((size_t)Riccati<J>::SIZE * nch), fstart((size_t)J * J * nch), part(nch);
    std::vector<double> acomp((size_t)Affine<J>::SIZE * nch), gstart((size_t)J * nch);
    for (int64_t ch = 0; ch < nch; ++ch) chol_chunk_body<J>(m, t, diag, n, comp.data(), nch, ch);
    host_tree<Riccati<J>>(comp, nch, fstart);
    *info = INT_MAX;
    for (int64_t ch = 0; ch < nch; ++ch)
        chol_replay_body<J>(m, t, diag, n, fstart.data(), nch, c, w, part.data(), info, x_fuse,
                            x_fuse ? acomp.data() : nullptr, ch);
    if (*info == INT_MAX) *info = 0;
    double s = 0.0;
    for (int64_t ch = 0; ch < nch; ++ch) s += part[ch];
    *logdet_half = s;
    if (x_fuse) {   // the fused path of qs_logp_impl: tree over the composites accumulated in the Cholesky replay
        host_tree<Affine<J>>(acomp, nch, gstart);
        for (int64_t ch = 0; ch < nch; ++ch)
            affine_replay_body<J, OP_LOWER_SOLVE>(m, t, diag, c, w, x_fuse, n, gstart.data(), nch, alpha, nullptr, ch);
    }
}

// the layout-specialised path of qs_fast.cu (qsf_run): chunk fold -> tree -> replay with the fused forward substitution and
// the per-chunk quadratic sums -> tree over the affine composites -> finish
template <int L>
void fast_factor_host(const QsModel& m, const double* t, const double* diag, int64_t n, double* c, double* w,
                      double* logdet_half, int* info, const double* x_fuse, double* sumsq) {
    constexpr int J = lay_J(L);
    const QsFastConst fc = qsf_constants(m);
    const int64_t nch = (n + m.chunk - 1) / m.chunk;
    std::vector<double> comp((size_t)Riccati<J>::SIZE * nch), fstart((size_t)J * J * nch), part(nch);
    std::vector<double> acomp((size_t)Affine<J>::SIZE * nch), gstart((size_t)J * nch), quad((size_t)QsfQuad<J>::SIZE * nch);
    for (int64_t ch = 0; ch < nch; ++ch) qsf_chunk_body<L>(m, fc, t, diag, n, comp.data(), nch, ch);
    host_tree<Riccati<J>>(comp, nch, fstart);
    *info = INT_MAX;
    for (int64_t ch = 0; ch < nch; ++ch)
        qsf_replay_body<L>(m, fc, t, diag, n, fstart.data(), nch, c, w, part.data(), info, x_fuse,
                           x_fuse ? acomp.data() : nullptr, x_fuse ? quad.data() : nullptr, ch);
    if (*info == INT_MAX) *info = 0;
    double s = 0.0;
    for (int64_t ch = 0; ch < nch; ++ch) s += part[ch];
    *logdet_half = s;
    if (x_fuse) {
        host_tree<Affine<J>>(acomp, nch, gstart);
        double ss = 0.0;
        for (int64_t ch = 0; ch < nch; ++ch) ss += qsf_quad_eval<J>(quad.data(), gstart.data(), nch, ch);
        *sumsq = ss;
    }
}

template <int L>
void fast_generators_host(const QsModel& m, const double* t, int64_t n, double* a_out, double* p_out) {
    constexpr int J = lay_J(L);
    const QsFastConst fc = qsf_constants(m);
    for (int64_t k = 0; k < n; ++k) {
        double a[J][J] = {}, p[J];
        qsf_gen<L>(m, fc, (k == 0) ? 0.0 : (t[k] - t[k - 1]), a, p);
        for (int i = 0; i < J; ++i) {
            p_out[k * J + i] = p[i];
            for (int j = 0; j < J; ++j) a_out[(k * J + i) * J + j] = a[i][j];
        }
    }
}

template <int J, int OP>
void affine_host(const QsModel& m, const double* t, const double* diag, const double* c, const double* w,
                 const double* x, int64_t n, double* out) {
    const int64_t nch = (n + m.chunk - 1) / m.chunk;
    std::vector<double> comp((size_t)Affine<J>::SIZE * nch), gstart((size_t)J * nch);
    for (int64_t ch = 0; ch < nch; ++ch) affine_chunk_body<J, OP>(m, t, c, w, x, n, comp.data(), nch, ch);
    host_tree<Affine<J>>(comp, nch, gstart);
    for (int64_t ch = 0; ch < nch; ++ch)
        affine_replay_body<J, OP>(m, t, diag, c, w, x, n, gstart.data(), nch, out, nullptr, ch);
}

template <int J>
void generators_host(const QsModel& m, const double* t, int64_t n, double* a_out, double* p_out) {
    for (int64_t k = 0; k < n; ++k) {
        double a[J][J], p[J];
        qs_gen<J>(m, (k == 0) ? 0.0 : (t[k] - t[k - 1]), a, p);
        for (int i = 0; i < J; ++i) {
            p_out[k * J + i] = p[i];
            for (int j = 0; j < J; ++j) a_out[(k * J + i) * J + j] = a[i][j];
        }
    }
}

}  // namespace

#define DISPATCH_J(Jv, CALL)                        \
    switch (Jv) {                                   \
        case 1: { constexpr int JJ = 1; CALL; } break; \
        case 2: { constexpr int JJ = 2; CALL; } break; \
        case 3: { constexpr int JJ = 3; CALL; } break; \
        case 4: { constexpr int JJ = 4; CALL; } break; \
        case 5: { constexpr int JJ = 5; CALL; } break; \
        case 6: { constexpr int JJ = 6; CALL; } break; \
        case 7: { constexpr int JJ = 7; CALL; } break; \
        case 8: { constexpr int JJ = 8; CALL; } break; \
        default: return 3;                          \
    }

extern "C" {

void hostcheck_set_tree(int mode) { g_tree_mode = mode; }

// model constants: J, q (h Pinf), h, d0
int hostcheck_model(const double* comps, int ncomp, int* J, double* q, double* h, double* d0) {
    try {
        QsModel m = build_model(comps, ncomp);
        *J = m.J;
        *d0 = m.d0;
        for (int i = 0; i < m.J; ++i) { q[i] = m.q[i]; h[i] = m.h[i]; }
        return 0;
    } catch (const std::exception&) { return 2; }
}

int hostcheck_generators(const double* comps, int ncomp, const double* t, int64_t n, double* a_out, double* p_out) {
    try {
        QsModel m = build_model(comps, ncomp);
        DISPATCH_J(m.J, (generators_host<JJ>(m, t, n, a_out, p_out)))
        return 0;
    } catch (const std::exception&) { return 2; }
}

int hostcheck_factor(const double* comps, int ncomp, const double* t, const double* diag, int64_t n, int chunk,
                     double* c, double* w, double* logdet_half, int* info, const double* x_fuse, double* alpha) {
    try {
        QsModel m = build_model(comps, ncomp);
        m.chunk = chunk;
        DISPATCH_J(m.J, (factor_host<JJ>(m, t, diag, n, c, w, logdet_half, info, x_fuse, alpha)))
        return 0;
    } catch (const std::exception&) { return 2; }
}

// op: the OP_* codes of qs_core.cuh (0 lower solve, 1 upper solve, 2 L z, 3 / 4 lower / upper part of K y)
int hostcheck_affine(const double* comps, int ncomp, int op, const double* t, const double* diag, const double* c,
                     const double* w, const double* x, int64_t n, int chunk, double* out) {
    try {
        QsModel m = build_model(comps, ncomp);
        m.chunk = chunk;
        switch (op) {
            case OP_LOWER_SOLVE: DISPATCH_J(m.J, (affine_host<JJ, OP_LOWER_SOLVE>(m, t, diag, c, w, x, n, out))) break;
            case OP_UPPER_SOLVE: DISPATCH_J(m.J, (affine_host<JJ, OP_UPPER_SOLVE>(m, t, diag, c, w, x, n, out))) break;
            case OP_LOWER_DOT: DISPATCH_J(m.J, (affine_host<JJ, OP_LOWER_DOT>(m, t, diag, c, w, x, n, out))) break;
            case OP_SYMM_LOWER: DISPATCH_J(m.J, (affine_host<JJ, OP_SYMM_LOWER>(m, t, diag, c, w, x, n, out))) break;
            case OP_SYMM_UPPER: DISPATCH_J(m.J, (affine_host<JJ, OP_SYMM_UPPER>(m, t, diag, c, w, x, n, out))) break;
            default: return 3;
        }
        return 0;
    } catch (const std::exception&) { return 2; }
}

// returns 4 if the model's block layout has no specialised code (the product then uses the generic path)
int hostcheck_fast_factor(const double* comps, int ncomp, const double* t, const double* diag, int64_t n, int chunk,
                          double* c, double* w, double* logdet_half, int* info, const double* x_fuse, double* sumsq) {
    try {
        QsModel m = build_model(comps, ncomp);
        m.chunk = chunk;
        switch (qsf_layout_of(m)) {
#define X(code) case code: fast_factor_host<code>(m, t, diag, n, c, w, logdet_half, info, x_fuse, sumsq); return 0;
            QSF_LAYOUTS(X)
#undef X
            default: return 4;
        }
    } catch (const std::exception&) { return 2; }
}

int hostcheck_fast_generators(const double* comps, int ncomp, const double* t, int64_t n, double* a_out, double* p_out) {
    try {
        QsModel m = build_model(comps, ncomp);
        switch (qsf_layout_of(m)) {
#define X(code) case code: fast_generators_host<code>(m, t, n, a_out, p_out); return 0;
            QSF_LAYOUTS(X)
#undef X
            default: return 4;
        }
    } catch (const std::exception&) { return 2; }
}

int hostcheck_inverse_diagonal(const double* comps, int ncomp, const double* t, int64_t n, const double* c,
                               const double* w, int chunk, double* out) {
    try {
        QsModel m = build_model(comps, ncomp);
        m.chunk = chunk;
        DISPATCH_J(m.J, (inv_diag_host<JJ>(m, t, c, w, n, out)))
        return 0;
    } catch (const std::exception&) { return 2; }
}

}  // extern "C"
