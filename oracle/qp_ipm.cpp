// TEST INFRASTRUCTURE ONLY -- CPU oracle (see qp_ipm.hpp for scope and citations).
#include "qp_ipm.hpp"

#include <algorithm>
#include <cmath>
#include <numeric>

namespace oracle {

namespace {
struct RowView {   // CSR of A
    std::vector<int> ptr, col;
    Vec val;
};
RowView ToRows(const Csc& A) {
    RowView r;
    r.ptr.assign(A.rows + 1, 0);
    for (int k : A.rowidx) r.ptr[k + 1]++;
    for (int i = 0; i < A.rows; i++) r.ptr[i + 1] += r.ptr[i];
    r.col.resize(A.nnz());
    r.val.resize(A.nnz());
    std::vector<int> fill(r.ptr.begin(), r.ptr.end() - 1);
    for (int j = 0; j < A.cols; j++)
        for (int k = A.colptr[j]; k < A.colptr[j + 1]; k++) {
            const int p = fill[A.rowidx[k]]++;
            r.col[p] = j;
            r.val[p] = A.val[k];
        }
    return r;
}

// dot product with four independent partial sums (fixed association order; lets the compiler keep the loop in SIMD
// registers without -ffast-math); AVX2 clone selected at load time where the host has it
__attribute__((target_clones("avx2", "default"))) double Dot4(const double* a, const double* b, int n) {
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    int k = 0;
    for (; k + 4 <= n; k += 4) {
        s0 += a[k] * b[k];
        s1 += a[k + 1] * b[k + 1];
        s2 += a[k + 2] * b[k + 2];
        s3 += a[k + 3] * b[k + 3];
    }
    for (; k < n; k++) s0 += a[k] * b[k];
    return (s0 + s1) + (s2 + s3);
}

// Envelope LDL' of a symmetric quasi-definite matrix: row i holds columns first[i] .. i-1 (ragged storage, so that
// re-assembly and factorisation touch the envelope only) and the diagonal separately.
struct ProfileLdl {
    int n = 0;
    std::vector<int> first;
    std::vector<size_t> off;
    Vec L, D, diag;
    void SetPattern(const std::vector<int>& f) {
        n = static_cast<int>(f.size());
        first = f;
        off.assign(n + 1, 0);
        for (int i = 0; i < n; i++) off[i + 1] = off[i] + static_cast<size_t>(i - first[i]);
        L.assign(off[n], 0.0);
        D.assign(n, 0.0);
        diag.assign(n, 0.0);
    }
    void Clear() {
        std::fill(L.begin(), L.end(), 0.0);
        std::fill(diag.begin(), diag.end(), 0.0);
    }
    double& At(int i, int j) { return (i == j) ? diag[i] : L[off[i] + (j - first[i])]; }   // i >= j >= first[i]
    bool Factor() {
        Vec tmp(n);
        for (int i = 0; i < n; i++) {
            double* Li = &L[off[i]] - first[i];
            for (int j = first[i]; j < i; j++) {
                const double* Lj = &L[off[j]] - first[j];
                const int k0 = std::max(first[i], first[j]);
                tmp[j] = Li[j] - Dot4(&tmp[k0], Lj + k0, j - k0);            // = L_ij * D_j
            }
            double d = diag[i];
            for (int j = first[i]; j < i; j++) {
                const double lij = tmp[j] / D[j];
                d -= lij * tmp[j];
                Li[j] = lij;
            }
            if (d == 0.0 || d != d) return false;
            D[i] = d;
        }
        return true;
    }
    void Solve(double* b) const {
        for (int i = 0; i < n; i++) {
            const double* Li = &L[off[i]] - first[i];
            b[i] -= Dot4(Li + first[i], b + first[i], i - first[i]);
        }
        for (int i = 0; i < n; i++) b[i] /= D[i];
        for (int i = n - 1; i >= 0; i--) {
            const double* Li = &L[off[i]] - first[i];
            const double xi = b[i];
            for (int k = first[i]; k < i; k++) b[k] -= Li[k] * xi;
        }
    }
};
double NormInf(const Vec& v) {
    double m = 0;
    for (double x : v) m = std::max(m, std::abs(x));
    return m;
}
double Dot(const Vec& a, const Vec& b) {
    double s = 0;
    for (size_t i = 0; i < a.size(); i++) s += a[i] * b[i];
    return s;
}
}  // namespace

// Homogeneous self-dual embedding, as Clarabel (Goulart & Chen 2024, sections 2-3) restricted to Zero / Nonnegative cones:
//   P x + A'y + q tau = 0 ,  A x + s - b tau = 0 ,  kappa + q'x + b'y + x'Px / tau = 0 ,  s o z = mu ,  tau kappa = mu
// (y = all multipliers, z = its Nonnegative part).  Per iteration one factorisation of the reduced quasi-definite matrix
//   [ P + eps I + A_I' W A_I , A_E' ; A_E , -delta I ] ,   W = 1 / (s/z + eps)
// (eps, delta: Clarabel's static regularisation, here folded into the eliminated cone block), three solves with it -- the
// constant right-hand side (-q ; b), the affine and the combined step -- sigma = (1 - alpha_aff)^3, step fraction 0.99.  Termination on the
// de-homogenised point; a primal infeasibility certificate is b'y < 0 with A'y ~ 0 (Clarabel's is_primal_infeasible).
// The QP here is strictly convex (mpc.cpp:1090-1095 adds 1e-3 I), so the dual-infeasible branch is not restated.
// csrc/bgg_ipm.cu runs the same iteration on the condensed QP.
IpmResult IpmSolve(const Csc& P, const Vec& q, const Csc& A, const Vec& b, const std::vector<char>& is_eq,
                   const std::vector<int>& order, const IpmSettings& st) {
    const int n = P.cols, m = A.rows;
    const RowView R = ToRows(A);
    std::vector<int> eq_rows, in_rows;
    for (int i = 0; i < m; i++) {
        if (is_eq[i]) {
            eq_rows.push_back(i);
        } else {
            bool nz = false;   // rows that are identically zero (0 <= b) carry no interior: keep them out
            for (int k = R.ptr[i]; k < R.ptr[i + 1]; k++) nz |= (R.val[k] != 0.0);
            if (nz) in_rows.push_back(i);
        }
    }
    const int me = static_cast<int>(eq_rows.size()), mi = static_cast<int>(in_rows.size());
    const int D = n + me;
    std::vector<int> pos(D);   // unknown -> position in the elimination order
    if (order.empty()) {
        std::iota(pos.begin(), pos.end(), 0);
    } else {
        for (int k = 0; k < D; k++) pos[order[k]] = k;
    }
    ProfileLdl F;
    {   // symbolic pass: envelope of the permuted matrix
        std::vector<int> first(D);
        std::iota(first.begin(), first.end(), 0);
        auto touch = [&](int a, int c) {
            const int pa = pos[a], pc = pos[c];
            const int i = std::max(pa, pc), j = std::min(pa, pc);
            first[i] = std::min(first[i], j);
        };
        for (int j = 0; j < n; j++)
            for (int k = P.colptr[j]; k < P.colptr[j + 1]; k++) touch(P.rowidx[k], j);
        for (int r = 0; r < mi; r++) {
            const int i = in_rows[r];
            for (int a = R.ptr[i]; a < R.ptr[i + 1]; a++)
                for (int c = R.ptr[i]; c <= a; c++) touch(R.col[a], R.col[c]);
        }
        for (int e = 0; e < me; e++) {
            const int i = eq_rows[e];
            for (int a = R.ptr[i]; a < R.ptr[i + 1]; a++) touch(n + e, R.col[a]);
        }
        F.SetPattern(first);
    }
    auto at = [&](int a, int c) -> double& {
        const int pa = pos[a], pc = pos[c];
        return F.At(std::max(pa, pc), std::min(pa, pc));
    };
    Vec W(mi, 1.0), Dg(mi, 1.0);
    auto build_and_factor = [&]() -> bool {
        F.Clear();
        for (int j = 0; j < n; j++) {
            for (int k = P.colptr[j]; k < P.colptr[j + 1]; k++)
                if (P.rowidx[k] >= j) at(P.rowidx[k], j) += P.val[k];
            at(j, j) += st.eps;
        }
        for (int r = 0; r < mi; r++) {
            const int i = in_rows[r];
            for (int a = R.ptr[i]; a < R.ptr[i + 1]; a++)
                for (int c = R.ptr[i]; c <= a; c++) at(R.col[a], R.col[c]) += W[r] * R.val[a] * R.val[c];
        }
        for (int e = 0; e < me; e++) {
            const int i = eq_rows[e];
            for (int a = R.ptr[i]; a < R.ptr[i + 1]; a++) at(n + e, R.col[a]) += R.val[a];
            at(n + e, n + e) += -st.delta;
        }
        return F.Factor();
    };
    auto rowdot = [&](int i, const Vec& v) {
        double s = 0;
        for (int k = R.ptr[i]; k < R.ptr[i + 1]; k++) s += R.val[k] * v[R.col[k]];
        return s;
    };
    auto add_rowT = [&](int i, double coef, Vec& out) {
        for (int k = R.ptr[i]; k < R.ptr[i + 1]; k++) out[R.col[k]] += coef * R.val[k];
    };
    // Solve the REGULARISED system   (P + eps I) dx + A_I'dz + A_E'dy = a1 ,  A_I dx - (Dg + eps) dz = a2 ,  A_E dx - delta dy = a3
    // (Dg = s / z).  The static regularisation is not refined away: with the step taken from this system the method is the
    // primal-dual proximal (exactly regularised) interior-point iteration, whose fixed point is the solution of the
    // unregularised QP; refining a nearly degenerate vertex against the unregularised system with a contraction factor
    // close to 1 stalls instead (tools/ipm_proto.py, instance 2344 of the config #2 batch).  `refine` steps of iterative
    // refinement against the same regularised matrix, applied matrix-free, remove the rounding of the factorisation.
    Vec work(D), t1(n), cx(n), cy(me), e1(n), e3(me), tz(mi);
    auto ldl_solve = [&](const Vec& r1, const Vec& r3, Vec& dx, Vec& dy) {
        for (int j = 0; j < n; j++) work[pos[j]] = r1[j];
        for (int e = 0; e < me; e++) work[pos[n + e]] = r3[e];
        F.Solve(work.data());
        for (int j = 0; j < n; j++) dx[j] = work[pos[j]];
        for (int e = 0; e < me; e++) dy[e] = work[pos[n + e]];
    };
    auto kkt_solve = [&](const Vec& a1, const Vec& a2, const Vec& a3, Vec& dx, Vec& dz, Vec& dy) {
        t1 = a1;
        for (int r = 0; r < mi; r++) add_rowT(in_rows[r], W[r] * a2[r], t1);
        ldl_solve(t1, a3, dx, dy);
        for (int rf = 0; rf < st.refine; rf++) {
            P.mul(dx.data(), e1.data());
            for (int j = 0; j < n; j++) e1[j] += st.eps * dx[j];
            for (int r = 0; r < mi; r++) add_rowT(in_rows[r], W[r] * rowdot(in_rows[r], dx), e1);
            for (int e = 0; e < me; e++) {
                add_rowT(eq_rows[e], dy[e], e1);
                e3[e] = a3[e] - (rowdot(eq_rows[e], dx) - st.delta * dy[e]);
            }
            for (int j = 0; j < n; j++) e1[j] = t1[j] - e1[j];
            ldl_solve(e1, e3, cx, cy);
            for (int j = 0; j < n; j++) dx[j] += cx[j];
            for (int e = 0; e < me; e++) dy[e] += cy[e];
        }
        for (int r = 0; r < mi; r++) dz[r] = W[r] * (rowdot(in_rows[r], dx) - a2[r]);
    };

    IpmResult res;
    Vec x(n, 0.0), s(mi), z(mi), y(me, 0.0), Px(n), rx(n), rz(mi), re(me);
    Vec x1(n), z1(mi), y1(me), x2(n), z2(mi), y2(me), dx(n), dz(mi), dy(me), ds(mi), a1(n), a2(mi), a3(me);
    Vec bi(mi), be(me), mq(n);
    for (int r = 0; r < mi; r++) bi[r] = b[in_rows[r]];
    for (int e = 0; e < me; e++) be[e] = b[eq_rows[e]];
    for (int j = 0; j < n; j++) mq[j] = -q[j];
    // ---- starting point (Clarabel's QP initialisation): unit scaling, s = -z, both shifted into the cone
    if (!build_and_factor()) {
        res.status = Other;
        res.no_iterate = true;
        res.x.assign(n, 0.0);
        res.y.assign(m, 0.0);
        res.s.assign(m, 0.0);
        return res;
    }
    kkt_solve(mq, bi, be, x, z, y);
    auto shift = [&](Vec& v) {
        double mn = 1e300;
        for (double e : v) mn = std::min(mn, e);
        if (mn < 1e-8)
            for (double& e : v) e += 1.0 - mn;
    };
    for (int r = 0; r < mi; r++) s[r] = -z[r];
    shift(s);
    shift(z);
    double tau = 1.0, kap = 1.0;

    const double nrm_q = std::max(1.0, NormInf(q));
    const double nrm_b = std::max(1.0, NormInf(b));
    res.status = MaxIter;
    int it = 0;
    double res_p = 0, res_d = 0, gap = 0, gscale = 1, bz = 0, aty = 0, ynorm = 1;
    bool have_point = false;
    Vec xg, zg, yg, sg;
    double taug = 1.0;
    for (it = 0; it <= st.max_iter; it++) {
        P.mul(x.data(), Px.data());
        const double xPx = Dot(x, Px);
        for (int j = 0; j < n; j++) rx[j] = Px[j] + q[j] * tau;
        Vec aty_v(n, 0.0);
        for (int r = 0; r < mi; r++) add_rowT(in_rows[r], z[r], aty_v);
        for (int e = 0; e < me; e++) add_rowT(eq_rows[e], y[e], aty_v);
        for (int j = 0; j < n; j++) rx[j] += aty_v[j];
        for (int r = 0; r < mi; r++) rz[r] = rowdot(in_rows[r], x) + s[r] - bi[r] * tau;
        for (int e = 0; e < me; e++) re[e] = rowdot(eq_rows[e], x) - be[e] * tau;
        const double qx = Dot(q, x);
        bz = Dot(bi, z) + Dot(be, y);
        const double rt = kap + qx + bz + xPx / tau;
        const double mu = (Dot(s, z) + tau * kap) / (mi + 1);
        const double pc = (0.5 * xPx / tau + qx) / tau, dc = (-bz - 0.5 * xPx / tau) / tau;
        res_p = std::max(NormInf(rz), NormInf(re)) / tau;
        res_d = NormInf(rx) / tau;
        gap = std::abs(pc - dc);
        gscale = std::max(1.0, std::min(std::abs(pc), std::abs(dc)));
        aty = NormInf(aty_v);
        ynorm = std::max(1.0, std::max(NormInf(z), NormInf(y)));
        if (!(res_p == res_p) || !(res_d == res_d) || !(mu == mu) || !(tau > 0)) {
            res.status = Other;
            break;
        }
        have_point = true;   // last iterate with finite residuals: what is returned
        xg = x; zg = z; yg = y; sg = s; taug = tau;
        if (res_d <= st.tol_feas * nrm_q && res_p <= st.tol_feas * nrm_b && gap <= st.tol_gap * gscale) {
            res.status = Solved;
            break;
        }
        if (bz < -st.tol_infeas && aty <= st.tol_infeas * ynorm * (-bz)) {
            res.status = PrimalInfeasible;
            break;
        }
        if (it == st.max_iter) break;
        for (int r = 0; r < mi; r++) {
            Dg[r] = s[r] / z[r];
            W[r] = 1.0 / (Dg[r] + st.eps);
        }
        if (!build_and_factor()) {
            res.status = Other;
            break;
        }
        kkt_solve(mq, bi, be, x1, z1, y1);
        const double den = kap / tau - Dot(q, x1) - Dot(bi, z1) - Dot(be, y1) + xPx / (tau * tau) - 2.0 * Dot(Px, x1) / tau;
        double dtau = 0, dkap = 0;
        // step for right-hand sides (d_x, d_z, d_e, d_tau, d_kappa, d_s): see the derivation in DESIGN.md section 3
        auto step = [&](double scale, const Vec& d_s, double d_kap) {
            for (int j = 0; j < n; j++) a1[j] = -scale * rx[j];
            for (int r = 0; r < mi; r++) a2[r] = -scale * rz[r] + d_s[r] / z[r];
            for (int e = 0; e < me; e++) a3[e] = -scale * re[e];
            kkt_solve(a1, a2, a3, x2, z2, y2);
            double num = scale * rt - d_kap / tau + Dot(q, x2) + Dot(bi, z2) + Dot(be, y2) + 2.0 * Dot(Px, x2) / tau;
            dtau = num / den;
            for (int j = 0; j < n; j++) dx[j] = x2[j] + dtau * x1[j];
            for (int r = 0; r < mi; r++) {
                dz[r] = z2[r] + dtau * z1[r];
                ds[r] = (-d_s[r] - s[r] * dz[r]) / z[r];
            }
            for (int e = 0; e < me; e++) dy[e] = y2[e] + dtau * y1[e];
            dkap = (-d_kap - kap * dtau) / tau;
        };
        auto max_step = [&]() {
            double a = 1.0;
            for (int r = 0; r < mi; r++) {
                if (ds[r] < 0) a = std::min(a, -s[r] / ds[r]);
                if (dz[r] < 0) a = std::min(a, -z[r] / dz[r]);
            }
            if (dtau < 0) a = std::min(a, -tau / dtau);
            if (dkap < 0) a = std::min(a, -kap / dkap);
            return a;
        };
        Vec d_s(mi);
        for (int r = 0; r < mi; r++) d_s[r] = s[r] * z[r];
        step(1.0, d_s, kap * tau);
        const double a_aff = max_step();
        const double sigma = (1.0 - a_aff) * (1.0 - a_aff) * (1.0 - a_aff);
        for (int r = 0; r < mi; r++) d_s[r] = s[r] * z[r] + ds[r] * dz[r] - sigma * mu;
        step(1.0 - sigma, d_s, kap * tau + dkap * dtau - sigma * mu);
        const double alpha = 0.99 * max_step();
        {   // a non-finite direction is not applied (same guard as the kernel): the last iterate is kept
            bool fin = std::isfinite(dtau) && std::isfinite(dkap) && std::isfinite(alpha);
            for (int j = 0; j < n && fin; j++) fin = std::isfinite(dx[j]);
            for (int r = 0; r < mi && fin; r++) fin = std::isfinite(dz[r]) && std::isfinite(ds[r]);
            if (!fin) {
                res.status = Other;
                break;
            }
        }
        for (int j = 0; j < n; j++) x[j] += alpha * dx[j];
        for (int r = 0; r < mi; r++) {
            s[r] += alpha * ds[r];
            z[r] += alpha * dz[r];
        }
        for (int e = 0; e < me; e++) y[e] += alpha * dy[e];
        tau += alpha * dtau;
        kap += alpha * dkap;
    }
    // Exit without meeting the tolerances (iteration limit or a numerical breakdown): Clarabel's reduced tolerances decide
    // between AlmostSolved, AlmostPrimalInfeasible and the plain failure status (reduced_tol_feas 1e-4, reduced_tol_gap 5e-5,
    // reduced_tol_infeas 5e-5).  A breakdown before the first complete residual evaluation stays `Other` with no iterate.
    if ((res.status == MaxIter || res.status == Other) && have_point) {
        if (res_d <= 1e-4 * nrm_q && res_p <= 1e-4 * nrm_b && gap <= 5e-5 * gscale)
            res.status = SolvedInacc;
        else if (bz < -5e-5 && aty <= 5e-5 * ynorm * (-bz))
            res.status = PrimalInfeasibleInacc;
    }
    if (it > st.max_iter) it = st.max_iter;
    res.iters = it;
    res.prim_res = res_p;
    res.dual_res = res_d;
    res.gap = gap;
    if (!have_point) {
        res.no_iterate = true;
        res.x.assign(n, 0.0);
        res.y.assign(m, 0.0);
        res.s.assign(m, 0.0);
        return res;
    }
    res.x.assign(n, 0.0);
    for (int j = 0; j < n; j++) res.x[j] = xg[j] / taug;
    res.y.assign(m, 0.0);
    res.s.assign(m, 0.0);
    for (int i = 0; i < m; i++)
        if (!is_eq[i]) res.s[i] = b[i] - rowdot(i, res.x);
    for (int r = 0; r < mi; r++) {
        res.y[in_rows[r]] = zg[r] / taug;
        res.s[in_rows[r]] = sg[r] / taug;
    }
    for (int e = 0; e < me; e++) res.y[eq_rows[e]] = yg[e] / taug;
    return res;
}

// The MPC's elimination order -- [x_k, multipliers of dynamics block k] per node, then the spline variables, then the
// remaining equality multipliers: keeps the envelope of the state chain 36 wide -- shared by the oracle's QpSolver and by
// the stand-in behind the reference's ClarabelInterface (ref_shim/ref_mpc_capi.cpp), so both run the same arithmetic.
IpmResult IpmSolveMpcOrder(const Csc& P, const Vec& q, const Csc& A, const Vec& b, const std::vector<char>& is_eq, int num_dynamics,
                           const IpmSettings& st) {
    const int n = P.cols, nblk = num_dynamics / 12;
    std::vector<int> order;
    int me = 0;
    for (char c : is_eq) me += c;
    for (int k = 0; k < nblk; k++) {
        for (int i = 0; i < 12; i++) order.push_back(12 * k + i);
        for (int i = 0; i < 12; i++) order.push_back(n + 12 * k + i);   // dynamics rows are the first equality rows
    }
    for (int j = 12 * nblk; j < n; j++) order.push_back(j);
    for (int e = num_dynamics; e < me; e++) order.push_back(n + e);
    return IpmSolve(P, q, A, b, is_eq, order, st);
}

QpSolution IpmQpSolver::Solve(const QpData& data, const Vec& /*warm_start*/, bool /*is_real_time*/) {
    const std::vector<char> eq = data.RowIsEquality();
    const IpmResult r = IpmSolveMpcOrder(data.P, data.cost_linear, data.A, data.ub, eq, data.num_dynamics, settings);
    QpSolution out;
    out.x = r.x;
    out.dual = r.y;
    out.slack = r.s;
    out.status = r.status;
    out.iters = r.iters;
    out.prim_res = r.prim_res;
    out.dual_res = r.dual_res;
    out.no_iterate = r.no_iterate;
    return out;
}

}  // namespace oracle
