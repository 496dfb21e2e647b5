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

// Envelope LDL' of a symmetric quasi-definite matrix stored densely (lower triangle), row i zero left of first[i].
struct ProfileLdl {
    int n = 0;
    Vec L, D;
    std::vector<int> first;
    bool Factor() {
        D.assign(n, 0.0);
        Vec tmp(n);
        for (int i = 0; i < n; i++) {
            double* Li = &L[static_cast<size_t>(i) * n];
            for (int j = first[i]; j < i; j++) {
                const double* Lj = &L[static_cast<size_t>(j) * n];
                double s = Li[j];
                for (int k = std::max(first[i], first[j]); k < j; k++) s -= tmp[k] * Lj[k];
                tmp[j] = s;            // = L_ij * D_j
            }
            double d = Li[i];
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
            const double* Li = &L[static_cast<size_t>(i) * n];
            double s = b[i];
            for (int k = first[i]; k < i; k++) s -= Li[k] * b[k];
            b[i] = s;
        }
        for (int i = 0; i < n; i++) b[i] /= D[i];
        for (int i = n - 1; i >= 0; i--) {
            const double* Li = &L[static_cast<size_t>(i) * n];
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
}  // namespace

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
    F.n = D;
    F.L.assign(static_cast<size_t>(D) * D, 0.0);
    F.first.assign(D, 0);
    auto at = [&](int a, int c) -> double& {
        const int pa = pos[a], pc = pos[c];
        const int i = std::max(pa, pc), j = std::min(pa, pc);
        F.first[i] = std::min(F.first[i], j);
        return F.L[static_cast<size_t>(i) * D + j];
    };
    Vec W(mi, 1.0);
    auto build_and_factor = [&]() -> bool {
        std::fill(F.L.begin(), F.L.end(), 0.0);
        for (int i = 0; i < D; i++) F.first[i] = i;
        for (int j = 0; j < n; j++)
            for (int k = P.colptr[j]; k < P.colptr[j + 1]; k++)
                if (P.rowidx[k] >= j) at(P.rowidx[k], j) += P.val[k];
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
    // solve the quasi-definite system for right-hand side (r1 ; r2) in natural unknown order
    Vec work(D);
    auto kkt_solve = [&](const Vec& r1, const Vec& r2, Vec& dz, Vec& dnu) {
        for (int j = 0; j < n; j++) work[pos[j]] = r1[j];
        for (int e = 0; e < me; e++) work[pos[n + e]] = r2[e];
        F.Solve(work.data());
        for (int j = 0; j < n; j++) dz[j] = work[pos[j]];
        for (int e = 0; e < me; e++) dnu[e] = work[pos[n + e]];
    };
    auto rowdot = [&](int i, const Vec& v) {
        double s = 0;
        for (int k = R.ptr[i]; k < R.ptr[i + 1]; k++) s += R.val[k] * v[R.col[k]];
        return s;
    };
    auto add_rowT = [&](int i, double coef, Vec& out) {
        for (int k = R.ptr[i]; k < R.ptr[i + 1]; k++) out[R.col[k]] += coef * R.val[k];
    };

    IpmResult res;
    Vec z(n, 0.0), s(mi), lam(mi), nu(me, 0.0), dz(n), dnu(me), ds(mi), dl(mi), rd(n), rp(mi), re(me), rc(mi), r1(n), r2(me), Pz(n);
    // ---- starting point: W = I
    if (!build_and_factor()) {
        res.status = Other;
        return res;
    }
    for (int j = 0; j < n; j++) r1[j] = -q[j];
    for (int r = 0; r < mi; r++) add_rowT(in_rows[r], b[in_rows[r]], r1);
    for (int e = 0; e < me; e++) r2[e] = b[eq_rows[e]];
    kkt_solve(r1, r2, z, nu);
    std::fill(nu.begin(), nu.end(), 0.0);
    double mn = 1e300;
    for (int r = 0; r < mi; r++) {
        s[r] = b[in_rows[r]] - rowdot(in_rows[r], z);
        mn = std::min(mn, s[r]);
    }
    const double shift = std::max(0.0, -1.5 * mn);
    double xi = 0, sl = 0, ss = 0;
    for (int r = 0; r < mi; r++) {
        const double v = std::max(s[r] + shift, 1e-2);
        s[r] = lam[r] = v;
        xi += v * v;
        sl += v;
    }
    for (int r = 0; r < mi; r++) {
        s[r] += 0.5 * xi / sl;
        ss += s[r];
    }
    for (int r = 0; r < mi; r++) lam[r] += 0.5 * xi / ss;

    const double nrm_q = std::max(1.0, NormInf(q));
    double nrm_b = 1.0;
    for (int r = 0; r < mi; r++) nrm_b = std::max(nrm_b, std::abs(b[in_rows[r]]));
    for (int e = 0; e < me; e++) nrm_b = std::max(nrm_b, std::abs(b[eq_rows[e]]));

    res.status = MaxIter;
    int it = 0;
    double n_rd = 0, n_rp = 0, n_re = 0, mu = 0, gscale = 1, last_rp = 0, last_re = 0, last_rd = 0, last_mu = 0, rp_ref = 0;
    for (it = 0; it <= st.max_iter; it++) {
        P.mul(z.data(), Pz.data());
        double pobj = 0;
        for (int j = 0; j < n; j++) {
            pobj += z[j] * (0.5 * Pz[j] + q[j]);
            rd[j] = Pz[j] + q[j];
        }
        for (int r = 0; r < mi; r++) add_rowT(in_rows[r], lam[r], rd);
        for (int e = 0; e < me; e++) add_rowT(eq_rows[e], nu[e], rd);
        double sdl = 0;
        for (int r = 0; r < mi; r++) {
            rp[r] = rowdot(in_rows[r], z) + s[r] - b[in_rows[r]];
            sdl += s[r] * lam[r];
        }
        for (int e = 0; e < me; e++) re[e] = rowdot(eq_rows[e], z) - b[eq_rows[e]];
        n_rd = NormInf(rd);
        n_rp = NormInf(rp);
        n_re = NormInf(re);
        mu = mi ? sdl / mi : 0.0;
        gscale = std::max(1.0, std::abs(pobj));
        if (n_rd != n_rd || n_rp != n_rp || mu != mu) {
            res.status = Other;
            n_rp = last_rp;
            n_re = last_re;
            n_rd = last_rd;
            mu = last_mu;
            break;
        }
        last_rp = n_rp;
        last_re = n_re;
        last_rd = n_rd;
        last_mu = mu;
        if (n_rd <= st.tol_feas * nrm_q && n_rp <= st.tol_feas * nrm_b && n_re <= st.tol_feas * nrm_b && sdl <= st.tol_gap * gscale) {
            res.status = Solved;
            break;
        }
        // Early exit on a stalled primal residual (stands in for the infeasibility certificate of Clarabel's homogeneous
        // embedding, which stops an infeasible QP long before the iteration limit): the primal residual shrinks by exactly
        // (1 - alpha) per step, so less than 10 % over ten iterations while it is still far from feasible means the
        // steps have collapsed.  Same rule as csrc/bgg_ipm.cu.
        {
            const double prim = std::max(n_rp, n_re);
            if (it == 0) rp_ref = prim;
            if (it > 0 && it % 10 == 0) {
                if (prim > 1e3 * st.tol_feas * nrm_b && prim >= 0.9 * rp_ref) break;   // classified below (PrimalInfeasible)
                rp_ref = prim;
            }
        }
        if (it == st.max_iter) break;
        for (int r = 0; r < mi; r++) W[r] = lam[r] / s[r];
        if (!build_and_factor()) {
            res.status = Other;
            break;
        }
        auto newton = [&](bool corrector, double sigmu) {
            for (int r = 0; r < mi; r++) {
                rc[r] = s[r] * lam[r];
                if (corrector) rc[r] += ds[r] * dl[r] - sigmu;
            }
            for (int j = 0; j < n; j++) r1[j] = -rd[j];
            for (int r = 0; r < mi; r++) add_rowT(in_rows[r], -(-rc[r] + lam[r] * rp[r]) / s[r], r1);
            for (int e = 0; e < me; e++) r2[e] = -re[e];
            kkt_solve(r1, r2, dz, dnu);
            for (int rf = 0; rf < st.refine; rf++) {   // refinement against the same regularised system
                Vec t1(n), t2(me), ez(n), enu(me);
                P.mul(dz.data(), t1.data());
                for (int r = 0; r < mi; r++) add_rowT(in_rows[r], W[r] * rowdot(in_rows[r], dz), t1);
                for (int e = 0; e < me; e++) {
                    add_rowT(eq_rows[e], dnu[e], t1);
                    t2[e] = rowdot(eq_rows[e], dz) - st.delta * dnu[e];
                }
                for (int j = 0; j < n; j++) t1[j] = r1[j] - t1[j];
                for (int e = 0; e < me; e++) t2[e] = r2[e] - t2[e];
                kkt_solve(t1, t2, ez, enu);
                for (int j = 0; j < n; j++) dz[j] += ez[j];
                for (int e = 0; e < me; e++) dnu[e] += enu[e];
            }
            for (int r = 0; r < mi; r++) {
                ds[r] = -rp[r] - rowdot(in_rows[r], dz);
                dl[r] = (-rc[r] - lam[r] * ds[r]) / s[r];
            }
        };
        auto max_step = [&]() {
            double a = 1e300;
            for (int r = 0; r < mi; r++) {
                if (ds[r] < 0) a = std::min(a, -s[r] / ds[r]);
                if (dl[r] < 0) a = std::min(a, -lam[r] / dl[r]);
            }
            return a;
        };
        newton(false, 0.0);
        const double a_aff = std::min(1.0, max_step());
        double mu_aff = 0;
        for (int r = 0; r < mi; r++) mu_aff += (s[r] + a_aff * ds[r]) * (lam[r] + a_aff * dl[r]);
        mu_aff = mi ? mu_aff / mi : 0.0;
        const double sr = (mu > 0) ? mu_aff / mu : 0.0;
        newton(true, sr * sr * sr * mu);
        const double alpha = std::min(1.0, 0.99 * max_step());
        for (int j = 0; j < n; j++) z[j] += alpha * dz[j];
        for (int r = 0; r < mi; r++) {
            s[r] += alpha * ds[r];
            lam[r] += alpha * dl[r];
        }
        for (int e = 0; e < me; e++) nu[e] += alpha * dnu[e];
    }
    // Exit classification without Clarabel's homogeneous embedding: residuals within 1e3 x tolerance -> SolvedInacc
    // ("AlmostSolved"); primal residual still far from feasible after the multipliers diverged -> PrimalInfeasible.
    if (res.status == MaxIter || res.status == Other) {
        const double loose = 1e3;
        if (n_rd <= loose * st.tol_feas * nrm_q && n_rp <= loose * st.tol_feas * nrm_b && n_re <= loose * st.tol_feas * nrm_b &&
            mu * mi <= loose * st.tol_gap * gscale)
            res.status = SolvedInacc;
        else if (n_rp > loose * st.tol_feas * nrm_b || n_re > loose * st.tol_feas * nrm_b)
            res.status = PrimalInfeasible;
    }
    res.iters = it;
    res.prim_res = std::max(n_rp, n_re);
    res.dual_res = n_rd;
    res.gap = mu * mi;
    res.x = z;
    res.y.assign(m, 0.0);
    res.s.assign(m, 0.0);
    for (int i = 0; i < m; i++)
        if (!is_eq[i]) res.s[i] = b[i] - rowdot(i, z);
    for (int r = 0; r < mi; r++) {
        res.y[in_rows[r]] = lam[r];
        res.s[in_rows[r]] = s[r];
    }
    for (int e = 0; e < me; e++) res.y[eq_rows[e]] = nu[e];
    return res;
}

QpSolution IpmQpSolver::Solve(const QpData& data, const Vec& /*warm_start*/, bool /*is_real_time*/) {
    const std::vector<char> eq = data.RowIsEquality();
    // elimination order: [x_k, multipliers of dynamics block k] per node, then the spline variables, then the
    // remaining equality multipliers -- keeps the envelope of the state chain 36 wide.
    const int n = data.num_vars, nblk = data.num_dynamics / 12;
    std::vector<int> order;
    int me = 0;
    for (char c : eq) me += c;
    for (int k = 0; k < nblk; k++) {
        for (int i = 0; i < 12; i++) order.push_back(12 * k + i);
        for (int i = 0; i < 12; i++) order.push_back(n + 12 * k + i);   // dynamics rows are the first equality rows
    }
    for (int j = 12 * nblk; j < n; j++) order.push_back(j);
    for (int e = data.num_dynamics; e < me; e++) order.push_back(n + e);
    const IpmResult r = IpmSolve(data.P, data.cost_linear, data.A, data.ub, eq, order, settings);
    QpSolution out;
    out.x = r.x;
    out.dual = r.y;
    out.slack = r.s;
    out.status = r.status;
    out.iters = r.iters;
    out.prim_res = r.prim_res;
    out.dual_res = r.dual_res;
    return out;
}

}  // namespace oracle
