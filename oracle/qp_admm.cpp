// TEST INFRASTRUCTURE ONLY -- CPU oracle (see qp_admm.hpp for scope and citations).
#include "qp_admm.hpp"

#include <algorithm>
#include <cmath>

namespace oracle {

namespace {
constexpr double kInfty = 1e30;        // OSQP_INFTY (and the reference's own "infinity", qp_interface.cpp:13-17)
constexpr double kMinScaling = 1e-4;   // OSQP MIN_SCALING
constexpr double kMaxScaling = 1e4;    // OSQP MAX_SCALING
constexpr double kRhoMin = 1e-6, kRhoMax = 1e6, kRhoTol = 1e-4, kRhoEqOverIneq = 1e3;

double NormInf(const Vec& v) {
    double m = 0;
    for (double x : v) m = std::max(m, std::abs(x));
    return m;
}
double LimitScaling(double v) {
    v = v < kMinScaling ? 1.0 : v;
    return v > kMaxScaling ? kMaxScaling : v;
}

struct Csr {
    int rows = 0, cols = 0;
    std::vector<int> rowptr, colidx;
    std::vector<int> src;   // index into the CSC value array
};
Csr ToCsr(const Csc& A) {
    Csr r;
    r.rows = A.rows;
    r.cols = A.cols;
    r.rowptr.assign(A.rows + 1, 0);
    for (int k : A.rowidx) r.rowptr[k + 1]++;
    for (int i = 0; i < A.rows; i++) r.rowptr[i + 1] += r.rowptr[i];
    r.colidx.resize(A.nnz());
    r.src.resize(A.nnz());
    std::vector<int> fill(r.rowptr.begin(), r.rowptr.end() - 1);
    for (int j = 0; j < A.cols; j++)
        for (int k = A.colptr[j]; k < A.colptr[j + 1]; k++) {
            const int pos = fill[A.rowidx[k]]++;
            r.colidx[pos] = j;
            r.src[pos] = k;
        }
    return r;
}

// Envelope (profile) Cholesky of a dense-stored SPD matrix whose row i is zero left of first[i].
struct ProfileChol {
    int n = 0;
    std::vector<double> L;       // row-major n x n, lower triangle
    std::vector<int> first;
    bool Factor() {
        for (int i = 0; i < n; i++) {
            double* Li = &L[static_cast<size_t>(i) * n];
            for (int j = first[i]; j <= i; j++) {
                const double* Lj = &L[static_cast<size_t>(j) * n];
                double s = Li[j];
                for (int k = std::max(first[i], first[j]); k < j; k++) s -= Li[k] * Lj[k];
                if (j < i) {
                    Li[j] = s / Lj[j];
                } else {
                    if (!(s > 0)) return false;
                    Li[i] = std::sqrt(s);
                }
            }
        }
        return true;
    }
    void Solve(double* b) const {
        for (int i = 0; i < n; i++) {
            const double* Li = &L[static_cast<size_t>(i) * n];
            double s = b[i];
            for (int k = first[i]; k < i; k++) s -= Li[k] * b[k];
            b[i] = s / Li[i];
        }
        for (int i = n - 1; i >= 0; i--) {
            const double* Li = &L[static_cast<size_t>(i) * n];
            const double xi = b[i] / Li[i];
            b[i] = xi;
            for (int k = first[i]; k < i; k++) b[k] -= Li[k] * xi;
        }
    }
};
}  // namespace

AdmmResult AdmmSolve(const Csc& P_in, const Vec& q_in, const Csc& A_in, const Vec& l_in, const Vec& u_in, const Vec& x0,
                     const Vec& y0, const AdmmSettings& st) {
    const int n = P_in.cols, m = A_in.rows;
    Csc P = P_in, A = A_in;
    Vec q = q_in, l = l_in, u = u_in;
    const Csr Ar = ToCsr(A);

    // ---- Ruiz equilibration + cost normalisation (osqp scaling.c: scale_data) ----
    Vec D(n, 1.0), E(m, 1.0);
    double c = 1.0;
    for (int it = 0; it < st.scaling; it++) {
        Vec Dt(n, 0.0), Et(m, 0.0);
        for (int j = 0; j < n; j++) {
            double mx = 0;
            for (int k = P.colptr[j]; k < P.colptr[j + 1]; k++) mx = std::max(mx, std::abs(P.val[k]));
            for (int k = A.colptr[j]; k < A.colptr[j + 1]; k++) {
                mx = std::max(mx, std::abs(A.val[k]));
                Et[A.rowidx[k]] = std::max(Et[A.rowidx[k]], std::abs(A.val[k]));
            }
            Dt[j] = mx;
        }
        for (double& v : Dt) v = 1.0 / std::sqrt(LimitScaling(v));
        for (double& v : Et) v = 1.0 / std::sqrt(LimitScaling(v));
        for (int j = 0; j < n; j++) {
            for (int k = P.colptr[j]; k < P.colptr[j + 1]; k++) P.val[k] *= Dt[P.rowidx[k]] * Dt[j];
            for (int k = A.colptr[j]; k < A.colptr[j + 1]; k++) A.val[k] *= Et[A.rowidx[k]] * Dt[j];
            D[j] *= Dt[j];
            q[j] *= Dt[j];
        }
        for (int i = 0; i < m; i++) E[i] *= Et[i];
        double mean = 0;
        for (int j = 0; j < n; j++) {
            double mx = 0;
            for (int k = P.colptr[j]; k < P.colptr[j + 1]; k++) mx = std::max(mx, std::abs(P.val[k]));
            mean += mx;
        }
        mean /= n;
        double ct = std::max(mean, LimitScaling(NormInf(q)));
        ct = 1.0 / LimitScaling(ct);
        for (double& v : P.val) v *= ct;
        for (double& v : q) v *= ct;
        c *= ct;
    }
    for (int i = 0; i < m; i++) {
        l[i] *= E[i];
        u[i] *= E[i];
    }

    // ---- rho vector (osqp auxil.c: set_rho_vec) ----
    std::vector<int> ctype(m);
    for (int i = 0; i < m; i++) {
        if (l[i] < -kInfty * kMinScaling && u[i] > kInfty * kMinScaling) ctype[i] = -1;
        else if (u[i] - l[i] < kRhoTol) ctype[i] = 1;
        else ctype[i] = 0;
    }
    double rho = std::min(std::max(st.rho, kRhoMin), kRhoMax);
    Vec rho_vec(m);
    auto set_rho_vec = [&]() {
        for (int i = 0; i < m; i++) rho_vec[i] = ctype[i] == -1 ? kRhoMin : (ctype[i] == 1 ? kRhoEqOverIneq * rho : rho);
    };
    set_rho_vec();

    // ---- reduced KKT: K = P + sigma I + A' diag(rho) A ----
    ProfileChol chol;
    chol.n = n;
    chol.L.assign(static_cast<size_t>(n) * n, 0.0);
    chol.first.assign(n, 0);
    auto build_and_factor = [&]() -> bool {
        std::fill(chol.L.begin(), chol.L.end(), 0.0);
        for (int i = 0; i < n; i++) chol.first[i] = i;
        for (int j = 0; j < n; j++)
            for (int k = P.colptr[j]; k < P.colptr[j + 1]; k++) {
                const int i = P.rowidx[k];
                if (i >= j) {
                    chol.L[static_cast<size_t>(i) * n + j] += P.val[k];
                    chol.first[i] = std::min(chol.first[i], j);
                }
            }
        for (int i = 0; i < n; i++) chol.L[static_cast<size_t>(i) * n + i] += st.sigma;
        for (int r = 0; r < m; r++) {
            for (int a = Ar.rowptr[r]; a < Ar.rowptr[r + 1]; a++) {
                const int ja = Ar.colidx[a];
                const double va = rho_vec[r] * A.val[Ar.src[a]];
                for (int b = Ar.rowptr[r]; b < Ar.rowptr[r + 1]; b++) {
                    const int jb = Ar.colidx[b];
                    if (jb > ja) continue;
                    chol.L[static_cast<size_t>(ja) * n + jb] += va * A.val[Ar.src[b]];
                    chol.first[ja] = std::min(chol.first[ja], jb);
                }
            }
        }
        return chol.Factor();
    };
    AdmmResult res;
    if (!build_and_factor()) {
        res.status = Other;
        return res;
    }

    // ---- warm start (osqp_warm_start: x scaled by Dinv, y by Einv*c, z = A x) ----
    Vec x(n), y(m), z(m);
    for (int j = 0; j < n; j++) x[j] = x0[j] / D[j];
    for (int i = 0; i < m; i++) y[i] = y0[i] / E[i] * c;
    A.mul(x.data(), z.data());

    Vec x_prev(n), z_prev(m), xt(n), zt(m), rhs(n), tmp_m(m), dy(m), Ax(m), Px(n), Aty(n);
    auto residuals = [&](double& pr, double& dr, double& eps_p, double& eps_d, double tol_mult) {
        A.mul(x.data(), Ax.data());
        P.mul(x.data(), Px.data());
        A.mul_t(y.data(), Aty.data());
        double nAx = 0, nz = 0;
        pr = 0;
        for (int i = 0; i < m; i++) {
            pr = std::max(pr, std::abs((Ax[i] - z[i]) / E[i]));
            nAx = std::max(nAx, std::abs(Ax[i] / E[i]));
            nz = std::max(nz, std::abs(z[i] / E[i]));
        }
        double nPx = 0, nAty = 0, nq = 0;
        dr = 0;
        for (int j = 0; j < n; j++) {
            dr = std::max(dr, std::abs((Px[j] + q[j] + Aty[j]) / D[j]));
            nPx = std::max(nPx, std::abs(Px[j] / D[j]));
            nAty = std::max(nAty, std::abs(Aty[j] / D[j]));
            nq = std::max(nq, std::abs(q[j] / D[j]));
        }
        dr /= c;
        eps_p = tol_mult * (st.eps_abs + st.eps_rel * std::max(nAx, nz));
        eps_d = tol_mult * (st.eps_abs + st.eps_rel * std::max({nPx, nAty, nq}) / c);
    };

    int iter = 0;
    res.status = MaxIter;
    for (iter = 1; iter <= st.max_iter; iter++) {
        x_prev = x;
        z_prev = z;
        // x~ : (P + sigma I + A' R A) x~ = sigma x_prev - q + A'(R z_prev - y)
        for (int i = 0; i < m; i++) tmp_m[i] = rho_vec[i] * z_prev[i] - y[i];
        A.mul_t(tmp_m.data(), rhs.data());
        for (int j = 0; j < n; j++) rhs[j] += st.sigma * x_prev[j] - q[j];
        chol.Solve(rhs.data());
        xt = rhs;
        A.mul(xt.data(), zt.data());
        for (int j = 0; j < n; j++) x[j] = st.alpha * xt[j] + (1 - st.alpha) * x_prev[j];
        for (int i = 0; i < m; i++) {
            const double zr = st.alpha * zt[i] + (1 - st.alpha) * z_prev[i];
            const double v = zr + y[i] / rho_vec[i];
            z[i] = std::min(std::max(v, l[i]), u[i]);
            dy[i] = rho_vec[i] * (zr - z[i]);
            y[i] += dy[i];
        }
        const bool check = (st.check_termination > 0 && iter % st.check_termination == 0) || iter == st.max_iter;
        if (check) {
            double pr, dr, ep, ed;
            residuals(pr, dr, ep, ed, 1.0);
            res.prim_res = pr;
            res.dual_res = dr;
            if (pr <= ep && dr <= ed) {
                res.status = Solved;
                break;
            }
            // primal infeasibility certificate (osqp auxil.c: is_primal_infeasible)
            Vec pdy(dy);
            for (int i = 0; i < m; i++) {
                const bool ub_inf = u[i] > kInfty * kMinScaling, lb_inf = l[i] < -kInfty * kMinScaling;
                if (ub_inf && lb_inf) pdy[i] = 0;
                else if (ub_inf) pdy[i] = std::min(pdy[i], 0.0);
                else if (lb_inf) pdy[i] = std::max(pdy[i], 0.0);
            }
            double ndy = 0;
            for (int i = 0; i < m; i++) ndy = std::max(ndy, std::abs(E[i] * pdy[i]));
            if (ndy > 1.0 / kInfty) {
                double lhs = 0;
                for (int i = 0; i < m; i++) lhs += u[i] * std::max(pdy[i], 0.0) + l[i] * std::min(pdy[i], 0.0);
                if (lhs < -st.eps_prim_inf * ndy) {
                    Vec Atdy(n);
                    A.mul_t(pdy.data(), Atdy.data());
                    double na = 0;
                    for (int j = 0; j < n; j++) na = std::max(na, std::abs(Atdy[j] / D[j]));
                    if (na < st.eps_prim_inf * ndy) {
                        res.status = PrimalInfeasible;
                        break;
                    }
                }
            }
        }
        if (st.adaptive_rho && st.adaptive_rho_interval > 0 && iter % st.adaptive_rho_interval == 0) {
            // osqp auxil.c: compute_rho_estimate / adapt_rho (scaled quantities)
            A.mul(x.data(), Ax.data());
            P.mul(x.data(), Px.data());
            A.mul_t(y.data(), Aty.data());
            double pr = 0, dr = 0;
            for (int i = 0; i < m; i++) pr = std::max(pr, std::abs(Ax[i] - z[i]));
            for (int j = 0; j < n; j++) dr = std::max(dr, std::abs(Px[j] + q[j] + Aty[j]));
            pr /= std::max(NormInf(z), NormInf(Ax)) + 1e-10;
            dr /= std::max({NormInf(q), NormInf(Aty), NormInf(Px)}) + 1e-10;
            double rho_new = rho * std::sqrt(pr / (dr + 1e-10));
            rho_new = std::min(std::max(rho_new, kRhoMin), kRhoMax);
            if (rho_new > rho * st.adaptive_rho_tolerance || rho_new < rho / st.adaptive_rho_tolerance) {
                rho = rho_new;
                set_rho_vec();
                res.rho_updates++;
                if (!build_and_factor()) {
                    res.status = Other;
                    break;
                }
            }
        }
    }
    res.iters = std::min(iter, st.max_iter);
    if (res.status == MaxIter) {   // OSQP's "solved inaccurate": 10x looser tolerances at max_iter
        double pr, dr, ep, ed;
        residuals(pr, dr, ep, ed, 10.0);
        if (pr <= ep && dr <= ed) res.status = SolvedInacc;
    }
    res.rho_final = rho;
    res.x.resize(n);
    res.y.resize(m);
    res.z.resize(m);
    for (int j = 0; j < n; j++) res.x[j] = D[j] * x[j];
    for (int i = 0; i < m; i++) {
        res.y[i] = E[i] * y[i] / c;
        res.z[i] = z[i] / E[i];
    }
    return res;
}

AdmmQpSolver::AdmmQpSolver() {
    // osqp_interface.cpp:261-266 (initial run) and :268-273 (real time); polishing is not restated.
    initial.eps_abs = 1e-5;
    initial.eps_rel = 1e-5;
    initial.max_iter = 3000;
    real_time.eps_abs = 1e-4;
    real_time.eps_rel = 1e-5;
    real_time.max_iter = 6000;
}

QpSolution AdmmQpSolver::Solve(const QpData& data, const Vec& warm_start, bool is_real_time) {
    const std::vector<char> eq = data.RowIsEquality();
    const int m = data.Total();
    Vec l(m), u(data.ub);
    for (int i = 0; i < m; i++) l[i] = eq[i] ? data.ub[i] : -kInfty;
    const Vec y0(m, 0.0);   // dual warm start is zeros, osqp_interface.cpp:53,74
    AdmmSettings s = is_real_time ? real_time : initial;
    const AdmmResult r = AdmmSolve(data.P, data.cost_linear, data.A, l, u, warm_start, y0, s);
    QpSolution out;
    out.x = r.x;
    out.dual = r.y;
    out.slack.resize(m);
    if (!r.x.empty()) {
        Vec Ax(m);
        data.A.mul(r.x.data(), Ax.data());
        for (int i = 0; i < m; i++) out.slack[i] = data.ub[i] - Ax[i];
    }
    out.status = r.status;
    out.iters = r.iters;
    out.prim_res = r.prim_res;
    out.dual_res = r.dual_res;
    return out;
}

}  // namespace oracle
