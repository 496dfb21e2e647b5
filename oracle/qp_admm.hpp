// TEST INFRASTRUCTURE ONLY -- CPU oracle.  Nothing under oracle/ may be imported, linked or executed by the
// product path; see foot_spline.hpp.
//
// CPU restatement of the OSQP ADMM algorithm (Stellato et al., "OSQP: an operator splitting solver for quadratic
// programs", Math. Prog. Comp. 2020; osqp v0.6/1.0 sources: scaling.c, auxil.c, osqp.c), the solver behind the
// reference's OSQPInterface (mpc/qp/osqp_interface.cpp:16-31 settings, :261-273 initial/real-time overrides,
// :49-74 rebuild + warm start with zero duals).  OSQP itself is an unpinned third-party dependency that is not in
// /root/reference and not installed here, so this follows the published algorithm; it is pinned on the reference's
// own 3-variable cross-solver QP (test/mpc_test.cpp:857-904, tolerance 1e-4 at :951-958) in
// tests/test_oracle_qp.py.  Differences from stock OSQP, all deliberate and deterministic:
//   * the KKT step solves the reduced system (P + sigma I + A^T diag(rho) A) x = rhs by an envelope Cholesky
//     (OSQP uses QDLDL on the quasi-definite KKT; same iterates up to round-off);
//   * rho adaptation happens at a fixed iteration interval (OSQP's default interval is timing based);
//   * no polishing.
#pragma once
#include "srb_mpc.hpp"

namespace oracle {

struct AdmmSettings {
    double rho = 1e-3, sigma = 1e-6, alpha = 1.6;           // osqp_interface.cpp:24-26
    double eps_abs = 1e-4, eps_rel = 1e-4;                  // :20-21
    double eps_prim_inf = 1e-4, eps_dual_inf = 1e-4;        // :18-19
    int max_iter = 1000;                                    // :23
    int scaling = 100;                                      // :28
    int check_termination = 25;                             // OSQP default
    bool adaptive_rho = true;                               // OSQP default
    int adaptive_rho_interval = 50;                         // fixed (see header comment)
    double adaptive_rho_tolerance = 5.0;                    // OSQP default
};

struct AdmmResult {
    Vec x, y, z;          // OSQP conventions: l <= A x <= u, y > 0 on an active upper bound
    SolveQuality status = Unsolved;
    int iters = 0, rho_updates = 0;
    double prim_res = 0, dual_res = 0, rho_final = 0;
};

// min 1/2 x'Px + q'x  s.t. l <= Ax <= u.  P holds both triangles (it is symmetric as stored by the reference).
AdmmResult AdmmSolve(const Csc& P, const Vec& q, const Csc& A, const Vec& l, const Vec& u, const Vec& x0, const Vec& y0,
                     const AdmmSettings& s);

// The QpSolver seam on top of it: Clarabel-form QpData in, Clarabel-convention solution out.
class AdmmQpSolver : public QpSolver {
public:
    AdmmSettings initial, real_time;
    AdmmQpSolver();
    QpSolution Solve(const QpData& data, const Vec& warm_start, bool is_real_time) override;
};

}  // namespace oracle
