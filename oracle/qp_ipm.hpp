// TEST INFRASTRUCTURE ONLY -- CPU oracle.  Nothing under oracle/ may be imported, linked or executed by the
// product path; see foot_spline.hpp.
//
// CPU interior-point solver for the reference's QP in Clarabel form (min 1/2 z'Pz + q'z, Az + s = b, s in Zero x
// Nonnegative cones; mpc/qp/clarabel_interface.cpp:29-70,72-155).  Clarabel (Rust, v unpinned, absent here) is a
// primal-dual interior-point method with Mehrotra predictor-corrector steps, static KKT regularisation and
// iterative refinement (Goulart & Chen, "Clarabel: an interior-point solver for conic programs with quadratic
// objectives", 2024); restricted to Zero/Nonnegative cones that is the textbook Mehrotra method restated here.
// The homogeneous embedding Clarabel adds for infeasibility certificates is not restated: infeasibility is reported
// from the residual at the iteration limit.  Pinned on the reference's own 3-variable cross-solver QP
// (test/mpc_test.cpp:857-904; primal within 1e-4, :951-953) and checked against KKT conditions on MPC-size QPs.
// Linear algebra: the quasi-definite system [P + A_I'WA_I, A_E'; A_E, -delta I] is factorised by an envelope LDL'
// on an ordering that interleaves each node's states with its dynamics multipliers (no Eigen/QDLDL here).
#pragma once
#include "srb_mpc.hpp"

namespace oracle {

struct IpmSettings {
    double tol_feas = 1e-8, tol_gap = 1e-8;   // Clarabel defaults; the reference tightens feas to 1e-10 (:18-27)
    double tol_infeas = 1e-8;                 // Clarabel's tol_infeas_abs / tol_infeas_rel
    double eps = 1e-10;                       // static regularisation of the (1,1) block and of the eliminated cone block
    double delta = 1e-10;                     // static regularisation of the equality block
    int max_iter = 50;                        // (Clarabel's default is 200; a QP of this family that needs more than 50 is reported MaxIter)
    int refine = 1;                           // iterative-refinement steps per solve, against the same regularised matrix
};

struct IpmResult {
    Vec x, y, s;           // Clarabel conventions: Ax + s = b, y the multipliers (>= 0 on Nonnegative rows)
    SolveQuality status = Unsolved;
    int iters = 0;
    double prim_res = 0, dual_res = 0, gap = 0;
    bool no_iterate = false;
};

// `order` (optional): elimination order over [z (n) ; equality multipliers (in row order)], a permutation of
// 0..n+m_eq-1 listing which unknown comes first.  Empty = natural order.
IpmResult IpmSolve(const Csc& P, const Vec& q, const Csc& A, const Vec& b, const std::vector<char>& is_eq,
                   const std::vector<int>& order, const IpmSettings& s);

IpmResult IpmSolveMpcOrder(const Csc& P, const Vec& q, const Csc& A, const Vec& b, const std::vector<char>& is_eq, int num_dynamics,
                           const IpmSettings& s);

class IpmQpSolver : public QpSolver {
public:
    IpmSettings settings;
    QpSolution Solve(const QpData& data, const Vec& warm_start, bool is_real_time) override;
};

}  // namespace oracle
