// bilevel-gait-gen_b200 -- per-instance solve workspace (HBM, L2-resident while a wave of instances is in flight).
//
// The reference rebuilds triplets -> Eigen::SparseMatrix -> solver every Solve() (mpc_single_rigid_body.cpp:49-107,
// qp_data.cpp:169-178).  Here one solve's QP lives in a *structured* form that is a deterministic function of the
// contact schedule: per-node discretised dynamics (Ad, cd and the spline weights that make up Bd), per-sample force
// rows, per-node foot-box rows and the few equality rows.  The CSC matrix of the reference (bit-identical sparsity)
// is produced from this form only when a caller asks for it (bgg_export_qp_csc).
#pragma once
#include "bgg_types.cuh"

namespace bgg {

struct WsHeader {
    int32_t nfv[kNumEE], npv[kNumEE];     // force / xy-position variables per foot and coordinate
    int32_t fbase[kNumEE], pbase[kNumEE]; // start of the foot's block inside the force / position variables
    int32_t nf, np, nu, n;                // spline variable counts and total decision variables 12(N+1)+nu
    int32_t n_samples;                    // force samples = 10 * (stance segments)
    int32_t n_eebox;                      // (N-3)*4*2 two-sided foot-box rows
    int32_t n_eq;                         // touch-down rows + foot-start rows
    int32_t n_td;                         // touch-down rows (0..8)
    int32_t m_ineq;                       // one-sided inequality rows: 6*n_samples + 2*n_eebox
    int32_t td_flag[kNumEE];
    int32_t error;                        // bit 0: too many knots, bit 1: nu > max_nu, bit 2: too many samples
    int32_t status, iters, ls_iters;
    int32_t no_iterate;                   // the solver stopped before its first finite residual evaluation: u holds nothing
    int32_t refined_iters;                // iterations of the last solve that ran with iterative refinement
    int32_t pass_state;                   // 0 pending, 1 solved by this launch sequence, 2 larger than the caps the first pass was
                                          // launched with: picked up by the second pass (worst-case shared memory), see bgg_capi.cu
    double t0;
    double cost_const;                    // 1/2 phi'P phi + q'phi: condensed objective + cost_const = full objective
    double alpha, cost, qp_cost, prim_res, dual_res, gap, eq_violation, step_norm, merit, merit_dd;
    double ee_box[2];
};

// Linearisation of one node: Ad = I + dt A, cd = dt C and the pieces Bd is made of.  Bd is never stored densely:
//   Bd[3+c , fcol(e,c,i)] = dt * fw[e][i]
//   Bd[9:12, fcol(e,c,i)] = dt * (rel[e] x e_c) * fw[e][i]
//   Bd[9:12, pcol(e,c,i)] = dt * (e_c x f[e])  * pw[e][i]            (single_rigid_body_model.cpp:113-148)
struct NodeLin {
    double Ad[kNx * kNx];
    double cd[kNx];
    double rel[kNumEE][3];    // r_e(t_k) - p_k
    double f[kNumEE][3];      // f_e(t_k)
    double fw[kNumEE][4];     // force-spline weights (same for x, y, z)
    double pw[kNumEE][2];     // xy-position-spline weights (same for x, y)
    int32_t fcnt[kNumEE], foff[kNumEE], pcnt[kNumEE], poff[kNumEE];
};

struct Sample {               // one of the 10 constraint samples of a stance (mpc.cpp:166-209, 352-414)
    double w[4];
    double time;
    int32_t ee, off, cnt, active;   // active == 0 when every weight is exactly 0 (the touch-down sample)
};

struct EqRow {                // touch-down / foot-start rows: w . u_pos[col] = rhs
    double w[2];
    double rhs;
    int32_t col[2];           // column inside the spline variables (>= nf)
    int32_t cnt, pad;         // pad: (foot * 2 + coord) group of the row
};

constexpr int kMaxEq = 16;
constexpr int kMaxContacts = 12;   // contact times (LiftOff / TouchDown knots) per foot inside kMaxKnots

struct GradInfo {             // result record of the gait-gradient kernel
    int32_t status;           // 0 ok, 1 last solve was not `Solved` (the reference returns false), 2 singular system
    int32_t n_theta;
    int32_t nct[kNumEE];      // contact times per foot
    int32_t pad[2];
    double pivot_min;         // smallest |pivot| of the LU factorisation (diagnostic)
    double resid;             // relative residual of the reduced adjoint system after refinement
};
constexpr int kMaxSamples = kNumEE * kMaxStances * kSamplesPerStance;

struct WsLayout {             // byte offsets inside one instance's workspace
    size_t stride;
    size_t hdr, nodes, samples, eq, zprev, H, g, phipos, xoff, u, lam, slack, nueq, zqp, dualx;
    size_t gdx, gdz, gdlam, gdnu, gdnue, gdH, ginfo;   // gait-gradient outputs (csrc/bgg_gradient.cu)
    size_t ktab;                                       // item table of the KKT assembly (csrc/bgg_kkt_mma.cuh)
    size_t ipm_spill;                                  // k_ipm's primal residual rows and foot-box right-hand sides when they do not fit on chip
    int32_t N, max_nu, max_rows, pad;
};

inline WsLayout make_layout(int N, int max_nu) {
    WsLayout L{};
    L.N = N;
    L.max_nu = max_nu;
    L.max_rows = 6 * kMaxSamples + 2 * (N - 3) * 8;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o += (bytes + 127) / 128 * 128; return r; };
    const size_t n_max = static_cast<size_t>(kNx) * (N + 1) + max_nu;
    L.hdr = take(sizeof(WsHeader));
    L.nodes = take(sizeof(NodeLin) * (N + 1));
    L.samples = take(sizeof(Sample) * kMaxSamples);
    L.eq = take(sizeof(EqRow) * kMaxEq);
    L.zprev = take(8 * n_max);
    L.H = take(8 * static_cast<size_t>(max_nu) * max_nu);             // full symmetric nu x nu (row stride nu)
    L.g = take(8 * static_cast<size_t>(max_nu));
    L.phipos = take(8 * static_cast<size_t>(2 * (N - 3)) * max_nu);   // rows (k-4)*2+c of the condensed position map
    L.xoff = take(8 * static_cast<size_t>(kNx) * (N + 1));            // phi_k : x_k = Phi_k u + phi_k
    L.u = take(8 * static_cast<size_t>(max_nu));
    L.lam = take(8 * static_cast<size_t>(L.max_rows));
    L.slack = take(8 * static_cast<size_t>(L.max_rows));
    L.nueq = take(8 * kMaxEq);
    L.zqp = take(8 * n_max);
    L.dualx = take(8 * static_cast<size_t>(kNx) * (N + 1));           // dynamics multipliers (adjoint recursion)
    L.gdx = take(8 * n_max);                                          // P z + q at prev_qp_sol
    L.gdz = take(8 * n_max);                                          // adjoint: dz
    L.gdlam = take(8 * static_cast<size_t>(L.max_rows));              // dlam (kernel row order)
    L.gdnu = take(8 * static_cast<size_t>(kNx) * (N + 1));            // dnu of the dynamics rows
    L.gdnue = take(8 * kMaxEq);                                       // dnu of the touch-down / foot-start rows
    L.gdH = take(8 * kNumEE * kMaxContacts);                          // dH/dtheta, foot-major
    L.ginfo = take(sizeof(GradInfo));
    L.ktab = take(12 * 1536 + 24 * 512);                              // kMaxKktItems x sizeof(KktItem) + kMaxKktPos x sizeof(KktPos)
    L.ipm_spill = take(8 * (static_cast<size_t>(L.max_rows) + 16 * static_cast<size_t>(N - 3)));
    L.stride = o;
    return L;
}

}  // namespace bgg
