// TEST INFRASTRUCTURE ONLY -- C entry points over the CPU oracle for ctypes (tests/, smoke(), bench.py's CPU
// baseline).  Not part of the product; the product's C-ABI is include/bgg.h.
#include <cmath>
#include <cstring>
#include <limits>
#include <memory>
#include <string>

#include "foot_spline.hpp"
#include "leg_kinematics.hpp"
#include "qp_admm.hpp"
#include "qp_ipm.hpp"
#include "srb_mpc.hpp"

using namespace oracle;

static thread_local std::string g_err;
#define ORC_TRY try {
#define ORC_CATCH(ret)                      \
    }                                       \
    catch (const std::exception& e) {       \
        g_err = e.what();                   \
        return ret;                         \
    }

static const double kNaN = std::numeric_limits<double>::quiet_NaN();

extern "C" {

const char* orc_last_error() { return g_err.c_str(); }
void orc_clear_error() { g_err.clear(); }

// ------------------------------------------------------------------ FootSpline
void* orc_spline_create(int num_contacts, const double* times, int start_in_contact, int num_force_polys) {
    ORC_TRY
    std::vector<double> t(times, times + num_contacts);
    return new FootSpline(num_contacts, t, start_in_contact != 0, num_force_polys);
    ORC_CATCH(nullptr)
}
void orc_spline_destroy(void* h) { delete static_cast<FootSpline*>(h); }
void* orc_spline_clone(void* h) { return new FootSpline(*static_cast<FootSpline*>(h)); }

double orc_spline_value(void* h, int type, int coord, double t) {
    ORC_TRY return static_cast<FootSpline*>(h)->ValueAt(static_cast<SplineType>(type), coord, t);
    ORC_CATCH(kNaN)
}
int orc_spline_lin(void* h, int type, int coord, double t, double* out) {
    ORC_TRY
    const auto v = static_cast<FootSpline*>(h)->GetPolyVarsLin(static_cast<SplineType>(type), coord, t);
    std::copy(v.begin(), v.end(), out);
    return static_cast<int>(v.size());
    ORC_CATCH(-1)
}
int orc_spline_vars_idx(void* h, int type, int coord, double t, int* idx, int* cnt) {
    ORC_TRY
    const auto p = static_cast<FootSpline*>(h)->GetVarsIdx(static_cast<SplineType>(type), coord, t);
    *idx = p.first;
    *cnt = p.second;
    return 0;
    ORC_CATCH(-1)
}
int orc_spline_is_force_mutable(void* h, double t) {
    ORC_TRY return static_cast<FootSpline*>(h)->IsForceMutable(t) ? 1 : 0;
    ORC_CATCH(-1)
}
int orc_spline_is_in_contact(void* h, double t) {
    ORC_TRY return static_cast<FootSpline*>(h)->IsInContact(t) ? 1 : 0;
    ORC_CATCH(-1)
}
int orc_spline_add_poly(void* h, double dt) {
    ORC_TRY static_cast<FootSpline*>(h)->AddPoly(dt);
    return 0;
    ORC_CATCH(-1)
}
int orc_spline_remove_poly(void* h, double t) {
    ORC_TRY static_cast<FootSpline*>(h)->RemovePoly(t);
    return 0;
    ORC_CATCH(-1)
}
double orc_spline_partial(void* h, int type, int coord, double t, int time_idx) {
    ORC_TRY return static_cast<FootSpline*>(h)->ComputePartialWrtTime(static_cast<SplineType>(type), coord, t, time_idx);
    ORC_CATCH(kNaN)
}
int orc_spline_coef_partial(void* h, int type, int coord, double t, int time_idx, double dtwdth, double* out) {
    ORC_TRY
    const auto v = static_cast<FootSpline*>(h)->ComputeCoefPartialWrtTime(static_cast<SplineType>(type), coord, t, time_idx, dtwdth);
    std::copy(v.begin(), v.end(), out);
    return static_cast<int>(v.size());
    ORC_CATCH(-1)
}
int orc_spline_set_vars(void* h, int type, int coord, int node, double v0, double v1) {
    ORC_TRY static_cast<FootSpline*>(h)->SetVars(static_cast<SplineType>(type), coord, node, v0, v1);
    return 0;
    ORC_CATCH(-1)
}
int orc_spline_set_contact_times(void* h, const double* t, int n) {
    ORC_TRY
    FootSpline* s = static_cast<FootSpline*>(h);
    std::vector<KnotTime> ct = s->GetContactTimes();
    if (static_cast<int>(ct.size()) != n) throw std::runtime_error("contact time count mismatch");
    for (int i = 0; i < n; i++) ct[i].t = t[i];
    s->SetContactTimes(ct);
    return 0;
    ORC_CATCH(-1)
}
int orc_spline_num_nodes(void* h) { return static_cast<FootSpline*>(h)->GetNumNodes(); }
int orc_spline_num_contacts(void* h) { return static_cast<FootSpline*>(h)->GetNumContacts(); }
int orc_spline_node_type(void* h, int type, int coord, int node) {
    ORC_TRY return static_cast<FootSpline*>(h)->GetNodeType(static_cast<SplineType>(type), coord, node);
    ORC_CATCH(-1)
}
int orc_spline_mutable_nodes(void* h, int type, int coord, int* out) {
    ORC_TRY
    const auto v = static_cast<FootSpline*>(h)->GetMutableNodes(static_cast<SplineType>(type), coord);
    std::copy(v.begin(), v.end(), out);
    return static_cast<int>(v.size());
    ORC_CATCH(-1)
}
int orc_spline_times(void* h, double* out, int* types) {
    const auto& kt = static_cast<FootSpline*>(h)->KnotTimes();
    for (size_t i = 0; i < kt.size(); i++) {
        out[i] = kt[i].t;
        if (types) types[i] = kt[i].type;
    }
    return static_cast<int>(kt.size());
}
int orc_spline_knots(void* h, int type, int coord, int* types, double* vals /* [n][2] */) {
    const auto& k = static_cast<FootSpline*>(h)->Knots(static_cast<SplineType>(type), coord);
    for (size_t i = 0; i < k.size(); i++) {
        types[i] = k[i].type;
        vals[2 * i] = k[i].v[0];
        vals[2 * i + 1] = k[i].v[1];
    }
    return static_cast<int>(k.size());
}
int orc_spline_as_qp_vec(void* h, int type, int coord, double* out) {
    ORC_TRY
    const auto v = static_cast<FootSpline*>(h)->GetSplineAsQPVec(static_cast<SplineType>(type), coord);
    std::copy(v.begin(), v.end(), out);
    return static_cast<int>(v.size());
    ORC_CATCH(-1)
}
int orc_spline_total_poly_vars(void* h, int type, int coord) {
    return static_cast<FootSpline*>(h)->GetTotalPolyVars(static_cast<SplineType>(type), coord);
}
double orc_spline_end_time(void* h) { return static_cast<FootSpline*>(h)->GetEndTime(); }
double orc_spline_start_time(void* h) { return static_cast<FootSpline*>(h)->GetStartTime(); }
double orc_spline_next_td(void* h, double t) {
    ORC_TRY return static_cast<FootSpline*>(h)->GetNextTouchDownTime(t);
    ORC_CATCH(kNaN)
}
double orc_spline_swing_time(void* h, double t) {
    ORC_TRY return static_cast<FootSpline*>(h)->GetSwingTime(t);
    ORC_CATCH(kNaN)
}
int orc_spline_set_to_touchdown(void* h, double t) {
    ORC_TRY static_cast<FootSpline*>(h)->SetToTouchdown(t);
    return 0;
    ORC_CATCH(-1)
}
int orc_spline_lower(void* h, int type, int coord, double t) {
    ORC_TRY return static_cast<FootSpline*>(h)->GetLowerNodeIdx(static_cast<SplineType>(type), coord, t);
    ORC_CATCH(-1)
}
int orc_spline_upper(void* h, int type, int coord, double t) {
    ORC_TRY return static_cast<FootSpline*>(h)->GetUpperNodeIdx(static_cast<SplineType>(type), coord, t);
    ORC_CATCH(-1)
}

// ------------------------------------------------------------------ generic ADMM QP (two-sided form)
struct OrcAdmmSettings {
    double rho, sigma, alpha, eps_abs, eps_rel, eps_prim_inf, eps_dual_inf, adaptive_rho_tolerance;
    int max_iter, scaling, check_termination, adaptive_rho, adaptive_rho_interval;
};
static AdmmSettings FromC(const OrcAdmmSettings& c) {
    AdmmSettings s;
    s.rho = c.rho; s.sigma = c.sigma; s.alpha = c.alpha; s.eps_abs = c.eps_abs; s.eps_rel = c.eps_rel;
    s.eps_prim_inf = c.eps_prim_inf; s.eps_dual_inf = c.eps_dual_inf; s.adaptive_rho_tolerance = c.adaptive_rho_tolerance;
    s.max_iter = c.max_iter; s.scaling = c.scaling; s.check_termination = c.check_termination;
    s.adaptive_rho = c.adaptive_rho != 0; s.adaptive_rho_interval = c.adaptive_rho_interval;
    return s;
}
void orc_admm_default_settings(OrcAdmmSettings* c) {
    const AdmmSettings s;
    c->rho = s.rho; c->sigma = s.sigma; c->alpha = s.alpha; c->eps_abs = s.eps_abs; c->eps_rel = s.eps_rel;
    c->eps_prim_inf = s.eps_prim_inf; c->eps_dual_inf = s.eps_dual_inf; c->adaptive_rho_tolerance = s.adaptive_rho_tolerance;
    c->max_iter = s.max_iter; c->scaling = s.scaling; c->check_termination = s.check_termination;
    c->adaptive_rho = s.adaptive_rho; c->adaptive_rho_interval = s.adaptive_rho_interval;
}
static Csc MakeCsc(int rows, int cols, const int* colptr, const int* rowidx, const double* val) {
    Csc m;
    m.rows = rows;
    m.cols = cols;
    m.colptr.assign(colptr, colptr + cols + 1);
    m.rowidx.assign(rowidx, rowidx + colptr[cols]);
    m.val.assign(val, val + colptr[cols]);
    return m;
}
// returns status (SolveQuality); info = [iters, prim_res, dual_res, rho_final, rho_updates]
int orc_admm_solve(int n, int m, const int* Pcolptr, const int* Prowidx, const double* Pval, const double* q,
                   const int* Acolptr, const int* Arowidx, const double* Aval, const double* l, const double* u,
                   const double* x0, const double* y0, const OrcAdmmSettings* cs, double* x, double* y, double* z,
                   double* info) {
    ORC_TRY
    const Csc P = MakeCsc(n, n, Pcolptr, Prowidx, Pval), A = MakeCsc(m, n, Acolptr, Arowidx, Aval);
    const AdmmResult r = AdmmSolve(P, Vec(q, q + n), A, Vec(l, l + m), Vec(u, u + m), Vec(x0, x0 + n), Vec(y0, y0 + m), FromC(*cs));
    if (!r.x.empty()) {
        std::copy(r.x.begin(), r.x.end(), x);
        std::copy(r.y.begin(), r.y.end(), y);
        std::copy(r.z.begin(), r.z.end(), z);
    }
    info[0] = r.iters; info[1] = r.prim_res; info[2] = r.dual_res; info[3] = r.rho_final; info[4] = r.rho_updates;
    return r.status;
    ORC_CATCH(-1)
}

// ------------------------------------------------------------------ SrbMpc
struct OrcMpcInfo {
    int num_nodes;
    double friction_coef, integrator_dt, force_bound, swing_height, foot_offset, ee_box_x, ee_box_y, force_cost;
};
struct OrcRobotConsts {
    double mass, Ir[9], Ir_inv[9], hip_xy[8], gravity[3];
};
struct MpcHandle {
    std::shared_ptr<AdmmQpSolver> solver;     // OSQP restatement (the reference's test-only OSQPInterface)
    std::shared_ptr<IpmQpSolver> ipm;         // interior point (the reference's live ClarabelInterface)
    std::unique_ptr<SrbMpc> mpc;
};

void* orc_mpc_create(const OrcMpcInfo* ci, const OrcRobotConsts* cr) {
    ORC_TRY
    MpcInfo info;
    info.num_nodes = ci->num_nodes; info.friction_coef = ci->friction_coef; info.integrator_dt = ci->integrator_dt;
    info.force_bound = ci->force_bound; info.swing_height = ci->swing_height; info.foot_offset = ci->foot_offset;
    info.ee_box_size[0] = ci->ee_box_x; info.ee_box_size[1] = ci->ee_box_y; info.force_cost = ci->force_cost;
    RobotConsts rc;
    rc.mass = cr->mass;
    std::memcpy(rc.Ir, cr->Ir, sizeof rc.Ir);
    std::memcpy(rc.Ir_inv, cr->Ir_inv, sizeof rc.Ir_inv);
    std::memcpy(rc.hip_xy, cr->hip_xy, sizeof rc.hip_xy);
    std::memcpy(rc.gravity, cr->gravity, sizeof rc.gravity);
    auto* h = new MpcHandle;
    h->solver = std::make_shared<AdmmQpSolver>();
    h->ipm = std::make_shared<IpmQpSolver>();
    h->mpc.reset(new SrbMpc(info, rc, h->ipm));   // the live path: interior point
    return h;
    ORC_CATCH(nullptr)
}
void orc_mpc_destroy(void* h) { delete static_cast<MpcHandle*>(h); }
void* orc_mpc_clone(void* h) {
    auto* src = static_cast<MpcHandle*>(h);
    auto* dst = new MpcHandle;
    dst->solver = std::make_shared<AdmmQpSolver>(*src->solver);
    dst->ipm = std::make_shared<IpmQpSolver>(*src->ipm);
    dst->mpc.reset(new SrbMpc(*src->mpc));
    dst->mpc->SetSolver(dst->ipm);
    return dst;
}
static SrbMpc& M(void* h) { return *static_cast<MpcHandle*>(h)->mpc; }

// which: 0 = interior point (default, the reference's live Clarabel path), 1 = ADMM (its OSQPInterface)
void orc_mpc_select_solver(void* h, int which) {
    auto* mh = static_cast<MpcHandle*>(h);
    if (which == 1) mh->mpc->SetSolver(mh->solver);
    else mh->mpc->SetSolver(mh->ipm);
}
void orc_mpc_set_ipm(void* h, double tol_feas, double tol_gap, int max_iter, int refine) {
    auto* mh = static_cast<MpcHandle*>(h);
    if (tol_feas > 0) mh->ipm->settings.tol_feas = tol_feas;
    if (tol_gap > 0) mh->ipm->settings.tol_gap = tol_gap;
    if (max_iter > 0) mh->ipm->settings.max_iter = max_iter;
    if (refine >= 0) mh->ipm->settings.refine = refine;
}
// generic Clarabel-form interior-point solve; info = [iters, prim_res, dual_res, gap]
int orc_ipm_solve(int n, int m, const int* Pcolptr, const int* Prowidx, const double* Pval, const double* q,
                  const int* Acolptr, const int* Arowidx, const double* Aval, const double* b, const char* is_eq,
                  double tol, int max_iter, double* x, double* y, double* s, double* info) {
    ORC_TRY
    const Csc P = MakeCsc(n, n, Pcolptr, Prowidx, Pval), A = MakeCsc(m, n, Acolptr, Arowidx, Aval);
    IpmSettings st;
    if (tol > 0) st.tol_feas = st.tol_gap = tol;
    if (max_iter > 0) st.max_iter = max_iter;
    const IpmResult r = IpmSolve(P, Vec(q, q + n), A, Vec(b, b + m), std::vector<char>(is_eq, is_eq + m), {}, st);
    if (!r.x.empty()) {
        std::copy(r.x.begin(), r.x.end(), x);
        std::copy(r.y.begin(), r.y.end(), y);
        std::copy(r.s.begin(), r.s.end(), s);
    }
    info[0] = r.iters; info[1] = r.prim_res; info[2] = r.dual_res; info[3] = r.gap;
    return r.status;
    ORC_CATCH(-1)
}

void orc_mpc_set_admm(void* h, const OrcAdmmSettings* initial, const OrcAdmmSettings* real_time) {
    auto* mh = static_cast<MpcHandle*>(h);
    if (initial) mh->solver->initial = FromC(*initial);
    if (real_time) mh->solver->real_time = FromC(*real_time);
}
void orc_mpc_get_admm(void* h, OrcAdmmSettings* initial, OrcAdmmSettings* real_time) {
    auto* mh = static_cast<MpcHandle*>(h);
    auto put = [](const AdmmSettings& s, OrcAdmmSettings* c) {
        c->rho = s.rho; c->sigma = s.sigma; c->alpha = s.alpha; c->eps_abs = s.eps_abs; c->eps_rel = s.eps_rel;
        c->eps_prim_inf = s.eps_prim_inf; c->eps_dual_inf = s.eps_dual_inf; c->adaptive_rho_tolerance = s.adaptive_rho_tolerance;
        c->max_iter = s.max_iter; c->scaling = s.scaling; c->check_termination = s.check_termination;
        c->adaptive_rho = s.adaptive_rho; c->adaptive_rho_interval = s.adaptive_rho_interval;
    };
    if (initial) put(mh->solver->initial, initial);
    if (real_time) put(mh->solver->real_time, real_time);
}

// state_des is a tangent (12) state; Q / Phi row-major 12x12
void orc_mpc_set_costs(void* h, const double* state_des, const double* Q, const double* Phi, const double* Phi_w) {
    Mat Qm(12, 12), Pm(12, 12);
    std::copy(Q, Q + 144, Qm.a.begin());
    std::copy(Phi, Phi + 144, Pm.a.begin());
    M(h).AddQuadraticTrackingCost(Vec(state_des, state_des + 12), Qm);
    M(h).SetQuadraticFinalCost(Pm);
    M(h).SetLinearFinalCost(Vec(Phi_w, Phi_w + 12));
}
void orc_mpc_set_warm_states(void* h, const double* states /* (N+1) x 13 */) {
    const int N = M(h).Info().num_nodes;
    std::vector<Vec> s;
    for (int i = 0; i <= N; i++) s.emplace_back(states + 13 * i, states + 13 * i + 13);
    M(h).SetStateTrajectoryWarmStart(s);
}
static std::vector<std::array<double, 3>> EE(const double* ee) {
    std::vector<std::array<double, 3>> v(4);
    for (int e = 0; e < 4; e++)
        for (int c = 0; c < 3; c++) v[e][c] = ee[3 * e + c];
    return v;
}
int orc_mpc_assemble(void* h, const double* state, double t0, const double* ee_start) {
    ORC_TRY M(h).AssembleOnly(Vec(state, state + 13), t0, EE(ee_start));
    return 0;
    ORC_CATCH(-1)
}
int orc_mpc_solve(void* h, const double* state, double t0, const double* ee_start, int real_time) {
    ORC_TRY
    if (real_time) M(h).GetRealTimeUpdate(Vec(state, state + 13), t0, EE(ee_start));
    else M(h).Solve(Vec(state, state + 13), t0, EE(ee_start));
    return M(h).LastQp().status;
    ORC_CATCH(-1)
}
int orc_mpc_initial_run(void* h, const double* state, const double* ee_start) {
    ORC_TRY M(h).CreateInitialRun(Vec(state, state + 13), EE(ee_start));
    return M(h).LastQp().status;
    ORC_CATCH(-1)
}
int orc_mpc_set_contact_times(void* h, int ee, const double* t, int n) {
    ORC_TRY
    auto ct = M(h).Trajectory().GetContactTimes();
    if (static_cast<int>(ct.at(ee).size()) != n) throw std::runtime_error("contact time count mismatch");
    for (int i = 0; i < n; i++) ct[ee][i].t = t[i];
    M(h).UpdateContactTimes(ct);
    return 0;
    ORC_CATCH(-1)
}
// MPC::AdjustForCurrentContacts (mpc.cpp:1195-1203)
int orc_mpc_adjust_for_contacts(void* h, double time, const int* in_contact) {
    ORC_TRY
    M(h).AdjustForCurrentContacts(time, {in_contact[0] != 0, in_contact[1] != 0, in_contact[2] != 0, in_contact[3] != 0});
    return 0;
    ORC_CATCH(-1)
}
// sizes: [n, m, nnzA, nnzP, num_dyn, num_force_box, num_cone, num_ee_loc, num_td, num_start, nf, np, num_eq, num_ineq]
void orc_mpc_sizes(void* h, int* out) {
    const QpData& d = M(h).Data();
    out[0] = d.num_vars; out[1] = d.Total(); out[2] = d.A.nnz(); out[3] = d.P.nnz(); out[4] = d.num_dynamics;
    out[5] = d.num_force_box; out[6] = d.num_cone; out[7] = d.num_ee_location; out[8] = d.num_td_pos;
    out[9] = d.num_start_ee; out[10] = M(h).Trajectory().GetTotalForceSplineVars();
    out[11] = M(h).Trajectory().GetTotalPosSplineVars(); out[12] = d.num_equality; out[13] = d.num_inequality;
}
void orc_mpc_get_A(void* h, int* colptr, int* rowidx, double* val) {
    const Csc& A = M(h).Data().A;
    std::copy(A.colptr.begin(), A.colptr.end(), colptr);
    std::copy(A.rowidx.begin(), A.rowidx.end(), rowidx);
    std::copy(A.val.begin(), A.val.end(), val);
}
void orc_mpc_get_P(void* h, int* colptr, int* rowidx, double* val) {
    const Csc& P = M(h).Data().P;
    std::copy(P.colptr.begin(), P.colptr.end(), colptr);
    std::copy(P.rowidx.begin(), P.rowidx.end(), rowidx);
    std::copy(P.val.begin(), P.val.end(), val);
}
void orc_mpc_get_vectors(void* h, double* q, double* ub, char* is_eq) {
    const QpData& d = M(h).Data();
    std::copy(d.cost_linear.begin(), d.cost_linear.end(), q);
    std::copy(d.ub.begin(), d.ub.end(), ub);
    const auto eq = d.RowIsEquality();
    std::copy(eq.begin(), eq.end(), is_eq);
}
void orc_mpc_get_prev_qp_sol(void* h, double* z) {
    const Vec& v = M(h).PrevQpSol();
    std::copy(v.begin(), v.end(), z);
}
// info = [status, iters, prim_res, dual_res]
void orc_mpc_get_qp_solution(void* h, double* x, double* dual, double* slack, double* info) {
    const QpSolution& s = M(h).LastQp();
    if (x) std::copy(s.x.begin(), s.x.end(), x);
    if (dual) std::copy(s.dual.begin(), s.dual.end(), dual);
    if (slack) std::copy(s.slack.begin(), s.slack.end(), slack);
    info[0] = s.status; info[1] = s.iters; info[2] = s.prim_res; info[3] = s.dual_res;
}
// stats = [alpha, eq_violation, step_norm, cost, merit, merit_dd, status, qp_iters, ee_box_x, ee_box_y]
void orc_mpc_get_stats(void* h, double* out) {
    const SolveStats& s = M(h).LastStats();
    out[0] = s.alpha; out[1] = s.eq_violation; out[2] = s.step_norm; out[3] = s.cost; out[4] = s.merit;
    out[5] = s.merit_dd; out[6] = s.status; out[7] = s.qp_iters;
    out[8] = M(h).Info().ee_box_size[0]; out[9] = M(h).Info().ee_box_size[1];
}
// dense per-node discretised dynamics of the last assembly: Ad [N][12][12], Bd [N][12][nu], cd [N][12]
void orc_mpc_get_node_dynamics(void* h, double* Ad, double* Bd, double* cd) {
    const auto& A = M(h).NodeA();
    const auto& B = M(h).NodeB();
    const auto& C = M(h).NodeC();
    for (size_t k = 0; k < A.size(); k++) {
        std::copy(A[k].a.begin(), A[k].a.end(), Ad + k * 144);
        std::copy(B[k].a.begin(), B[k].a.end(), Bd + k * B[k].a.size());
        std::copy(C[k].begin(), C[k].end(), cd + k * 12);
    }
}
void orc_mpc_get_states(void* h, double* states /* (N+1) x 13 */) {
    const Traj& t = M(h).Trajectory();
    for (int i = 0; i < t.NumStates(); i++) std::copy(t.GetState(i).begin(), t.GetState(i).end(), states + 13 * i);
}
void orc_mpc_set_state(void* h, int node, const double* s) {
    Traj t = M(h).Trajectory();
    t.SetState(node, Vec(s, s + 13));
    M(h).SetWarmStartTrajectory(t);
}
// Borrowed pointer to foot `ee`'s spline inside the MPC's trajectory (valid until the next solve / clone).
void* orc_mpc_foot(void* h, int ee) { return const_cast<FootSpline*>(&M(h).Trajectory().Foot(ee)); }
double orc_mpc_init_time(void* h) { return M(h).Trajectory().InitTime(); }
double orc_mpc_cost(void* h) { return M(h).GetCost(); }
void orc_mpc_force_at(void* h, int ee, double t, double* out) { M(h).Trajectory().GetForce(ee, t, out); }
void orc_mpc_ee_at(void* h, int ee, double t, double* out) { M(h).Trajectory().GetEndEffectorLocation(ee, t, out); }

// Contact-time parameter partials of the last solve's QP (gait_partials.cpp).  Triplets are written up to `cap`
// entries each; returns 0, 1 when the last solve was not `Solved` (the reference returns false), -1 on error.
// counts = [nnz dA, nnz dG, num_eq, num_ineq]
int orc_mpc_param_partials(void* h, int ee, int contact_idx, int cap, int* counts, int* Ar, int* Ac, double* Av, int* Gr,
                           int* Gc, double* Gv, double* db) {
    ORC_TRY
    ParamPartials pp;
    if (!M(h).ComputeParamPartialsClarabel(M(h).Trajectory(), pp, ee, contact_idx)) return 1;
    counts[0] = static_cast<int>(pp.dA.v.size());
    counts[1] = static_cast<int>(pp.dG.v.size());
    counts[2] = pp.num_eq;
    counts[3] = pp.num_ineq;
    if (counts[0] > cap || counts[1] > cap) throw std::runtime_error("triplet capacity too small");
    std::copy(pp.dA.ri.begin(), pp.dA.ri.end(), Ar);
    std::copy(pp.dA.ci.begin(), pp.dA.ci.end(), Ac);
    std::copy(pp.dA.v.begin(), pp.dA.v.end(), Av);
    std::copy(pp.dG.ri.begin(), pp.dG.ri.end(), Gr);
    std::copy(pp.dG.ci.begin(), pp.dG.ci.end(), Gc);
    std::copy(pp.dG.v.begin(), pp.dG.v.end(), Gv);
    std::copy(pp.db.begin(), pp.db.end(), db);
    return 0;
    ORC_CATCH(-1)
}
// number of contact times per foot (Trajectory::GetNumContactNodes) and their values / types
int orc_mpc_num_contacts(void* h, int ee) { return M(h).Trajectory().Foot(ee).GetNumContacts(); }
void orc_mpc_get_contact_times(void* h, int ee, double* t, int* type) {
    const auto ct = M(h).Trajectory().GetContactTimes();
    for (size_t i = 0; i < ct.at(ee).size(); i++) {
        t[i] = ct[ee][i].t;
        type[i] = static_cast<int>(ct[ee][i].type);
    }
}

// merit evaluation taps (mpc.cpp:749-788) for the line-search kernel's parity test
double orc_mpc_merit(void* h, const double* z) {
    ORC_TRY
    const int n = M(h).Data().num_vars;
    return M(h).GetMeritValue(Vec(z, z + n));
    ORC_CATCH(kNaN)
}

// quaternion helpers
void orc_quat_log3(const double* q, double* out) { QuatLog3(q, out); }
void orc_quat_exp3(const double* v, double* out) { QuatExp3(v, out); }

// ---- inverse kinematics (leg_kinematics.cpp).  `kin_flat`: 4 legs x (t[4][3], R[4][9], axis[3][3]) = 228 doubles
static kin::RobotKin Kin(const double* kin_flat) {
    kin::RobotKin rk;
    static_assert(sizeof(kin::RobotKin) == 228 * sizeof(double), "RobotKin is 228 packed doubles");
    std::memcpy(&rk, kin_flat, sizeof rk);
    return rk;
}
void orc_kin_exp6(const double* nu, double* R, double* p) {
    kin::Se3 M;
    kin::Exp6(nu, M);
    std::memcpy(R, M.R, sizeof M.R);
    std::memcpy(p, M.p, sizeof M.p);
}
void orc_kin_log6(const double* R, const double* p, double* out) {
    kin::Se3 M;
    std::memcpy(M.R, R, sizeof M.R);
    std::memcpy(M.p, p, sizeof M.p);
    kin::Log6(M, out);
}
void orc_kin_jlog6(const double* R, const double* p, double* J) {
    kin::Se3 M;
    std::memcpy(M.R, R, sizeof M.R);
    std::memcpy(M.p, p, sizeof M.p);
    kin::Jlog6(M, J);
}
// feet: world positions [4][3] and rotations [4][9]; J: the 6 x 18 LOCAL foot-frame Jacobian of foot `ee`
void orc_kin_fk(const double* kin_flat, const double* q, int ee, double* feet_p, double* feet_R, double* J) {
    const kin::RobotKin rk = Kin(kin_flat);
    kin::Se3 joints[13], feet[4];
    kin::ForwardKinematics(rk, q, joints, feet);
    for (int e = 0; e < 4; e++) {
        std::memcpy(feet_p + 3 * e, feet[e].p, sizeof feet[e].p);
        std::memcpy(feet_R + 9 * e, feet[e].R, sizeof feet[e].R);
    }
    kin::FootJacobianLocal(rk, joints, feet, ee, J);
}
void orc_kin_integrate(const double* q, const double* v, double* out) { kin::Integrate(q, v, out); }
int orc_ik(const double* kin_flat, const double* state, const double* ee_des, const double* joint_guess, double* q_out, int* iters) {
    double ee[4][3];
    std::memcpy(ee, ee_des, sizeof ee);
    return kin::InverseKinematics(Kin(kin_flat), state, ee, joint_guess, q_out, iters);
}
int orc_mpc_targets_from_traj(void* h, const double* kin_flat, double time, double dt, double mass, const double* Ir_inv, double* q_des,
                              double* v_des, double* force_des) {
    ORC_TRY return kin::GetTargetsFromTraj(Kin(kin_flat), M(h).Trajectory(), time, dt, mass, Ir_inv, q_des, v_des, force_des);
    ORC_CATCH(-1)
}

}  // extern "C"
