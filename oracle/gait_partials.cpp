// TEST INFRASTRUCTURE ONLY -- CPU oracle.  Nothing under oracle/ may be imported, linked or executed by the
// product path; see foot_spline.hpp.
//
// Restatement of the contact-time parameter partials of the assembled QP, the inputs of the gait optimiser's gradient:
//   MPCSingleRigidBody::ComputeParamPartialsClarabel        mpc/mpc_single_rigid_body.cpp:642-792
//   MPC::AddForceBoxConstraintPartials                       mpc/mpc.cpp:416-531
//   MPC::AddFrictionConeConstraintPartials                   mpc/mpc.cpp:240-350
//   MPCSingleRigidBody::AddTDPositionConstraintPartial       mpc/mpc_single_rigid_body.cpp:889-927
// Quirks of the reference are kept as they are: the foot-start partial always lands on rows 0-1 of its block
// (:733-752, `idx` restarts at 0 for every foot), the touch-down partial uses swing_time / 2 where the constraint
// itself uses td_fraction (= 0.75) (:900 vs :855), and the sample-time sensitivity of the force rows
// (dtimedth) is passed to the coefficient partials.
#include <cassert>
#include <stdexcept>

#include "srb_mpc.hpp"

namespace oracle {

namespace {
constexpr int kFbPerForce = 10;   // FB_PER_FORCE, mpc.h:320

void SetRowT(TripletBuilder& b, const Vec& row, int r0, int c0, double scale = 1.0) { b.SetRow(row, r0, c0, scale); }
}  // namespace

void SrbMpc::AddForceBoxConstraintPartials(TripletBuilder& builder, int contact_idx, int start_idx, int ee) const {
    const int force_idx = ForceSplineStartIdx();
    const auto ct = prev_traj_.GetContactTimes();
    int row_idx = 0;
    for (int i = 0; i < ee; i++)
        for (size_t t = 0; t + 1 < ct.at(i).size(); t++)
            if (ct[i][t].type == TouchDown) row_idx += kFbPerForce;
    for (int t = 0; t < contact_idx; t++)
        if (ct.at(ee).at(t).type == TouchDown) row_idx += kFbPerForce;
    if (ct.at(ee).at(contact_idx).type == LiftOff && contact_idx > 0) row_idx -= kFbPerForce;
    const int coord = 2;
    for (int j = 0; j < 2; j++) {   // using_clarabel_: + rows then - rows
        const bool td = ct[ee][contact_idx].type == TouchDown && contact_idx < static_cast<int>(ct[ee].size()) - 1;
        const bool lo = ct[ee][contact_idx].type == LiftOff && contact_idx > 0;
        if (td || lo) {
            const double lower = td ? ct[ee][contact_idx].t : ct[ee][contact_idx - 1].t;
            const double upper = td ? ct[ee][contact_idx + 1].t : ct[ee][contact_idx].t;
            for (int i = 0; i < kFbPerForce; i++) {
                const double frac = static_cast<double>(i) / static_cast<double>(kFbPerForce);
                const double time = frac * (upper - lower) + lower;
                if (!prev_traj_.IsForceMutable(ee, time)) throw std::runtime_error("force not mutable at a sample");
                const double dtimedth = td ? -frac + 1.0 : frac;
                const auto vi = prev_traj_.GetForceSplineIndex(ee, time, coord);
                const Vec p = prev_traj_.Foot(ee).ComputeCoefPartialWrtTime(Force, coord, time, contact_idx, dtimedth);
                SetRowT(builder, p, start_idx + row_idx, force_idx + vi.first, j == 0 ? 1.0 : -1.0);
                row_idx++;
            }
        }
        row_idx += data_.num_force_box / 2 - kFbPerForce;
    }
}

void SrbMpc::AddFrictionConeConstraintPartials(TripletBuilder& builder, int contact_idx, int start_idx, int ee) const {
    const int force_idx = ForceSplineStartIdx();
    const auto ct = prev_traj_.GetContactTimes();
    int row_idx = 0;
    for (int i = 0; i < ee; i++)
        for (size_t t = 0; t + 1 < ct.at(i).size(); t++)
            if (ct[i][t].type == TouchDown) row_idx += 4 * kFbPerForce;
    for (int t = 0; t < contact_idx; t++)
        if (ct.at(ee).at(t).type == TouchDown) row_idx += 4 * kFbPerForce;
    if (ct.at(ee).at(contact_idx).type == LiftOff && contact_idx > 0) row_idx -= 4 * kFbPerForce;
    const bool td = ct[ee][contact_idx].type == TouchDown && contact_idx < static_cast<int>(ct[ee].size()) - 1;
    const bool lo = contact_idx > 0 && ct[ee][contact_idx].type == LiftOff;
    if (!td && !lo) return;
    const double lower = td ? ct[ee][contact_idx].t : ct[ee][contact_idx - 1].t;
    const double upper = td ? ct[ee][contact_idx + 1].t : ct[ee][contact_idx].t;
    for (int i = 0; i < kFbPerForce; i++) {
        for (int coord = 0; coord < 3; coord++) {
            const double frac = static_cast<double>(i) / static_cast<double>(kFbPerForce);
            const double time = frac * (upper - lower) + lower;
            const double dtimedth = td ? -frac + 1.0 : frac;
            const auto vi = prev_traj_.GetForceSplineIndex(ee, time, coord);
            const Vec p = prev_traj_.Foot(ee).ComputeCoefPartialWrtTime(Force, coord, time, contact_idx, dtimedth);
            for (int fc = 0; fc < 4; fc++)
                SetRowT(builder, p, start_idx + row_idx + fc, force_idx + vi.first, friction_pyramid_[fc][coord]);
        }
        row_idx += 4;
    }
}

void SrbMpc::AddTDPositionConstraintPartial(TripletBuilder& builder, Vec& b, int contact_idx, int eq_idx, int ee) const {
    const int start_pos_idx = PosSplineStartIdx();
    int row_idx = 0;
    for (int i = 0; i < ee; i++)
        if (prev_traj_.GetNextContactTime(i, init_time_) - init_time_ < td_fraction_ * prev_traj_.GetCurrentSwingTime(i)) row_idx += 2;
    if (prev_traj_.GetNextContactTime(ee, init_time_) - init_time_ < prev_traj_.GetCurrentSwingTime(ee) / 2) {
        const double td_time = prev_traj_.GetNextContactTime(ee, init_time_);
        double pp[3];
        prev_traj_.GetPositionPartialWrtContactTime(ee, td_time, contact_idx, pp);
        b.at(eq_idx + row_idx) = pp[0];
        b.at(eq_idx + row_idx + 1) = pp[1];
        for (int coord = 0; coord < 2; coord++) {
            const auto vi = prev_traj_.GetPositionSplineIndex(ee, td_time, coord);
            const Vec lin = prev_traj_.Foot(ee).ComputeCoefPartialWrtTime(Position, coord, td_time, contact_idx, 0);
            SetRowT(builder, lin, eq_idx + row_idx, start_pos_idx + vi.first);
            row_idx++;
        }
    }
}

bool SrbMpc::ComputeParamPartialsClarabel(const Traj& traj, ParamPartials& out, int ee, int contact_idx) const {
    if (last_qp_.status != Solved) return false;
    out.dA.Reserve();
    out.dG.Reserve();
    out.num_eq = data_.num_equality;
    out.num_ineq = data_.num_inequality;
    out.num_vars = data_.num_vars;
    out.db.assign(data_.num_equality, 0.0);
    out.dh.assign(data_.num_inequality, 0.0);
    const int N = info_.num_nodes;
    const double dt = info_.integrator_dt;
    int eq_idx = 0, ineq_idx = 0;
    for (Constraint c : data_.constraints) {
        if (c == Dynamics) {
            for (int node = 0; node < N; node++) {
                Mat dA, dB;
                Vec dC;
                model_.ComputeLinearizationPartialWrtContactTimes(dA, dB, dC, traj.GetState(node), traj, GetTime(node), ee, contact_idx);
                for (double& v : dA.a) v = dt * v;
                for (double& v : dB.a) v = dt * v;
                for (double& v : dC) v = dt * v;
                out.dA.SetMatrix(dA, eq_idx + (node + 1) * 12, node * 12);
                out.dA.SetMatrix(dB, eq_idx + (node + 1) * 12, ForceSplineStartIdx());
                for (int i = 0; i < 12; i++) out.db[eq_idx + (node + 1) * 12 + i] = -dC[i];
            }
            eq_idx += data_.num_dynamics;
        } else if (c == EndEffectorLocation) {
            const int spline_offset = PosSplineStartIdx();
            int idx = 2 * ee;
            for (int node = 4; node < N + 1; node++) {
                const double time = GetTime(node);
                for (int coord = 0; coord < 2; coord++) {
                    const auto vi = traj.GetPositionSplineIndex(ee, time, coord);
                    const Vec p = traj.Foot(ee).ComputeCoefPartialWrtTime(Position, coord, time, contact_idx, 0);
                    out.dG.SetRow(p, ineq_idx + idx, spline_offset + vi.first, 1.0);
                    out.dG.SetRow(p, ineq_idx + idx + data_.num_ee_location / 2, spline_offset + vi.first, -1.0);
                    idx++;
                }
                idx += 2 * (4 - 1);
            }
        } else if (c == EndEffectorStart) {
            Mat M(data_.num_start_ee, traj.GetTotalPosSplineVars());
            int idx = 0;
            for (int coord = 0; coord < 2; coord++) {
                const auto vi = prev_traj_.GetPositionSplineIndex(ee, GetTime(0), coord);
                const Vec p = traj.Foot(ee).ComputeCoefPartialWrtTime(Position, coord, GetTime(0), contact_idx, 0);
                for (int k = 0; k < vi.second; k++) M(idx, vi.first + k) = p.at(k);
                idx++;
            }
            out.dA.SetMatrix(M, eq_idx, PosSplineStartIdx());
            eq_idx += data_.num_start_ee;
            ineq_idx += data_.num_ee_location;
        } else if (c == ForceBox) {
            AddForceBoxConstraintPartials(out.dG, contact_idx, ineq_idx, ee);
            ineq_idx += data_.num_force_box;
        } else if (c == FrictionCone) {
            AddFrictionConeConstraintPartials(out.dG, contact_idx, ineq_idx, ee);
            ineq_idx += data_.num_cone;
        } else if (c == TDPosition) {
            AddTDPositionConstraintPartial(out.dA, out.db, contact_idx, eq_idx, ee);
            eq_idx += data_.num_td_pos;
        }
    }
    return true;
}

}  // namespace oracle
