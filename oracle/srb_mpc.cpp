// TEST INFRASTRUCTURE ONLY -- CPU oracle (see srb_mpc.hpp for scope and citations).
#include "srb_mpc.hpp"

#include <algorithm>
#include <cassert>
#include <cmath>
#include <limits>
#include <numeric>

namespace oracle {

// ------------------------------------------------------------------------------------------------ sparse helpers
void Csc::mul(const double* x, double* y) const {
    std::fill(y, y + rows, 0.0);
    for (int j = 0; j < cols; j++)
        for (int k = colptr[j]; k < colptr[j + 1]; k++) y[rowidx[k]] += val[k] * x[j];
}
void Csc::mul_t(const double* x, double* y) const {
    for (int j = 0; j < cols; j++) {
        double s = 0;
        for (int k = colptr[j]; k < colptr[j + 1]; k++) s += val[k] * x[rowidx[k]];
        y[j] = s;
    }
}

// Eigen::SparseMatrix::setFromTriplets (qp_data.cpp:176-177): column-compressed, duplicates summed, explicit zeros
// that result from a sum are kept, rows ascending inside a column.
Csc TripletBuilder::Build(int rows, int cols) const {
    Csc m;
    m.rows = rows;
    m.cols = cols;
    std::vector<size_t> order(v.size());
    std::iota(order.begin(), order.end(), size_t(0));
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) {
        if (ci[a] != ci[b]) return ci[a] < ci[b];
        return ri[a] < ri[b];
    });
    m.colptr.assign(cols + 1, 0);
    int last_r = -1, last_c = -1;
    for (size_t o : order) {
        if (ri[o] < 0 || ri[o] >= rows || ci[o] < 0 || ci[o] >= cols) throw std::runtime_error("triplet out of range");
        if (ri[o] == last_r && ci[o] == last_c) {
            m.val.back() += v[o];
        } else {
            m.rowidx.push_back(ri[o]);
            m.val.push_back(v[o]);
            m.colptr[ci[o] + 1]++;
            last_r = ri[o];
            last_c = ci[o];
        }
    }
    for (int j = 0; j < cols; j++) m.colptr[j + 1] += m.colptr[j];
    return m;
}

// ------------------------------------------------------------------------------------------------ QpData
int QpData::Total() const {   // qp_data.cpp:61-97
    int n = 0;
    for (Constraint c : constraints) {
        switch (c) {
            case Dynamics: n += num_dynamics; break;
            case ForceBox: n += num_force_box; break;
            case FrictionCone: n += num_cone; break;
            case EndEffectorLocation: n += num_ee_location; break;
            case TDPosition: n += num_td_pos; break;
            case EndEffectorStart: n += num_start_ee; break;
        }
    }
    return n;
}

void QpData::InitQPMats() {   // qp_data.cpp:99-167 (using_clarabel_ == true)
    constraint_mat.Reserve();
    cost_mat.Reserve();
    dynamics_constants.assign(num_dynamics, 0.0);
    ee_location_lb.assign(num_ee_location / 2, 0.0);
    ee_location_ub.assign(num_ee_location / 2, 0.0);
    start_ee_constants.assign(num_start_ee, 0.0);
    force_box_lb.assign(num_force_box / 2, 0.0);
    force_box_ub.assign(num_force_box / 2, 0.0);
    friction_cone_ub.assign(num_cone, 0.0);
    td_pos_constants.assign(num_td_pos, 0.0);
    cost_linear.assign(num_vars, 0.0);
    ub.assign(Total(), 0.0);
}

void QpData::ConstructSparseMats() {   // qp_data.cpp:169-178
    A = constraint_mat.Build(Total(), num_vars);
    P = cost_mat.Build(num_vars, num_vars);
}

void QpData::ConstructVectors() {   // qp_data.cpp:200-289 (using_clarabel_ == true)
    int idx = 0;
    num_inequality = 0;
    num_equality = 0;
    auto put = [&](const Vec& v, double sign) {
        for (double x : v) ub[idx++] = sign * x;
    };
    for (Constraint c : constraints) {
        switch (c) {
            case Dynamics: put(dynamics_constants, 1); num_equality += num_dynamics; break;
            case ForceBox: put(force_box_ub, 1); put(force_box_lb, -1); num_inequality += num_force_box; break;
            case FrictionCone: put(friction_cone_ub, 1); num_inequality += num_cone; break;
            case EndEffectorLocation: put(ee_location_ub, 1); put(ee_location_lb, -1); num_inequality += num_ee_location; break;
            case TDPosition: put(td_pos_constants, 1); num_equality += num_td_pos; break;
            case EndEffectorStart: put(start_ee_constants, 1); num_equality += num_start_ee; break;
        }
    }
    assert(idx == Total());
}

std::vector<char> QpData::RowIsEquality() const {
    std::vector<char> eq;
    for (Constraint c : constraints) {
        switch (c) {
            case Dynamics: eq.insert(eq.end(), num_dynamics, 1); break;
            case ForceBox: eq.insert(eq.end(), num_force_box, 0); break;
            case FrictionCone: eq.insert(eq.end(), num_cone, 0); break;
            case EndEffectorLocation: eq.insert(eq.end(), num_ee_location, 0); break;
            case TDPosition: eq.insert(eq.end(), num_td_pos, 1); break;
            case EndEffectorStart: eq.insert(eq.end(), num_start_ee, 1); break;
        }
    }
    return eq;
}

// ------------------------------------------------------------------------------------------------ Traj
Traj::Traj(int len, const std::vector<std::vector<double>>& switching_times, double node_dt, double swing_height,
           double foot_offset)
    : swing_height_(swing_height), foot_offset_(foot_offset), node_dt_(node_dt) {   // trajectory.cpp:11-48
    states_.assign(len, Vec(13, 0.0));
    for (size_t i = 0; i < switching_times.size(); i++) {
        const bool in_contact = (i == 1 || i == 2);
        ee_.emplace_back(static_cast<int>(switching_times[i].size()), switching_times[i], in_contact, 3);
    }
    init_time_ = 0;
    UpdateSplineVarsCount();
    SetSwingPosZ();
}

void Traj::UpdateSplineVarsCount() {   // trajectory.cpp:252-265
    pos_vars_ = 0;
    force_vars_ = 0;
    for (const auto& s : ee_) {
        pos_vars_ += 2 * s.GetTotalPolyVars(Position, 0);
        force_vars_ += 3 * s.GetTotalPolyVars(Force, 0);
    }
}

void Traj::SetSwingPosZ() {   // trajectory.cpp:303-317
    for (auto& s : ee_) {
        for (int node : s.GetMutableNodes(Position, 2)) {
            if (s.GetNodeType(Position, 2, node) == FullDeriv) s.SetVars(Position, 2, node, swing_height_, 0);
            else s.SetVars(Position, 2, node, foot_offset_, 0);
        }
    }
}

void Traj::UpdateForceSpline(int ee, int coord, const double* vars, int n) {   // trajectory.cpp:83-97
    int idx = 0;
    for (int node : ee_[ee].GetMutableNodes(Force, coord)) {
        ee_[ee].SetVars(Force, coord, node, vars[idx], vars[idx + 1]);
        idx += 2;
    }
    assert(idx == n);
    (void)n;
}

void Traj::UpdatePositionSpline(int ee, int coord, const double* vars, int n) {   // trajectory.cpp:99-111
    int idx = 0;
    for (int node : ee_[ee].GetMutableNodes(Position, coord)) {
        ee_[ee].SetVars(Position, coord, node, vars[idx], 0);
        idx++;
    }
    assert(idx == n);
    (void)n;
}

std::pair<int, int> Traj::GetPositionSplineIndex(int ee, double time, int coord) const {   // trajectory.cpp:113-133
    if (coord == 2) throw std::runtime_error("The chosen spline is not mutable and thus does not provide a index.");
    int before = 0;
    for (int e = 0; e < ee; e++) before += 2 * ee_[e].GetTotalPolyVars(Position, coord);
    int into = 0;
    for (int j = 0; j < coord; j++) into += ee_[ee].GetTotalPolyVars(Position, coord);
    const auto vi = ee_[ee].GetVarsIdx(Position, coord, time);
    return {before + into + vi.first, vi.second};
}

std::pair<int, int> Traj::GetForceSplineIndex(int ee, double time, int coord) const {   // trajectory.cpp:363-378
    int before = 0;
    for (int e = 0; e < ee; e++) before += 3 * ee_[e].GetTotalPolyVars(Force, coord);
    int into = 0;
    for (int j = 0; j < coord; j++) into += ee_[ee].GetTotalPolyVars(Force, coord);
    const auto vi = ee_[ee].GetVarsIdx(Force, coord, time);
    return {before + into + vi.first, vi.second};
}

void Traj::AddPolys(double final_time) {   // trajectory.cpp:225-238
    for (auto& s : ee_) {
        while (s.GetEndTime() < final_time) {
            const std::vector<KnotTime> ct = s.GetContactTimes();
            const double last_diff = ct[ct.size() - 1].t - ct[ct.size() - 2].t;
            s.AddPoly(std::max(last_diff, 0.2));
        }
    }
    SetSwingPosZ();
    UpdateSplineVarsCount();
}

void Traj::RemoveUnusedPolys(double init_time) {   // trajectory.cpp:240-246
    for (auto& s : ee_) s.RemovePoly(init_time);
    UpdateSplineVarsCount();
}

std::vector<bool> Traj::GetContacts(double time) const {   // trajectory.cpp:272-286
    std::vector<bool> out(ee_.size());
    for (size_t e = 0; e < ee_.size(); e++) out[e] = (ee_[e].GetVarsIdx(Position, 0, time).second == 1);
    return out;
}

std::vector<std::vector<KnotTime>> Traj::GetContactTimes() const {
    std::vector<std::vector<KnotTime>> out;
    for (const auto& s : ee_) out.push_back(s.GetContactTimes());
    return out;
}

Vec Traj::GetPositionSplineLin(int ee, int coord, double time) const {   // trajectory.cpp:348-361
    if (coord == 2) throw std::runtime_error("You cannot request spline linearizations for position z axis.");
    return ee_[ee].GetPolyVarsLin(Position, coord, time);
}

void Traj::GetForce(int ee, double time, double out[3]) const {
    for (int c = 0; c < 3; c++) out[c] = ee_[ee].ValueAt(Force, c, time);
}

void Traj::GetEndEffectorLocation(int ee, double time, double out[3]) const {
    for (int c = 0; c < 3; c++) out[c] = ee_[ee].ValueAt(Position, c, time);
}

Vec Traj::SplinesAsVec() const {   // trajectory.cpp:429-452
    Vec f, p;
    for (const auto& s : ee_) {
        for (int coord = 0; coord < 3; coord++) {
            const Vec fv = s.GetSplineAsQPVec(Force, coord);
            f.insert(f.end(), fv.begin(), fv.end());
            if (coord < 2) {
                const Vec pv = s.GetSplineAsQPVec(Position, coord);
                p.insert(p.end(), pv.begin(), pv.end());
            }
        }
    }
    assert(static_cast<int>(f.size()) == force_vars_ && static_cast<int>(p.size()) == pos_vars_);
    f.insert(f.end(), p.begin(), p.end());
    return f;
}

void Traj::GetForcePartialWrtContactTime(int ee, double time, int contact_idx, double out[3]) const {
    for (int c = 0; c < 3; c++) out[c] = ee_[ee].ComputePartialWrtTime(Force, c, time, contact_idx);
}

void Traj::GetPositionPartialWrtContactTime(int ee, double time, int contact_idx, double out[3]) const {
    out[2] = 0;
    for (int c = 0; c < 2; c++) out[c] = ee_[ee].ComputePartialWrtTime(Position, c, time, contact_idx);
}

void Traj::UpdateContactTimes(std::vector<std::vector<KnotTime>>& ct) {
    for (size_t e = 0; e < ee_.size(); e++) ee_[e].SetContactTimes(ct.at(e));
}

// ------------------------------------------------------------------------------------------------ quaternion maps
// pinocchio::quaternion::log3 (explog-quaternion.hpp): small-angle Taylor branch below eps^(1/3) on |vec|^2.
void QuatLog3(const double q[4], double out[3]) {
    const double eps = std::numeric_limits<double>::epsilon();
    static const double ts_prec = std::pow(eps, 1.0 / 3.0);
    const double n2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2];
    const double norm = std::sqrt(n2 + eps * eps);
    const double sgn = (q[3] >= 0) ? 1.0 : -1.0;
    const double w = sgn * q[3];
    const double theta_2 = std::atan2(norm, w);
    const double y_x = norm / w;
    const double y_x_sq = n2 / (w * w);
    const double theta = (n2 < ts_prec) ? 2.0 * (1.0 - y_x_sq / 3.0) * y_x : 2.0 * theta_2;
    const double th2_2 = theta * theta / 4.0;
    const double inv_sinc = (n2 < ts_prec) ? 2.0 * (1.0 + th2_2 / 6.0 + 7.0 / 360.0 * th2_2 * th2_2)
                                           : theta / std::sin(theta_2);
    for (int k = 0; k < 3; k++) out[k] = inv_sinc * (sgn * q[k]);
}

// pinocchio::quaternion::exp3: Taylor branch when |v|^2 <= eps^(1/4).
void QuatExp3(const double v[3], double q[4]) {
    const double eps = std::numeric_limits<double>::epsilon();
    static const double ts_prec = std::pow(eps, 1.0 / 4.0);
    const double t2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
    const double t = std::sqrt(t2 + eps * eps);
    if (t2 > ts_prec) {
        const double s = std::sin(t / 2), c = std::cos(t / 2);
        for (int k = 0; k < 3; k++) q[k] = s * (v[k] / t);
        q[3] = c;
    } else {
        const double k_ = 0.5 - t2 / 48.0;
        for (int k = 0; k < 3; k++) q[k] = k_ * v[k];
        q[3] = 1.0 - t2 / 8.0;
    }
}

// pinocchio::quaternion::firstOrderNormalize: q *= (3 - |q|^2) / 2
void QuatFirstOrderNormalize(double q[4]) {
    const double n2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
    const double alpha = (3.0 - n2) / 2.0;
    for (int k = 0; k < 4; k++) q[k] *= alpha;
}

// ------------------------------------------------------------------------------------------------ SrbModel
namespace {
inline void cross(const double a[3], const double b[3], double o[3]) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}
inline void mat3v(const double M[9], const double v[3], double o[3]) {
    for (int i = 0; i < 3; i++) o[i] = M[3 * i] * v[0] + M[3 * i + 1] * v[1] + M[3 * i + 2] * v[2];
}
const double kE[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
}  // namespace

Vec SrbModel::ManifoldToTangent(const Vec& s) const {
    Vec t(12);
    for (int i = 0; i < 6; i++) t[i] = s[i];
    QuatLog3(&s[6], &t[6]);   // reference quaternion hard-wired to identity, single_rigid_body_model.cpp:182
    for (int i = 0; i < 3; i++) t[9 + i] = s[10 + i];
    return t;
}

Vec SrbModel::TangentToManifold(const Vec& s) const {
    Vec m(13);
    for (int i = 0; i < 6; i++) m[i] = s[i];
    QuatExp3(&s[6], &m[6]);
    for (int i = 0; i < 3; i++) m[10 + i] = s[9 + i];
    return m;
}

Vec SrbModel::CalcDynamics(const Vec& x, const Traj& traj, double time) const {   // :222-256
    Vec xdot(12);
    const double* om = &x[9];
    for (int i = 0; i < 3; i++) xdot[i] = x[3 + i] / rc_.mass;
    for (int i = 0; i < 3; i++) xdot[3 + i] = rc_.mass * rc_.gravity[i];
    mat3v(rc_.Ir_inv, om, &xdot[6]);
    double Iw[3], c[3];
    mat3v(rc_.Ir, om, Iw);
    cross(om, Iw, c);
    for (int i = 0; i < 3; i++) xdot[9 + i] = -c[i];
    for (int e = 0; e < traj.NumEE(); e++) {
        double f[3], r[3], rel[3], t[3];
        traj.GetForce(e, time, f);
        traj.GetEndEffectorLocation(e, time, r);
        for (int i = 0; i < 3; i++) xdot[3 + i] += f[i];
        for (int i = 0; i < 3; i++) rel[i] = r[i] - x[i];
        cross(rel, f, t);
        for (int i = 0; i < 3; i++) xdot[9 + i] += t[i];
    }
    return xdot;
}

void SrbModel::GetLinearDynamics(const Vec& state, const Vec& /*ref_state*/, const Traj& traj, double /*dt*/,
                                 double time, Mat& A, Mat& B, Vec& C) const {   // :55-169
    const double* p = &state[0];
    const double* omega = &state[10];
    A = Mat(12, 12);
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) A(i, 3 + j) = kE[i][j] / rc_.mass;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) A(6 + i, 9 + j) = rc_.Ir_inv[3 * i + j];
    double Iw[3];
    mat3v(rc_.Ir, omega, Iw);
    for (int i = 0; i < 3; i++) {
        double c1[3], c2[3];
        const double Icol[3] = {rc_.Ir[i], rc_.Ir[3 + i], rc_.Ir[6 + i]};
        cross(kE[i], Iw, c1);
        cross(omega, Icol, c2);
        for (int r = 0; r < 3; r++) A(9 + r, 9 + i) = -c1[r] - c2[r];
        for (int e = 0; e < traj.NumEE(); e++) {
            double f[3], c3[3];
            traj.GetForce(e, time, f);
            cross(kE[i], f, c3);
            for (int r = 0; r < 3; r++) A(9 + r, i) += -c3[r];
        }
    }
    const int nf = traj.GetTotalForceSplineVars();
    const int nu = nf + traj.GetTotalPosSplineVars();
    B = Mat(12, nu);
    for (int e = 0; e < traj.NumEE(); e++) {
        double r[3], f[3], rel[3];
        traj.GetEndEffectorLocation(e, time, r);
        for (int i = 0; i < 3; i++) rel[i] = r[i] - p[i];
        traj.GetForce(e, time, f);
        for (int coord = 0; coord < 3; coord++) {
            if (traj.IsForceMutable(e, time)) {
                const Vec w = traj.GetForceSplineLin(e, coord, time);
                const auto vi = traj.GetForceSplineIndex(e, time, coord);
                double rc[3];
                cross(rel, kE[coord], rc);
                for (int k = 0; k < vi.second; k++) B(3 + coord, vi.first + k) = w.at(k);
                for (size_t k = 0; k < w.size(); k++)
                    for (int rr = 0; rr < 3; rr++) B(9 + rr, vi.first + static_cast<int>(k)) = rc[rr] * w[k];
            }
            if (coord != 2) {
                const Vec w = traj.GetPositionSplineLin(e, coord, time);
                const auto vi = traj.GetPositionSplineIndex(e, time, coord);
                double ef[3];
                cross(kE[coord], f, ef);
                for (size_t k = 0; k < w.size(); k++)
                    for (int rr = 0; rr < 3; rr++) B(9 + rr, nf + vi.first + static_cast<int>(k)) = ef[rr] * w[k];
            }
        }
    }
    const Vec x = ManifoldToTangent(state);
    const Vec u = traj.SplinesAsVec();
    C.assign(12, 0.0);
    for (int i = 0; i < 12; i++) {
        double s = 0;
        for (int j = 0; j < 12; j++) s += -A(i, j) * x[j];
        double s2 = 0;
        for (int j = 0; j < nu; j++) s2 += B(i, j) * u[j];
        C[i] = s - s2;
    }
    const Vec f0 = CalcDynamics(x, traj, time);
    for (int i = 0; i < 12; i++) C[i] += f0[i];
}

void SrbModel::ComputeLinearizationPartialWrtContactTimes(Mat& dA, Mat& dB, Vec& dC, const Vec& state, const Traj& traj,
                                                          double time, int ee, int contact_idx) const {   // :458-555
    const int nf = traj.GetTotalForceSplineVars();
    const int nu = nf + traj.GetTotalPosSplineVars();
    dA = Mat(12, 12);
    dB = Mat(12, nu);
    dC.assign(12, 0.0);
    double fp[3], pp[3];
    traj.GetForcePartialWrtContactTime(ee, time, contact_idx, fp);
    traj.GetPositionPartialWrtContactTime(ee, time, contact_idx, pp);
    for (int coord = 0; coord < 3; coord++) {
        double c[3];
        cross(kE[coord], fp, c);
        for (int r = 0; r < 3; r++) dA(9 + r, coord) += -c[r];
    }
    double rel[3], r[3], f[3];
    traj.GetEndEffectorLocation(ee, time, r);
    for (int i = 0; i < 3; i++) rel[i] = r[i] - state[i];
    traj.GetForce(ee, time, f);
    for (int coord = 0; coord < 3; coord++) {
        if (traj.IsForceMutable(ee, time)) {
            const Vec dw = traj.Foot(ee).ComputeCoefPartialWrtTime(Force, coord, time, contact_idx, 0);
            const Vec w = traj.GetForceSplineLin(ee, coord, time);
            const auto vi = traj.GetForceSplineIndex(ee, time, coord);
            double rc[3], pc[3];
            cross(rel, kE[coord], rc);
            cross(pp, kE[coord], pc);
            for (int k = 0; k < vi.second; k++) dB(3 + coord, vi.first + k) = dw.at(k);
            for (size_t k = 0; k < dw.size(); k++)
                for (int rr = 0; rr < 3; rr++) dB(9 + rr, vi.first + static_cast<int>(k)) = rc[rr] * dw[k] + pc[rr] * w.at(k);
        }
        if (coord != 2) {
            const Vec dw = traj.Foot(ee).ComputeCoefPartialWrtTime(Position, coord, time, contact_idx, 0);
            const Vec w = traj.GetPositionSplineLin(ee, coord, time);
            const auto vi = traj.GetPositionSplineIndex(ee, time, coord);
            double ef[3], efp[3];
            cross(kE[coord], f, ef);
            cross(kE[coord], fp, efp);
            for (size_t k = 0; k < dw.size(); k++)
                for (int rr = 0; rr < 3; rr++)
                    dB(9 + rr, nf + vi.first + static_cast<int>(k)) = ef[rr] * dw[k] + efp[rr] * w.at(k);
        }
    }
    const Vec x = ManifoldToTangent(state);
    const Vec u = traj.SplinesAsVec();
    for (int i = 0; i < 12; i++) {
        double s = 0;
        for (int j = 0; j < 12; j++) s += -dA(i, j) * x[j];
        double s2 = 0;
        for (int j = 0; j < nu; j++) s2 += dB(i, j) * u[j];
        dC[i] = s - s2;
    }
    for (int i = 0; i < 3; i++) dC[3 + i] += fp[i];
    double c1[3], c2[3];
    cross(rel, fp, c1);
    cross(pp, f, c2);
    for (int i = 0; i < 3; i++) dC[9 + i] += c1[i] + c2[i];
}

// ------------------------------------------------------------------------------------------------ SrbMpc
static std::vector<std::vector<double>> DefaultSwitchingTimes(int num_ee) {   // mpc.cpp:566-588
    return std::vector<std::vector<double>>(num_ee, std::vector<double>{0, 0.3, 0.6, 0.9, 1.2});
}

SrbMpc::SrbMpc(const MpcInfo& info, const RobotConsts& rc, std::shared_ptr<QpSolver> solver)
    : info_(info), model_(rc), solver_(std::move(solver)),
      prev_traj_(info.num_nodes + 1, DefaultSwitchingTimes(4), info.integrator_dt, info.swing_height, info.foot_offset) {
    // mpc.cpp:38-76, mpc_single_rigid_body.cpp:8-23
    data_.constraints = {Dynamics, ForceBox, FrictionCone, EndEffectorLocation, TDPosition, EndEffectorStart};
    num_inputs_ = prev_traj_.GetTotalPosSplineVars() + prev_traj_.GetTotalForceSplineVars();
    const double mu = info_.friction_coef;   // SetFrictionPyramid, mpc.cpp:153-163
    const double h[3] = {1, 0, 0}, l[3] = {0, 1, 0}, n[3] = {0, 0, 1};
    for (int c = 0; c < 3; c++) {
        friction_pyramid_[0][c] = h[c] - n[c] * mu;
        friction_pyramid_[1][c] = -(h[c] + n[c] * mu);
        friction_pyramid_[2][c] = l[c] - n[c] * mu;
        friction_pyramid_[3][c] = -(l[c] + n[c] * mu);
    }
    Phi_ = Mat(12, 12);
    Phi_w_.assign(12, 0.0);
    Q_ = Mat(12, 12);
    w_.assign(12, 0.0);
    // SetInitQPSizes, mpc_single_rigid_body.cpp:323-341
    data_.num_vars = (info_.num_nodes + 1) * 12 + num_inputs_;
    data_.num_dynamics = (info_.num_nodes + 1) * 12;
    data_.num_cone = NumFricConeConstraints();
    data_.num_force_box = NumForceBoxConstraints();
    data_.num_ee_location = 2 * (info_.num_nodes - 3) * 2 * 4;
    data_.num_td_pos = NumTDConstraints();
    data_.num_start_ee = 2 * 4;
    data_.InitQPMats();
    prev_qp_sol_.assign(data_.num_vars, 0.0);
    ee_bounds_[0] = info_.ee_box_size[0];
    ee_bounds_[1] = info_.ee_box_size[1];
}

void SrbMpc::AddQuadraticTrackingCost(const Vec& state_des, const Mat& Q) {
    Q_ = Q;
    w_.assign(12, 0.0);
    for (int i = 0; i < 12; i++) {
        double s = 0;
        for (int j = 0; j < 12; j++) s += (-1 * Q(i, j)) * state_des[j];
        w_[i] = s;
    }
}

void SrbMpc::SetStateTrajectoryWarmStart(const std::vector<Vec>& states) {
    for (int node = 0; node < info_.num_nodes + 1; node++) prev_traj_.SetState(node, states.at(node));
}

void SrbMpc::SetWarmStartTrajectory(const Traj& t) {
    prev_traj_ = t;
    num_inputs_ = prev_traj_.GetTotalPosSplineVars() + prev_traj_.GetTotalForceSplineVars();
    init_time_ = t.GetTime(0);
}

void SrbMpc::AdjustForCurrentContacts(double time, const std::vector<bool>& in_contact) {
    for (int ee = 0; ee < 4; ee++) {
        if (in_contact.at(ee) && !prev_traj_.Foot(ee).IsInContact(time) &&
            std::abs(prev_traj_.GetNextContactTime(ee, time) - time) < 7e-2) {
            prev_traj_.SetEEInContact(ee, time);
        }
    }
}

int SrbMpc::NumForceBoxConstraints() const {   // mpc.cpp:1101-1113
    int n = 0;
    for (const auto& ct : prev_traj_.GetContactTimes())
        for (size_t i = 0; i + 1 < ct.size(); i++)
            if (ct[i].type == TouchDown) n += 2 * 10;
    return n;
}

int SrbMpc::NumFricConeConstraints() const {   // mpc.cpp:1115-1127
    int n = 0;
    for (const auto& ct : prev_traj_.GetContactTimes())
        for (size_t i = 0; i + 1 < ct.size(); i++)
            if (ct[i].type == TouchDown) n += 4 * 10;
    return n;
}

int SrbMpc::NumTDConstraints() const {   // mpc.cpp:1205-1214
    int n = 0;
    for (int ee = 0; ee < 4; ee++) {
        if (prev_traj_.GetNextContactTime(ee, init_time_) - init_time_ < td_fraction_ * prev_traj_.GetCurrentSwingTime(ee)) n += 2;
    }
    return n;
}

void SrbMpc::UpdateQPSizes() {   // mpc.cpp:610-624
    num_inputs_ = prev_traj_.GetTotalPosSplineVars() + prev_traj_.GetTotalForceSplineVars();
    data_.num_vars = (info_.num_nodes + 1) * 12 + num_inputs_;
    data_.num_force_box = NumForceBoxConstraints();
    data_.num_cone = NumFricConeConstraints();
    data_.num_td_pos = NumTDConstraints();
}

void SrbMpc::AddCosts() {   // mpc_single_rigid_body.cpp:61-65 -> mpc.cpp:791-802,542-564,1090-1095
    const int nf = prev_traj_.GetTotalForceSplineVars();
    const int N = info_.num_nodes;
    for (int node = 0; node < N; node++) data_.cost_mat.SetMatrix(Q_, node * 12, node * 12);
    if (nf > 0 && info_.force_cost != 0) data_.cost_mat.SetDiagonalMatrix(info_.force_cost, ForceSplineStartIdx(), ForceSplineStartIdx(), nf);
    for (int node = 0; node < N; node++)
        for (int i = 0; i < 12; i++) data_.cost_linear[node * 12 + i] = w_[i];
    data_.cost_mat.SetMatrix(Phi_, N * 12, N * 12);
    for (int i = 0; i < 12; i++) data_.cost_linear[N * 12 + i] = Phi_w_[i];
    data_.cost_mat.SetDiagonalMatrix(1e-3, 0, 0, data_.num_vars);
}

void SrbMpc::AddDynamicsConstraints(const Vec& state) {   // mpc_single_rigid_body.cpp:218-265
    data_.constraint_mat.SetDiagonalMatrix(-1, 0, 0, 12);
    const Vec x0 = model_.ManifoldToTangent(state);
    for (int i = 0; i < 12; i++) data_.dynamics_constants[i] = -x0[i];
    const double dt = info_.integrator_dt;
    const int N = info_.num_nodes;
    node_A_.assign(N, Mat());
    node_B_.assign(N, Mat());
    node_C_.assign(N, Vec());
    for (int node = 0; node < N; node++) {
        Mat A, B;
        Vec C;
        model_.GetLinearDynamics(prev_traj_.GetState(node), state, prev_traj_, dt, GetTime(node), A, B, C);
        for (int i = 0; i < 12; i++)
            for (int j = 0; j < 12; j++) A(i, j) = (i == j ? 1.0 : 0.0) + dt * A(i, j);
        for (double& b : B.a) b = dt * b;
        for (double& c : C) c = dt * c;
        const int row = constraint_idx_ + (node + 1) * 12;
        data_.constraint_mat.SetMatrix(A, row, node * 12);
        data_.constraint_mat.SetDiagonalMatrix(-1, row, (node + 1) * 12, 12);
        data_.constraint_mat.SetMatrix(B, row, ForceSplineStartIdx());
        for (int i = 0; i < 12; i++) data_.dynamics_constants[row + i] = -C[i];
        node_A_[node] = A;
        node_B_[node] = B;
        node_C_[node] = C;
    }
    constraint_idx_ += (N + 1) * 12;
}

void SrbMpc::AddForceBoxConstraints() {   // mpc.cpp:352-414
    const int fstart = ForceSplineStartIdx();
    int row = 0;
    const auto ct = prev_traj_.GetContactTimes();
    for (int pass = 0; pass < 2; pass++) {
        for (int ee = 0; ee < 4; ee++) {
            for (size_t ti = 0; ti + 1 < ct[ee].size(); ti++) {
                if (ct[ee][ti].type != TouchDown) continue;
                for (int i = 0; i < 10; i++) {
                    const double lower = ct[ee][ti].t, upper = ct[ee][ti + 1].t;
                    const double time = (static_cast<double>(i) / 10.0) * (upper - lower) + lower;
                    if (!prev_traj_.IsForceMutable(ee, time)) throw std::runtime_error("Force is not mutable here.");
                    const auto vi = prev_traj_.GetForceSplineIndex(ee, time, 2);
                    const Vec w = prev_traj_.GetForceSplineLin(ee, 2, time);
                    data_.constraint_mat.SetRow(w, constraint_idx_ + row, fstart + vi.first, pass == 0 ? 1.0 : -1.0);
                    if (pass == 0) {
                        data_.force_box_lb[row] = -0.0;
                        data_.force_box_ub[row] = info_.force_bound;
                    }
                    row++;
                }
            }
        }
    }
    assert(row == data_.num_force_box);
    constraint_idx_ += row;
}

void SrbMpc::AddFrictionConeConstraints() {   // mpc.cpp:166-209
    const int fstart = ForceSplineStartIdx();
    int row = 0;
    const auto ct = prev_traj_.GetContactTimes();
    for (int ee = 0; ee < 4; ee++) {
        for (size_t ti = 0; ti + 1 < ct[ee].size(); ti++) {
            if (ct[ee][ti].type != TouchDown) continue;
            for (int i = 0; i < 10; i++) {
                for (int coord = 0; coord < 3; coord++) {
                    const double lower = ct[ee][ti].t, upper = ct[ee][ti + 1].t;
                    const double time = (static_cast<double>(i) / 10.0) * (upper - lower) + lower;
                    const auto vi = prev_traj_.GetForceSplineIndex(ee, time, coord);
                    const Vec w = prev_traj_.GetForceSplineLin(ee, coord, time);
                    for (int fc = 0; fc < 4; fc++) {
                        data_.constraint_mat.SetRow(w, constraint_idx_ + row + fc, fstart + vi.first, friction_pyramid_[fc][coord]);
                        data_.friction_cone_ub[row + fc] = 0;
                    }
                }
                row += 4;
            }
        }
    }
    assert(row == data_.num_cone);
    constraint_idx_ += row;
}

void SrbMpc::AddEELocationConstraints() {   // mpc_single_rigid_body.cpp:381-443
    const int pstart = PosSplineStartIdx();
    const double bounds[2] = {info_.ee_box_size[0] / 2, info_.ee_box_size[1] / 2};
    Mat A(data_.num_ee_location, data_.num_vars);
    int idx = 0;
    for (int pass = 0; pass < 2; pass++) {
        for (int node = 4; node < info_.num_nodes + 1; node++) {
            for (int ee = 0; ee < 4; ee++) {
                if (pass == 0) {
                    for (int c = 0; c < 2; c++) {
                        data_.ee_location_ub[idx + c] = bounds[c] + model_.Consts().hip_xy[ee][c];
                        data_.ee_location_lb[idx + c] = -bounds[c] + model_.Consts().hip_xy[ee][c];
                    }
                }
                for (int coord = 0; coord < 2; coord++) {
                    A(idx, node * 12 + coord) = (pass == 0) ? -1 : 1;
                    const auto vi = prev_traj_.GetPositionSplineIndex(ee, GetTime(node), coord);
                    const Vec w = prev_traj_.GetPositionSplineLin(ee, coord, GetTime(node));
                    for (int k = 0; k < vi.second; k++) A(idx, pstart + vi.first + k) = (pass == 0) ? w.at(k) : -w.at(k);
                    idx++;
                }
            }
        }
    }
    data_.constraint_mat.SetMatrix(A, constraint_idx_, 0);
    assert(idx == data_.num_ee_location);
    constraint_idx_ += idx;
}

void SrbMpc::AddTDPositionConstraints() {   // mpc_single_rigid_body.cpp:849-887
    const int pstart = PosSplineStartIdx();
    int row = 0;
    for (int ee = 0; ee < 4; ee++) {
        if (prev_traj_.GetNextContactTime(ee, init_time_) - init_time_ < td_fraction_ * prev_traj_.GetCurrentSwingTime(ee)) {
            const double td_time = prev_traj_.GetNextContactTime(ee, init_time_);
            double loc[3];
            prev_traj_.GetEndEffectorLocation(ee, td_time, loc);
            data_.td_pos_constants[row] = loc[0];
            data_.td_pos_constants[row + 1] = loc[1];
            for (int coord = 0; coord < 2; coord++) {
                const auto vi = prev_traj_.GetPositionSplineIndex(ee, td_time, coord);
                const Vec w = prev_traj_.GetPositionSplineLin(ee, coord, td_time);
                data_.constraint_mat.SetRow(w, constraint_idx_ + row, pstart + vi.first);
                row++;
            }
        }
    }
    assert(row == data_.num_td_pos);
    constraint_idx_ += row;
}

void SrbMpc::AddEEStartConstraints(const std::vector<std::array<double, 3>>& ee_start) {   // :445-475
    int idx = 0;
    Mat M(data_.num_start_ee, prev_traj_.GetTotalPosSplineVars());
    for (int ee = 0; ee < 4; ee++) {
        data_.start_ee_constants[idx] = ee_start.at(ee)[0];
        data_.start_ee_constants[idx + 1] = ee_start.at(ee)[1];
        for (int coord = 0; coord < 2; coord++) {
            const auto vi = prev_traj_.GetPositionSplineIndex(ee, GetTime(0), coord);
            const Vec w = prev_traj_.GetPositionSplineLin(ee, coord, GetTime(0));
            for (int k = 0; k < vi.second; k++) M(idx, vi.first + k) = w.at(k);
            idx++;
        }
    }
    data_.constraint_mat.SetMatrix(M, constraint_idx_, PosSplineStartIdx());
    constraint_idx_ += idx;
}

Vec SrbMpc::ConvertTrajToQPVec(const Traj& traj) const {   // mpc_single_rigid_body.cpp:343-357
    Vec q(traj.GetTotalVariables(), 0.0);
    for (int i = 0; i < info_.num_nodes + 1; i++) {
        const Vec t = model_.ManifoldToTangent(traj.GetState(i));
        std::copy(t.begin(), t.end(), q.begin() + i * 12);
    }
    const Vec u = traj.SplinesAsVec();
    std::copy(u.begin(), u.end(), q.end() - u.size());
    return q;
}

Traj SrbMpc::ConvertQPSolToTrajectory(const Vec& z) const {   // mpc_single_rigid_body.cpp:275-321
    Traj traj(prev_traj_);
    int fi = ForceSplineStartIdx();
    int pi = PosSplineStartIdx();
    for (int ee = 0; ee < 4; ee++) {
        for (int coord = 0; coord < 3; coord++) {
            int nv = traj.Foot(ee).GetTotalPolyVars(Force, coord);
            traj.UpdateForceSpline(ee, coord, &z[fi], nv);
            fi += nv;
            if (coord < 2) {
                nv = traj.Foot(ee).GetTotalPolyVars(Position, coord);
                traj.UpdatePositionSpline(ee, coord, z.data() + pi, nv);
                pi += nv;
            }
        }
    }
    assert(pi == static_cast<int>(z.size()));
    for (int node = 0; node < info_.num_nodes + 1; node++) {
        Vec t(z.begin() + node * 12, z.begin() + node * 12 + 12);
        Vec m = model_.TangentToManifold(t);
        QuatFirstOrderNormalize(&m[6]);
        traj.SetState(node, m);
    }
    return traj;
}

double SrbMpc::GetCostValue(const Vec& x) const {   // mpc.cpp:759-761
    Vec Px(x.size());
    data_.P.mul(x.data(), Px.data());
    double a = 0, b = 0;
    for (size_t i = 0; i < x.size(); i++) {
        a += x[i] * Px[i];
        b += data_.cost_linear[i] * x[i];
    }
    return 0.5 * a + b;
}

Vec SrbMpc::GetEqualityConstraintValues(const Traj& traj) const {   // mpc.cpp:764-776, rk_integrator.cpp:14-30
    const int N = info_.num_nodes;
    Vec d(static_cast<size_t>(N) * 12, 0.0);
    for (int node = 0; node < N; node++) {
        const Vec xn = model_.ManifoldToTangent(traj.GetState(node + 1));
        const Vec x = model_.ManifoldToTangent(traj.GetState(node));
        const Vec f = model_.CalcDynamics(x, traj, GetTime(node));
        for (int i = 0; i < 12; i++) d[node * 12 + i] = xn[i] - (x[i] + info_.integrator_dt * f[i]);
    }
    return d;
}

static double L1(const Vec& v) {
    double s = 0;
    for (double x : v) s += std::abs(x);
    return s;
}

double SrbMpc::GetMeritValue(const Vec& x) const {   // mpc.cpp:749-753
    const Traj t = ConvertQPSolToTrajectory(x);
    return mu_ * L1(GetEqualityConstraintValues(t)) + GetCostValue(x);
}

double SrbMpc::GetMeritGradient(const Vec& x, const Vec& p) const {   // mpc.cpp:783-788
    const Traj t = ConvertQPSolToTrajectory(x);
    Vec g(x.size());
    data_.P.mul(x.data(), g.data());
    double s = 0;
    for (size_t i = 0; i < x.size(); i++) s += (g[i] + data_.cost_linear[i]) * p[i];
    return s - mu_ * L1(GetEqualityConstraintValues(t));
}

double SrbMpc::LineSearch(const Vec& direction) const {   // mpc.cpp:730-747
    double alpha = 1;
    const double merit = GetMeritValue(prev_qp_sol_);
    Vec tmp(prev_qp_sol_.size());
    for (size_t i = 0; i < tmp.size(); i++) tmp[i] = alpha * direction[i] + prev_qp_sol_[i];
    double merit_step = GetMeritValue(tmp);
    const double dd = GetMeritGradient(prev_qp_sol_, direction);
    int i = 0;
    while ((merit - merit_step) < -0.00001 * alpha * dd && i < 10) {
        alpha *= 0.5;
        for (size_t k = 0; k < tmp.size(); k++) tmp[k] = (alpha * direction[k]) + prev_qp_sol_[k];
        merit_step = GetMeritValue(tmp);
        i++;
    }
    return alpha;
}

void SrbMpc::Prepare(const Vec& state, double init_time, const std::vector<std::array<double, 3>>& ee_start) {
    // mpc_single_rigid_body.cpp:25-107
    init_time_ = init_time;
    prev_traj_.SetInitTime(init_time);
    prev_traj_.AddPolys(info_.integrator_dt * info_.num_nodes + init_time);
    prev_traj_.RemoveUnusedPolys(init_time);
    UpdateQPSizes();
    data_.InitQPMats();
    prev_traj_.SetState(0, state);
    prev_qp_sol_ = ConvertTrajToQPVec(prev_traj_);
    assert(static_cast<int>(prev_qp_sol_.size()) == data_.num_vars);
    AddCosts();
    constraint_idx_ = 0;
    for (Constraint c : data_.constraints) {
        switch (c) {
            case Dynamics: AddDynamicsConstraints(prev_traj_.GetState(0)); break;
            case ForceBox: AddForceBoxConstraints(); break;
            case FrictionCone: AddFrictionConeConstraints(); break;
            case EndEffectorLocation: AddEELocationConstraints(); break;
            case TDPosition: AddTDPositionConstraints(); break;
            case EndEffectorStart: AddEEStartConstraints(ee_start); break;
        }
    }
    data_.ConstructSparseMats();
    data_.ConstructVectors();
}

void SrbMpc::AssembleOnly(const Vec& state, double init_time, const std::vector<std::array<double, 3>>& ee_start) {
    Prepare(state, init_time, ee_start);
}

const Traj& SrbMpc::Solve(const Vec& state, double init_time, const std::vector<std::array<double, 3>>& ee_start) {
    Prepare(state, init_time, ee_start);
    last_qp_ = solver_->Solve(data_, prev_qp_sol_, in_real_time_);   // :110-129
    Vec sol = last_qp_.x;
    if (last_qp_.status == PrimalInfeasible || last_qp_.no_iterate) sol = prev_qp_sol_;      // the "Primal infeasible." throw/catch
    if (last_qp_.status != SolvedInacc && last_qp_.status != Solved && last_qp_.status != MaxIter) {   // :136-144
        info_.ee_box_size[0] += 0.05;
        info_.ee_box_size[1] += 0.05;
    } else {
        info_.ee_box_size[0] = std::max(info_.ee_box_size[0] - 0.05, ee_bounds_[0]);
        info_.ee_box_size[1] = std::max(info_.ee_box_size[1] - 0.05, ee_bounds_[1]);
    }
    Vec p(sol.size());
    for (size_t i = 0; i < p.size(); i++) p[i] = sol[i] - prev_qp_sol_[i];
    double alpha = 1;
    if (sol.size() == prev_qp_sol_.size()) alpha = LineSearch(p);
    for (size_t i = 0; i < p.size(); i++) prev_qp_sol_[i] = (alpha * p[i]) + prev_qp_sol_[i];
    prev_traj_ = ConvertQPSolToTrajectory(prev_qp_sol_);
    // RecordStats, mpc.cpp:804-816
    stats_.alpha = alpha;
    stats_.eq_violation = L1(GetEqualityConstraintValues(prev_traj_));
    double sn = 0;
    for (double x : p) sn += x * x;
    stats_.step_norm = std::sqrt(sn);
    stats_.cost = GetCostValue(prev_qp_sol_);
    stats_.merit = mu_ * stats_.eq_violation + GetCostValue(ConvertTrajToQPVec(prev_traj_));
    Vec back(prev_qp_sol_);
    for (size_t i = 0; i < back.size(); i++) back[i] -= alpha * p[i];
    stats_.merit_dd = GetMeritGradient(back, p);
    stats_.status = last_qp_.status;
    stats_.qp_iters = last_qp_.iters;
    return prev_traj_;
}

const Traj& SrbMpc::CreateInitialRun(const Vec& state, const std::vector<std::array<double, 3>>& ee_start) {
    in_real_time_ = false;
    for (int i = 0; i < 10; i++) Solve(state, 0, ee_start);   // `converged` is never set, mpc.cpp:78-90
    return prev_traj_;
}

const Traj& SrbMpc::GetRealTimeUpdate(const Vec& state, double init_time, const std::vector<std::array<double, 3>>& ee_start) {
    in_real_time_ = true;   // mpc.cpp:92-108
    return Solve(state, init_time, ee_start);
}

}  // namespace oracle
