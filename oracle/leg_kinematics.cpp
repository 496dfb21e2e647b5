// TEST INFRASTRUCTURE ONLY -- see leg_kinematics.hpp.
#include "leg_kinematics.hpp"

#include "srb_mpc.hpp"

#include <cmath>
#include <cstring>

namespace oracle {
namespace kin {
namespace {

// pinocchio/math/taylor-expansion.hpp: TaylorSeriesExpansion<double>::precision<3>() = eps^(1/4)
const double kTaylor3 = std::sqrt(std::sqrt(2.220446049250313e-16));
const double kPi = 3.14159265358979323846;

void Cross(const double a[3], const double b[3], double o[3]) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}
double Dot3(const double a[3], const double b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
void MatVec(const double R[9], const double v[3], double o[3]) {
    for (int i = 0; i < 3; i++) o[i] = R[3 * i] * v[0] + R[3 * i + 1] * v[1] + R[3 * i + 2] * v[2];
}
void MatTVec(const double R[9], const double v[3], double o[3]) {
    for (int i = 0; i < 3; i++) o[i] = R[i] * v[0] + R[3 + i] * v[1] + R[6 + i] * v[2];
}
void MatMul(const double A[9], const double B[9], double C[9]) {
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
void MatTMul(const double A[9], const double B[9], double C[9]) {   // A' B
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) C[3 * i + j] = A[i] * B[j] + A[3 + i] * B[3 + j] + A[6 + i] * B[6 + j];
}
// rotation by angle about a unit axis: what pinocchio's revolute joints (RX, RY, RZ, unaligned) evaluate
void AxisAngle(const double a[3], double ang, double R[9]) {
    const double c = std::cos(ang), s = std::sin(ang), v = 1.0 - c;
    R[0] = a[0] * a[0] * v + c;        R[1] = a[0] * a[1] * v - a[2] * s; R[2] = a[0] * a[2] * v + a[1] * s;
    R[3] = a[1] * a[0] * v + a[2] * s; R[4] = a[1] * a[1] * v + c;        R[5] = a[1] * a[2] * v - a[0] * s;
    R[6] = a[2] * a[0] * v - a[1] * s; R[7] = a[2] * a[1] * v + a[0] * s; R[8] = a[2] * a[2] * v + c;
}

}  // namespace

void QuatToMatrix(const double q[4], double R[9]) {
    const double x = q[0], y = q[1], z = q[2], w = q[3];
    const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
    const double twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
    R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
    R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}

void MatrixToQuat(const double m[9], double q[4]) {
    double t = m[0] + m[4] + m[8];
    if (t > 0) {
        t = std::sqrt(t + 1.0);
        q[3] = 0.5 * t;
        t = 0.5 / t;
        q[0] = (m[7] - m[5]) * t;
        q[1] = (m[2] - m[6]) * t;
        q[2] = (m[3] - m[1]) * t;
    } else {
        int i = 0;
        if (m[4] > m[0]) i = 1;
        if (m[8] > m[4 * i]) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = std::sqrt(m[4 * i] - m[4 * j] - m[4 * k] + 1.0);
        q[i] = 0.5 * t;
        t = 0.5 / t;
        q[3] = (m[3 * k + j] - m[3 * j + k]) * t;
        q[j] = (m[3 * j + i] + m[3 * i + j]) * t;
        q[k] = (m[3 * k + i] + m[3 * i + k]) * t;
    }
}

Se3 Inverse(const Se3& a) {
    Se3 o;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) o.R[3 * i + j] = a.R[3 * j + i];
    double t[3];
    MatVec(o.R, a.p, t);
    for (int i = 0; i < 3; i++) o.p[i] = -t[i];
    return o;
}
Se3 Mul(const Se3& a, const Se3& b) {
    Se3 o;
    MatMul(a.R, b.R, o.R);
    double t[3];
    MatVec(a.R, b.p, t);
    for (int i = 0; i < 3; i++) o.p[i] = a.p[i] + t[i];
    return o;
}
Se3 ActInv(const Se3& a, const Se3& b) {
    Se3 o;
    MatTMul(a.R, b.R, o.R);
    const double d[3] = {b.p[0] - a.p[0], b.p[1] - a.p[1], b.p[2] - a.p[2]};
    MatTVec(a.R, d, o.p);
    return o;
}

// pinocchio/spatial/explog.hpp: exp6(MotionDense)
void Exp6(const double nu[6], Se3& M) {
    const double* v = nu;
    const double* w = nu + 3;
    const double t2 = Dot3(w, w), t = std::sqrt(t2), wv = Dot3(w, v);
    double alpha_wxv, alpha_v, alpha_w, diag;
    if (t > kTaylor3) {
        const double ct = std::cos(t), st = std::sin(t), inv_t2 = 1.0 / t2;
        alpha_wxv = (1.0 - ct) * inv_t2;
        alpha_v = st / t;
        alpha_w = (1.0 - alpha_v) * inv_t2 * wv;
        diag = ct;
    } else {
        alpha_wxv = 0.5 - t2 / 24.0;
        alpha_v = 1.0 - t2 / 6.0;
        alpha_w = (1.0 / 6.0 - t2 / 120.0) * wv;
        diag = 1.0 - t2 / 2.0;
    }
    double wxv[3];
    Cross(w, v, wxv);
    for (int i = 0; i < 3; i++) M.p[i] = alpha_v * v[i] + alpha_w * w[i] + alpha_wxv * wxv[i];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) M.R[3 * i + j] = alpha_wxv * w[i] * w[j];
    M.R[1] -= alpha_v * w[2]; M.R[3] += alpha_v * w[2];
    M.R[2] += alpha_v * w[1]; M.R[6] -= alpha_v * w[1];
    M.R[5] -= alpha_v * w[0]; M.R[7] += alpha_v * w[0];
    M.R[0] += diag; M.R[4] += diag; M.R[8] += diag;
}

// pinocchio/spatial/log.hxx: log3_impl
void Log3(const double R[9], double w[3], double& theta) {
    const double tr = R[0] + R[4] + R[8];
    if (tr >= 3.0) theta = 0.0;
    else if (tr <= -1.0) theta = kPi;
    else theta = std::acos((tr - 1.0) / 2.0);
    if (theta >= kPi - 1e-2) {   // near pi: axis from the diagonal, signs from the skew part
        const double cphi = -(tr - 1.0) / 2.0, beta = theta * theta / (1.0 + cphi);
        const double tmp[3] = {(R[0] + cphi) * beta, (R[4] + cphi) * beta, (R[8] + cphi) * beta};
        w[0] = (R[7] > R[5] ? 1.0 : -1.0) * (tmp[0] > 0 ? std::sqrt(tmp[0]) : 0.0);
        w[1] = (R[2] > R[6] ? 1.0 : -1.0) * (tmp[1] > 0 ? std::sqrt(tmp[1]) : 0.0);
        w[2] = (R[3] > R[1] ? 1.0 : -1.0) * (tmp[2] > 0 ? std::sqrt(tmp[2]) : 0.0);
    } else {
        const double t = ((theta > kTaylor3) ? theta / std::sin(theta) : 1.0) / 2.0;
        w[0] = t * (R[7] - R[5]);
        w[1] = t * (R[2] - R[6]);
        w[2] = t * (R[3] - R[1]);
    }
}

// pinocchio/spatial/log.hxx: Jlog3_impl
void Jlog3(double theta, const double w[3], double J[9]) {
    double alpha, diag;
    if (theta < kTaylor3) {
        alpha = 1.0 / 12.0 + theta * theta / 720.0;
        diag = 0.5 * (2.0 - theta * theta / 6.0);
    } else {
        const double ct = std::cos(theta), st = std::sin(theta), st_1mct = st / (1.0 - ct);
        alpha = 1.0 / (theta * theta) - st_1mct / (2.0 * theta);
        diag = 0.5 * (theta * st_1mct);
    }
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) J[3 * i + j] = alpha * w[i] * w[j];
    J[0] += diag; J[4] += diag; J[8] += diag;
    // addSkew(0.5 w, J)
    J[1] -= 0.5 * w[2]; J[3] += 0.5 * w[2];
    J[2] += 0.5 * w[1]; J[6] -= 0.5 * w[1];
    J[5] -= 0.5 * w[0]; J[7] += 0.5 * w[0];
}

// pinocchio/spatial/log.hxx: log6_impl
void Log6(const Se3& M, double out[6]) {
    double w[3], t;
    Log3(M.R, w, t);
    const double t2 = t * t;
    double alpha, beta;
    if (t < kTaylor3) {
        alpha = 1.0 - t2 / 12.0 - t2 * t2 / 720.0;
        beta = 1.0 / 12.0 + t2 / 720.0;
    } else {
        const double st = std::sin(t), ct = std::cos(t);
        alpha = t * st / (2.0 * (1.0 - ct));
        beta = 1.0 / t2 - st / (2.0 * t * (1.0 - ct));
    }
    double wxp[3];
    Cross(w, M.p, wxp);
    const double wp = Dot3(w, M.p);
    for (int i = 0; i < 3; i++) {
        out[i] = alpha * M.p[i] - 0.5 * wxp[i] + (beta * wp) * w[i];
        out[3 + i] = w[i];
    }
}

// pinocchio/spatial/log.hxx: Jlog6_impl -- [[A, B], [0, A]], A = Jlog3(R), B = C A
void Jlog6(const Se3& M, double J[36]) {
    double w[3], t, A[9];
    Log3(M.R, w, t);
    Jlog3(t, w, A);
    const double t2 = t * t;
    double beta, beta_dot_over_theta;
    if (t < kTaylor3) {
        beta = 1.0 / 12.0 + t2 / 720.0;
        beta_dot_over_theta = 1.0 / 360.0;
    } else {
        const double tinv = 1.0 / t, t2inv = tinv * tinv, st = std::sin(t), ct = std::cos(t), inv_2_2ct = 1.0 / (2.0 * (1.0 - ct));
        beta = t2inv - st * tinv * inv_2_2ct;
        beta_dot_over_theta = -2.0 * t2inv * t2inv + (1.0 + st * tinv) * t2inv * inv_2_2ct;
    }
    const double* p = M.p;
    const double wTp = Dot3(w, p);
    double v3[3], C[9], B[9];
    for (int i = 0; i < 3; i++) v3[i] = (beta_dot_over_theta * wTp) * w[i] - (t2 * beta_dot_over_theta + 2.0 * beta) * p[i];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) C[3 * i + j] = v3[i] * w[j] + beta * w[i] * p[j];
    C[0] += wTp * beta; C[4] += wTp * beta; C[8] += wTp * beta;
    C[1] -= 0.5 * p[2]; C[3] += 0.5 * p[2];
    C[2] += 0.5 * p[1]; C[6] -= 0.5 * p[1];
    C[5] -= 0.5 * p[0]; C[7] += 0.5 * p[0];
    MatMul(C, A, B);
    for (int i = 0; i < 36; i++) J[i] = 0.0;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            J[6 * i + j] = A[3 * i + j];
            J[6 * i + 3 + j] = B[3 * i + j];
            J[6 * (3 + i) + 3 + j] = A[3 * i + j];
        }
}

void ForwardKinematics(const RobotKin& rk, const double q[kNq], Se3 joints[13], Se3 feet[4]) {
    QuatToMatrix(q + 3, joints[0].R);
    for (int i = 0; i < 3; i++) joints[0].p[i] = q[i];
    for (int ee = 0; ee < 4; ee++) {
        const LegChain& lc = rk.leg[ee];
        Se3 parent = joints[0];
        for (int j = 0; j < 3; j++) {
            Se3 place, rot;
            std::memcpy(place.R, lc.R[j], sizeof place.R);
            std::memcpy(place.p, lc.t[j], sizeof place.p);
            AxisAngle(lc.axis[j], q[7 + 3 * ee + j], rot.R);
            rot.p[0] = rot.p[1] = rot.p[2] = 0.0;
            parent = Mul(parent, Mul(place, rot));   // oMi = oMi[parent] * liMi,  liMi = jointPlacement * joint transform
            joints[1 + 3 * ee + j] = parent;
        }
        Se3 fp;
        std::memcpy(fp.R, lc.R[3], sizeof fp.R);
        std::memcpy(fp.p, lc.t[3], sizeof fp.p);
        feet[ee] = Mul(parent, fp);   // updateFramePlacements: oMf = oMi[parent joint] * frame placement
    }
}

void FootJacobianLocal(const RobotKin& rk, const Se3 joints[13], const Se3 feet[4], int ee, double J[6 * kNv]) {
    for (int i = 0; i < 6 * kNv; i++) J[i] = 0.0;
    // a joint's motion subspace S, expressed in the foot frame: (jMf)^-1 acting on S, jMf = oMi[j]^-1 oMf
    const Se3 bMf = ActInv(joints[0], feet[ee]);
    for (int k = 0; k < 3; k++) {   // free flyer: S = I6 in the base frame
        double e[3] = {0, 0, 0}, exp_[3], lin[3], ang[3];
        e[k] = 1.0;
        MatTVec(bMf.R, e, lin);   // linear column k: R'(v - p x w), w = 0
        for (int i = 0; i < 3; i++) J[kNv * i + k] = lin[i];
        Cross(e, bMf.p, exp_);    // angular column k: linear part R'(w x p), angular part R' w
        MatTVec(bMf.R, exp_, lin);
        MatTVec(bMf.R, e, ang);
        for (int i = 0; i < 3; i++) {
            J[kNv * i + 3 + k] = lin[i];
            J[kNv * (3 + i) + 3 + k] = ang[i];
        }
    }
    for (int j = 0; j < 3; j++) {
        const Se3 jMf = ActInv(joints[1 + 3 * ee + j], feet[ee]);
        const double* a = rk.leg[ee].axis[j];
        double axp[3], lin[3], ang[3];
        Cross(a, jMf.p, axp);
        MatTVec(jMf.R, axp, lin);
        MatTVec(jMf.R, a, ang);
        const int col = 6 + 3 * ee + j;
        for (int i = 0; i < 3; i++) {
            J[kNv * i + col] = lin[i];
            J[kNv * (3 + i) + col] = ang[i];
        }
    }
}

// pinocchio/multibody/liegroup/special-euclidean.hpp: SpecialEuclideanOperationTpl<3>::integrate_impl; revolute joints add
void Integrate(const double q[kNq], const double v[kNv], double out[kNq]) {
    Se3 M0, E;
    QuatToMatrix(q + 3, M0.R);
    for (int i = 0; i < 3; i++) M0.p[i] = q[i];
    Exp6(v, E);
    const Se3 M1 = Mul(M0, E);
    for (int i = 0; i < 3; i++) out[i] = M1.p[i];
    double rq[4];
    MatrixToQuat(M1.R, rq);
    const double dot = rq[0] * q[3] + rq[1] * q[4] + rq[2] * q[5] + rq[3] * q[6];
    if (dot < 0)
        for (double& c : rq) c = -c;
    const double n2 = rq[0] * rq[0] + rq[1] * rq[1] + rq[2] * rq[2] + rq[3] * rq[3];   // quaternion::firstOrderNormalize
    const double a = (3.0 - n2) / 2.0;
    for (int i = 0; i < 4; i++) out[3 + i] = rq[i] * a;
    for (int i = 0; i < 12; i++) out[7 + i] = q[7 + i] + v[6 + i];
}

// single_rigid_body_model.cpp:314-425
int InverseKinematics(const RobotKin& rk, const double state[13], const double ee_des[4][3], const double joint_guess[12], double q[kNq],
                      int iters[4]) {
    const double eps = 5e-6, DT = 1e-1, damp = 1e-6;   // :345-348
    const int IT_MAX = 1000;
    Se3 body_des;                                      // :336-337 (the interpolated quaternion is used as it comes)
    QuatToMatrix(state + 6, body_des.R);
    for (int i = 0; i < 3; i++) body_des.p[i] = state[i];
    for (int i = 0; i < 3; i++) q[i] = state[i];        // :339-342
    for (int i = 0; i < 4; i++) q[3 + i] = state[6 + i];
    for (int i = 0; i < 12; i++) q[7 + i] = joint_guess[i];
    bool success = false;                              // :363 -- set once, never cleared between feet
    for (int ee = 0; ee < 4; ee++) {
        iters[ee] = IT_MAX;
        for (int it = 0; it < IT_MAX; it++) {
            const double n2 = q[3] * q[3] + q[4] * q[4] + q[5] * q[5] + q[6] * q[6];   // :372-377 firstOrderNormalize
            const double a = (3.0 - n2) / 2.0;
            for (int i = 0; i < 4; i++) q[3 + i] *= a;
            Se3 joints[13], feet[4];
            ForwardKinematics(rk, q, joints, feet);
            double err[9];
            const double d[3] = {ee_des[ee][0] - feet[ee].p[0], ee_des[ee][1] - feet[ee].p[1], ee_des[ee][2] - feet[ee].p[2]};
            MatTVec(feet[ee].R, d, err);               // :382-384 oMf.actInv(SE3(I, target)).translation()
            const Se3 body_err = ActInv(joints[0], body_des);   // :386-388
            Log6(body_err, err + 3);
            double nrm = 0;
            for (double e : err) nrm += e * e;
            if (std::sqrt(nrm) < eps) {                // :390-393
                success = true;
                iters[ee] = it;
                break;
            }
            double Jf[6 * kNv], J[9 * kNv], Jl[36];
            FootJacobianLocal(rk, joints, feet, ee, Jf);
            for (int i = 0; i < 3 * kNv; i++) J[i] = -Jf[i];   // :395-396
            Jlog6(Inverse(body_err), Jl);              // :398, ComputeJacobianForIK: J = -Jlog6(err^-1) * [I6 0]
            for (int i = 0; i < 6; i++)
                for (int j = 0; j < kNv; j++) J[kNv * (3 + i) + j] = (j < 6) ? -Jl[6 * i + j] : 0.0;
            double A[81];                              // :402-404
            for (int i = 0; i < 9; i++)
                for (int j = 0; j < 9; j++) {
                    double s = 0;
                    for (int k = 0; k < kNv; k++) s += J[kNv * i + k] * J[kNv * j + k];
                    A[9 * i + j] = s + (i == j ? damp : 0.0);
                }
            // JJt.ldlt().solve(err): L D L' without pivoting (the matrix is positive definite)
            double y[9];
            for (int j = 0; j < 9; j++) {
                double dj = A[9 * j + j];
                for (int k = 0; k < j; k++) dj -= A[9 * j + k] * A[9 * j + k] * A[9 * k + k];
                A[9 * j + j] = dj;
                for (int i = j + 1; i < 9; i++) {
                    double s = A[9 * i + j];
                    for (int k = 0; k < j; k++) s -= A[9 * i + k] * A[9 * j + k] * A[9 * k + k];
                    A[9 * i + j] = s / dj;
                }
            }
            for (int i = 0; i < 9; i++) {
                double s = err[i];
                for (int k = 0; k < i; k++) s -= A[9 * i + k] * y[k];
                y[i] = s;
            }
            for (int i = 0; i < 9; i++) y[i] /= A[9 * i + i];
            for (int i = 8; i >= 0; i--) {
                double s = y[i];
                for (int k = i + 1; k < 9; k++) s -= A[9 * k + i] * y[k];
                y[i] = s;
            }
            double v[kNv], qn[kNq];                    // :405-406
            for (int j = 0; j < kNv; j++) {
                double s = 0;
                for (int i = 0; i < 9; i++) s += J[kNv * i + j] * y[i];
                v[j] = -s * DT;
            }
            Integrate(q, v, qn);
            std::memcpy(q, qn, sizeof qn);
        }
        if (!success) return 1;                        // :417-420
    }
    return 0;
}

// controllers/mpc_controller.cpp:414-511
int GetTargetsFromTraj(const RobotKin& rk, const Traj& traj, double time, double dt, double mass, const double Ir_inv[9], double q_des[kNq],
                       double v_des[kNv], double force_des[12]) {
    if (time < traj.GetTime(0)) time = traj.GetTime(0);                       // :415-417
    const int node = static_cast<int>(std::ceil((time - traj.InitTime()) / traj.NodeDt()));   // :420, Trajectory::GetNode (trajectory.cpp:479-481)
    if (node + 1 >= traj.NumStates() || node < 0) return 3;
    auto lerp = [&](int a, int b, double t_b, double t_a, double at, double out[13]) {
        // (x_b - x_a) * (1 - (t_b - at) / (t_b - t_a)) + x_a
        const Vec& xa = traj.GetState(a);
        const Vec& xb = traj.GetState(b);
        const double w = 1 - (t_b - at) / (t_b - t_a);
        for (int i = 0; i < 13; i++) out[i] = (xb[i] - xa[i]) * w + xa[i];
    };
    double s1[13], s2[13];
    if (node > 0) {                                                            // :431-447
        lerp(node - 1, node, traj.GetTime(node), traj.GetTime(node - 1), time, s1);
        if (time + dt < traj.GetTime(node)) return 2;
        lerp(node, node + 1, traj.GetTime(node + 1), traj.GetTime(node), time + dt, s2);
    } else {                                                                   // :448-456
        lerp(node, node + 1, traj.GetTime(node + 1), traj.GetTime(node), time, s1);
        lerp(node, node + 1, traj.GetTime(node + 1), traj.GetTime(node), time + dt, s2);
    }
    double ee1[4][3], ee2[4][3];
    for (int ee = 0; ee < 4; ee++) {                                           // :460-464, :477-481 / :491-495
        traj.GetEndEffectorLocation(ee, time, ee1[ee]);
        traj.GetEndEffectorLocation(ee, time + dt, ee2[ee]);
    }
    int iters[4];
    double q1[kNq], q2[kNq];
    if (InverseKinematics(rk, s1, ee1, q_des + 7, q1, iters)) return 1;        // :466-468
    for (int i = 0; i < 3; i++) {                                              // :471-473
        v_des[i] = s1[3 + i] / mass;
        v_des[3 + i] = Ir_inv[3 * i] * s1[10] + Ir_inv[3 * i + 1] * s1[11] + Ir_inv[3 * i + 2] * s1[12];
    }
    if (InverseKinematics(rk, s2, ee2, q1 + 7, q2, iters)) return 1;           // :483-485 / :497-499 (guess = the new q_des_)
    for (int i = 0; i < 12; i++) v_des[6 + i] = (q2[7 + i] - q1[7 + i]) / dt;  // :487 and :501 are the same difference
    for (int i = 0; i < kNq; i++) q_des[i] = q1[i];
    for (int ee = 0; ee < 4; ee++) traj.GetForce(ee, time, force_des + 3 * ee);   // :511-513 (nodes_ahead = 0)
    return 0;
}

}  // namespace kin
}  // namespace oracle
