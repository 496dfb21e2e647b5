// Joint-space targets from the MPC trajectory: MPCController::GetTargetsFromTraj
// (controllers/mpc_controller.cpp:414-511) over SingleRigidBodyModel::InverseKinematics
// (mpc/models/single_rigid_body_model.cpp:314-425, ComputeJacobianForIK :430-441).
//
// The reference does the rigid-body arithmetic with pinocchio on the full 19-dof model; the A1's legs are three-joint chains off
// the floating base, so everything the IK loop needs is closed form here: forward kinematics of the base and ONE leg, the LOCAL
// foot Jacobian (9 non-zero columns: 6 base + 3 leg), log6 / Jlog6 of the body error, a 9 x 9 damped normal-equation solve, the
// free-flyer integrate (exp6, rotation -> quaternion, first-order normalisation).  One thread per IK problem: the loop is ~90
// dependent iterations per foot with a few hundred flops each -- latency, not throughput; the batch supplies the parallelism.
#include <cuda_runtime.h>

#include "bgg_kernels.cuh"
#include "bgg_spline.cuh"

namespace bgg {
namespace {

constexpr double kTaylor3 = 1.220703125e-4;   // pinocchio TaylorSeriesExpansion<double>::precision<3>() = eps^(1/4) = 2^-13
constexpr double kPiD = 3.14159265358979323846;

struct Se3 {
    double R[9];
    double p[3];
};

__device__ __forceinline__ void cross3(const double* a, const double* b, double* o) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ __forceinline__ double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
__device__ __forceinline__ void mat_vec(const double* R, const double* v, double* o) {
    #pragma unroll
    for (int i = 0; i < 3; ++i) o[i] = R[3 * i] * v[0] + R[3 * i + 1] * v[1] + R[3 * i + 2] * v[2];
}
__device__ __forceinline__ void mat_t_vec(const double* R, const double* v, double* o) {
    #pragma unroll
    for (int i = 0; i < 3; ++i) o[i] = R[i] * v[0] + R[3 + i] * v[1] + R[6 + i] * v[2];
}
__device__ __forceinline__ void mat_mul(const double* A, const double* B, double* C) {
    #pragma unroll
    for (int i = 0; i < 3; ++i)
        #pragma unroll
        for (int j = 0; j < 3; ++j) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
__device__ __forceinline__ void mat_t_mul(const double* A, const double* B, double* C) {
    #pragma unroll
    for (int i = 0; i < 3; ++i)
        #pragma unroll
        for (int j = 0; j < 3; ++j) C[3 * i + j] = A[i] * B[j] + A[3 + i] * B[3 + j] + A[6 + i] * B[6 + j];
}
__device__ __forceinline__ Se3 se3_mul(const Se3& a, const Se3& b) {
    Se3 o;
    mat_mul(a.R, b.R, o.R);
    double t[3];
    mat_vec(a.R, b.p, t);
    for (int i = 0; i < 3; ++i) o.p[i] = a.p[i] + t[i];
    return o;
}
__device__ __forceinline__ Se3 se3_act_inv(const Se3& a, const Se3& b) {   // a^-1 b
    Se3 o;
    mat_t_mul(a.R, b.R, o.R);
    const double d[3] = {b.p[0] - a.p[0], b.p[1] - a.p[1], b.p[2] - a.p[2]};
    mat_t_vec(a.R, d, o.p);
    return o;
}
__device__ __forceinline__ Se3 se3_inverse(const Se3& a) {
    Se3 o;
    #pragma unroll
    for (int i = 0; i < 3; ++i)
        #pragma unroll
        for (int j = 0; j < 3; ++j) o.R[3 * i + j] = a.R[3 * j + i];
    double t[3];
    mat_vec(o.R, a.p, t);
    for (int i = 0; i < 3; ++i) o.p[i] = -t[i];
    return o;
}
// Eigen::Quaternion::toRotationMatrix (no normalisation), coefficients x y z w
__device__ void quat_to_matrix(const double* q, double* R) {
    const double x = q[0], y = q[1], z = q[2], w = q[3];
    const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
    const double twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
    R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
    R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}
// Eigen's rotation matrix -> quaternion (what pinocchio's assignQuaternion evaluates)
__device__ void matrix_to_quat(const double* m, double* q) {
    double t = m[0] + m[4] + m[8];
    if (t > 0) {
        t = sqrt(t + 1.0);
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
        t = sqrt(m[4 * i] - m[4 * j] - m[4 * k] + 1.0);
        q[i] = 0.5 * t;
        t = 0.5 / t;
        q[3] = (m[3 * k + j] - m[3 * j + k]) * t;
        q[j] = (m[3 * j + i] + m[3 * i + j]) * t;
        q[k] = (m[3 * k + i] + m[3 * i + k]) * t;
    }
}
__device__ void axis_angle(const double* a, double ang, double* R) {
    double s, c;
    sincos(ang, &s, &c);
    const double v = 1.0 - c;
    R[0] = a[0] * a[0] * v + c;        R[1] = a[0] * a[1] * v - a[2] * s; R[2] = a[0] * a[2] * v + a[1] * s;
    R[3] = a[1] * a[0] * v + a[2] * s; R[4] = a[1] * a[1] * v + c;        R[5] = a[1] * a[2] * v - a[0] * s;
    R[6] = a[2] * a[0] * v - a[1] * s; R[7] = a[2] * a[1] * v + a[0] * s; R[8] = a[2] * a[2] * v + c;
}
// pinocchio exp6 (spatial/explog.hpp); nu = (v, w)
__device__ void exp6(const double* nu, Se3& M) {
    const double* v = nu;
    const double* w = nu + 3;
    const double t2 = dot3(w, w), t = sqrt(t2), wv = dot3(w, v);
    double alpha_wxv, alpha_v, alpha_w, diag;
    if (t > kTaylor3) {
        double st, ct;
        sincos(t, &st, &ct);
        const double inv_t2 = 1.0 / t2;
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
    cross3(w, v, wxv);
    for (int i = 0; i < 3; ++i) M.p[i] = alpha_v * v[i] + alpha_w * w[i] + alpha_wxv * wxv[i];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) M.R[3 * i + j] = alpha_wxv * w[i] * w[j];
    M.R[1] -= alpha_v * w[2]; M.R[3] += alpha_v * w[2];
    M.R[2] += alpha_v * w[1]; M.R[6] -= alpha_v * w[1];
    M.R[5] -= alpha_v * w[0]; M.R[7] += alpha_v * w[0];
    M.R[0] += diag; M.R[4] += diag; M.R[8] += diag;
}
// pinocchio log3 (spatial/log.hxx)
__device__ void log3(const double* R, double* w, double& theta) {
    const double tr = R[0] + R[4] + R[8];
    if (tr >= 3.0) theta = 0.0;
    else if (tr <= -1.0) theta = kPiD;
    else theta = acos((tr - 1.0) / 2.0);
    if (theta >= kPiD - 1e-2) {
        const double cphi = -(tr - 1.0) / 2.0, beta = theta * theta / (1.0 + cphi);
        const double t0 = (R[0] + cphi) * beta, t1 = (R[4] + cphi) * beta, t2 = (R[8] + cphi) * beta;
        w[0] = (R[7] > R[5] ? 1.0 : -1.0) * (t0 > 0 ? sqrt(t0) : 0.0);
        w[1] = (R[2] > R[6] ? 1.0 : -1.0) * (t1 > 0 ? sqrt(t1) : 0.0);
        w[2] = (R[3] > R[1] ? 1.0 : -1.0) * (t2 > 0 ? sqrt(t2) : 0.0);
    } else {
        const double t = ((theta > kTaylor3) ? theta / sin(theta) : 1.0) / 2.0;
        w[0] = t * (R[7] - R[5]);
        w[1] = t * (R[2] - R[6]);
        w[2] = t * (R[3] - R[1]);
    }
}
__device__ void jlog3(double theta, const double* w, double* J) {
    double alpha, diag;
    if (theta < kTaylor3) {
        alpha = 1.0 / 12.0 + theta * theta / 720.0;
        diag = 0.5 * (2.0 - theta * theta / 6.0);
    } else {
        double st, ct;
        sincos(theta, &st, &ct);
        const double st_1mct = st / (1.0 - ct);
        alpha = 1.0 / (theta * theta) - st_1mct / (2.0 * theta);
        diag = 0.5 * (theta * st_1mct);
    }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) J[3 * i + j] = alpha * w[i] * w[j];
    J[0] += diag; J[4] += diag; J[8] += diag;
    J[1] -= 0.5 * w[2]; J[3] += 0.5 * w[2];
    J[2] += 0.5 * w[1]; J[6] -= 0.5 * w[1];
    J[5] -= 0.5 * w[0]; J[7] += 0.5 * w[0];
}
__device__ void log6(const Se3& M, double* out) {
    double w[3], t;
    log3(M.R, w, t);
    const double t2 = t * t;
    double alpha, beta;
    if (t < kTaylor3) {
        alpha = 1.0 - t2 / 12.0 - t2 * t2 / 720.0;
        beta = 1.0 / 12.0 + t2 / 720.0;
    } else {
        double st, ct;
        sincos(t, &st, &ct);
        alpha = t * st / (2.0 * (1.0 - ct));
        beta = 1.0 / t2 - st / (2.0 * t * (1.0 - ct));
    }
    double wxp[3];
    cross3(w, M.p, wxp);
    const double wp = dot3(w, M.p);
    for (int i = 0; i < 3; ++i) {
        out[i] = alpha * M.p[i] - 0.5 * wxp[i] + (beta * wp) * w[i];
        out[3 + i] = w[i];
    }
}
// pinocchio Jlog6: [[A, B], [0, A]], A = Jlog3(R), B = C A.  Only A and B are returned.
__device__ void jlog6(const Se3& M, double* A, double* B) {
    double w[3], t;
    log3(M.R, w, t);
    jlog3(t, w, A);
    const double t2 = t * t;
    double beta, bdot;
    if (t < kTaylor3) {
        beta = 1.0 / 12.0 + t2 / 720.0;
        bdot = 1.0 / 360.0;
    } else {
        double st, ct;
        sincos(t, &st, &ct);
        const double tinv = 1.0 / t, t2inv = tinv * tinv, inv_2_2ct = 1.0 / (2.0 * (1.0 - ct));
        beta = t2inv - st * tinv * inv_2_2ct;
        bdot = -2.0 * t2inv * t2inv + (1.0 + st * tinv) * t2inv * inv_2_2ct;
    }
    const double* p = M.p;
    const double wTp = dot3(w, p);
    double v3[3], C[9];
    for (int i = 0; i < 3; ++i) v3[i] = (bdot * wTp) * w[i] - (t2 * bdot + 2.0 * beta) * p[i];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) C[3 * i + j] = v3[i] * w[j] + beta * w[i] * p[j];
    C[0] += wTp * beta; C[4] += wTp * beta; C[8] += wTp * beta;
    C[1] -= 0.5 * p[2]; C[3] += 0.5 * p[2];
    C[2] += 0.5 * p[1]; C[6] -= 0.5 * p[1];
    C[5] -= 0.5 * p[0]; C[7] += 0.5 * p[0];
    mat_mul(C, A, B);
}

// SingleRigidBodyModel::InverseKinematics.  q: [p, quat xyzw, 12 joints]; returns 0 or 1 ("IK did not converge.")
__device__ int inverse_kinematics(const RobotKin& rk, const double* state, const double* ee_des, const double* joint_guess, double* q, int* iters) {
    const double eps = 5e-6, DT = 1e-1, damp = 1e-6;   // single_rigid_body_model.cpp:345-348
    const int IT_MAX = 1000;
    Se3 body_des;                                      // :336-337
    quat_to_matrix(state + 6, body_des.R);
    for (int i = 0; i < 3; ++i) body_des.p[i] = state[i];
    for (int i = 0; i < 3; ++i) q[i] = state[i];        // :339-342
    for (int i = 0; i < 4; ++i) q[3 + i] = state[6 + i];
    for (int i = 0; i < 12; ++i) q[7 + i] = joint_guess[i];
    bool success = false;                              // :363 -- set once, never cleared between feet
    for (int ee = 0; ee < kNumEE; ++ee) {
        const LegChain& lc = rk.leg[ee];
        int used = IT_MAX;
        for (int it = 0; it < IT_MAX; ++it) {
            const double n2 = q[3] * q[3] + q[4] * q[4] + q[5] * q[5] + q[6] * q[6];   // :372-377 firstOrderNormalize
            const double nrm_fix = (3.0 - n2) / 2.0;
            for (int i = 0; i < 4; ++i) q[3 + i] *= nrm_fix;
            // forward kinematics of the base and this leg (:379-380)
            Se3 base, jt[3], foot;
            quat_to_matrix(q + 3, base.R);
            for (int i = 0; i < 3; ++i) base.p[i] = q[i];
            Se3 parent = base;
            for (int j = 0; j < 3; ++j) {
                Se3 place, rot;
                for (int i = 0; i < 9; ++i) place.R[i] = lc.R[j][i];
                for (int i = 0; i < 3; ++i) place.p[i] = lc.t[j][i];
                axis_angle(lc.axis[j], q[7 + 3 * ee + j], rot.R);
                rot.p[0] = rot.p[1] = rot.p[2] = 0.0;
                parent = se3_mul(parent, se3_mul(place, rot));
                jt[j] = parent;
            }
            {
                Se3 fp;
                for (int i = 0; i < 9; ++i) fp.R[i] = lc.R[3][i];
                for (int i = 0; i < 3; ++i) fp.p[i] = lc.t[3][i];
                foot = se3_mul(parent, fp);
            }
            double err[9];
            const double d[3] = {ee_des[3 * ee] - foot.p[0], ee_des[3 * ee + 1] - foot.p[1], ee_des[3 * ee + 2] - foot.p[2]};
            mat_t_vec(foot.R, d, err);                                  // :382-384
            const Se3 body_err = se3_act_inv(base, body_des);           // :386-388
            log6(body_err, err + 3);
            double nrm = 0;
            for (int i = 0; i < 9; ++i) nrm += err[i] * err[i];
            if (sqrt(nrm) < eps) {                                      // :390-393
                success = true;
                used = it;
                break;
            }
            // J (9 x 9 non-zero columns: 6 base + this leg's 3): rows 0-2 = -(LOCAL foot Jacobian, linear part) (:395-396),
            // rows 3-8 = -Jlog6(body_err^-1) [I6 0] (:398-399, ComputeJacobianForIK)
            double J[81];
            for (int i = 0; i < 81; ++i) J[i] = 0.0;
            const Se3 bMf = se3_act_inv(base, foot);
            for (int k = 0; k < 3; ++k) {
                double e[3] = {0, 0, 0}, exp_[3], lin[3];
                e[k] = 1.0;
                mat_t_vec(bMf.R, e, lin);
                for (int i = 0; i < 3; ++i) J[9 * i + k] = -lin[i];
                cross3(e, bMf.p, exp_);
                mat_t_vec(bMf.R, exp_, lin);
                for (int i = 0; i < 3; ++i) J[9 * i + 3 + k] = -lin[i];
            }
            for (int j = 0; j < 3; ++j) {
                const Se3 jMf = se3_act_inv(jt[j], foot);
                double axp[3], lin[3];
                cross3(lc.axis[j], jMf.p, axp);
                mat_t_vec(jMf.R, axp, lin);
                for (int i = 0; i < 3; ++i) J[9 * i + 6 + j] = -lin[i];
            }
            {
                double A[9], Bm[9];
                jlog6(se3_inverse(body_err), A, Bm);
                for (int i = 0; i < 3; ++i)
                    for (int j = 0; j < 3; ++j) {
                        J[9 * (3 + i) + j] = -A[3 * i + j];
                        J[9 * (3 + i) + 3 + j] = -Bm[3 * i + j];
                        J[9 * (6 + i) + 3 + j] = -A[3 * i + j];
                    }
            }
            double M[81];                                               // :402-404
            for (int i = 0; i < 9; ++i)
                for (int j = 0; j <= i; ++j) {
                    double s = 0;
                    for (int k = 0; k < 9; ++k) s += J[9 * i + k] * J[9 * j + k];
                    M[9 * i + j] = s + (i == j ? damp : 0.0);
                }
            double y[9];                                                // L D L' (positive definite: no pivoting needed)
            for (int j = 0; j < 9; ++j) {
                double dj = M[9 * j + j];
                for (int k = 0; k < j; ++k) dj -= M[9 * j + k] * M[9 * j + k] * M[9 * k + k];
                M[9 * j + j] = dj;
                for (int i = j + 1; i < 9; ++i) {
                    double s = M[9 * i + j];
                    for (int k = 0; k < j; ++k) s -= M[9 * i + k] * M[9 * j + k] * M[9 * k + k];
                    M[9 * i + j] = s / dj;
                }
            }
            for (int i = 0; i < 9; ++i) {
                double s = err[i];
                for (int k = 0; k < i; ++k) s -= M[9 * i + k] * y[k];
                y[i] = s;
            }
            for (int i = 0; i < 9; ++i) y[i] /= M[9 * i + i];
            for (int i = 8; i >= 0; --i) {
                double s = y[i];
                for (int k = i + 1; k < 9; ++k) s -= M[9 * k + i] * y[k];
                y[i] = s;
            }
            double v[9];                                                // :405 v = -J' (JJ' + damp)^-1 err, times DT
            for (int j = 0; j < 9; ++j) {
                double s = 0;
                for (int i = 0; i < 9; ++i) s += J[9 * i + j] * y[i];
                v[j] = -s * DT;
            }
            // :406 pinocchio::integrate: free flyer M0 exp6(v), rotation -> quaternion on the side of the old one, first-order
            // normalisation; revolute joints add
            Se3 E;
            exp6(v, E);
            const Se3 M1 = se3_mul(base, E);
            double rq[4];
            matrix_to_quat(M1.R, rq);
            const double dq = rq[0] * q[3] + rq[1] * q[4] + rq[2] * q[5] + rq[3] * q[6];
            if (dq < 0)
                for (int i = 0; i < 4; ++i) rq[i] = -rq[i];
            const double m2 = rq[0] * rq[0] + rq[1] * rq[1] + rq[2] * rq[2] + rq[3] * rq[3];
            const double a = (3.0 - m2) / 2.0;
            for (int i = 0; i < 3; ++i) q[i] = M1.p[i];
            for (int i = 0; i < 4; ++i) q[3 + i] = rq[i] * a;
            for (int j = 0; j < 3; ++j) q[7 + 3 * ee + j] += v[6 + j];
        }
        if (iters) iters[ee] = used;
        if (!success) return 1;                                         // :417-420
    }
    return 0;
}

__global__ void __launch_bounds__(32) k_ik(RobotKin rk, int count, const double* __restrict__ state, const double* __restrict__ ee_des,
                                           const double* __restrict__ joint_guess, double* __restrict__ q_out, int* __restrict__ status,
                                           int* __restrict__ iters) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= count) return;
    double q[19];
    int it[4] = {0, 0, 0, 0};
    const int rc = inverse_kinematics(rk, state + 13 * b, ee_des + 12 * b, joint_guess + 12 * b, q, it);
    for (int i = 0; i < 19; ++i) q_out[19 * b + i] = q[i];
    status[b] = rc;
    if (iters)
        for (int i = 0; i < 4; ++i) iters[4 * b + i] = it[i];
}

// MPCController::GetTargetsFromTraj (controllers/mpc_controller.cpp:414-511) for every instance of the batch
__global__ void __launch_bounds__(32) k_targets_from_traj(Params P, RobotKin rk, const Instance* __restrict__ inst, int B, const double* __restrict__ time_in,
                                                          double* __restrict__ q_des, double* __restrict__ v_des, double* __restrict__ force_des,
                                                          int* __restrict__ status) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const Instance& I = inst[b];
    const double dt = P.dt;
    double time = time_in[b];
    if (time < I.init_time) time = I.init_time;                                   // :415-417 (GetTime(0) = init_time)
    const int node = static_cast<int>(ceil((time - I.init_time) / dt));           // :420, Trajectory::GetNode (trajectory.cpp:479-481)
    if (node < 0 || node + 1 > P.N) {                                             // the reference's states_.at(node + 1) throws
        status[b] = 3;
        return;
    }
    auto t_of = [&](int k) { return I.init_time + dt * k; };                      // Trajectory::GetTime (trajectory.cpp:413-415)
    auto lerp = [&](int a, int bb, double at, double* out) {                      // (x_b - x_a) (1 - (t_b - at) / (t_b - t_a)) + x_a
        const double w = 1 - (t_of(bb) - at) / (t_of(bb) - t_of(a));
        for (int i = 0; i < kNxMan; ++i) out[i] = (I.states[bb][i] - I.states[a][i]) * w + I.states[a][i];
    };
    double s1[kNxMan], s2[kNxMan];
    if (node > 0) {                                                               // :431-447
        lerp(node - 1, node, time, s1);
        if (time + dt < t_of(node)) {
            status[b] = 2;                                                        // "bad interp."
            return;
        }
        lerp(node, node + 1, time + dt, s2);
    } else {                                                                      // :448-456
        lerp(node, node + 1, time, s1);
        lerp(node, node + 1, time + dt, s2);
    }
    double ee1[12], ee2[12];
    for (int e = 0; e < kNumEE; ++e)
        for (int c = 0; c < 3; ++c) {                                             // :460-464, :477-481 / :491-495
            ee1[3 * e + c] = value_at(I.foot[e], false, c, time);
            ee2[3 * e + c] = value_at(I.foot[e], false, c, time + dt);
        }
    double q1[19], q2[19];
    if (inverse_kinematics(rk, s1, ee1, q_des + 19 * b + 7, q1, nullptr)) {       // :466-468
        status[b] = 1;
        return;
    }
    if (inverse_kinematics(rk, s2, ee2, q1 + 7, q2, nullptr)) {                   // :483-485 / :497-499, guess = the new q_des_
        status[b] = 1;
        return;
    }
    double* v = v_des + 18 * b;
    for (int i = 0; i < 3; ++i) {                                                 // :471-473
        v[i] = s1[3 + i] / P.mass;
        v[3 + i] = P.Ir_inv[3 * i] * s1[10] + P.Ir_inv[3 * i + 1] * s1[11] + P.Ir_inv[3 * i + 2] * s1[12];
    }
    for (int i = 0; i < 12; ++i) v[6 + i] = (q2[7 + i] - q1[7 + i]) / dt;         // :487 and :501 are the same difference
    for (int i = 0; i < 19; ++i) q_des[19 * b + i] = q1[i];
    for (int e = 0; e < kNumEE; ++e)
        for (int c = 0; c < 3; ++c) force_des[12 * b + 3 * e + c] = value_at(I.foot[e], true, c, time);   // :511-513
    status[b] = 0;
}

}  // namespace

void launch_ik(const RobotKin& rk, int count, const double* state, const double* ee_des, const double* joint_guess, double* q, int* status,
               int* iters, cudaStream_t stream) {
    k_ik<<<(count + 31) / 32, 32, 0, stream>>>(rk, count, state, ee_des, joint_guess, q, status, iters);
}

void launch_targets_from_traj(const Params& P, const RobotKin& rk, const Instance* inst, int B, const double* time, double* q_des, double* v_des,
                              double* force_des, int* status, cudaStream_t stream) {
    k_targets_from_traj<<<(B + 31) / 32, 32, 0, stream>>>(P, rk, inst, B, time, q_des, v_des, force_des, status);
}

}  // namespace bgg
