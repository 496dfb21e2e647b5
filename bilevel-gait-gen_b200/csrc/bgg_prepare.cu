// bilevel-gait-gen_b200 -- kernel 1+2+3a: horizon maintenance, contact-spline evaluation, single-rigid-body
// linearisation + Euler discretisation, and the structured constraint rows, one CTA per MPC instance.
//
// Replaces, for a whole batch at once, steps 1-5 of the reference's Solve() (mpc_single_rigid_body.cpp:25-103):
//   prev_traj_.AddPolys / RemoveUnusedPolys           trajectory.cpp:225-246
//   ConvertTrajToQPVec                                mpc_single_rigid_body.cpp:343-357
//   GetLinearDynamics + Euler discretisation          single_rigid_body_model.cpp:55-169, mpc_single_rigid_body.cpp:236-248
//   force-box / friction-cone sample rows             mpc.cpp:166-209, 352-414
//   foot-box, touch-down and foot-start rows          mpc_single_rigid_body.cpp:381-475, 849-887
// Compiled with -fmad=false: entries that are exactly 0.0 in the reference (and therefore absent from its sparse
// matrix, utils/sparse_matrix_builder.cpp:25) must be exactly 0.0 here as well.
#include "bgg_kernels.cuh"

namespace bgg {

__device__ void quat_log3(const double q[4], double out[3]) {
    // pinocchio::quaternion::log3 (published algorithm): Taylor branch below eps^(1/3) on |vec|^2
    const double eps = 2.220446049250313e-16;
    const double ts_prec = 6.055454452393343e-06;   // eps^(1/3)
    const double n2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2];
    const double norm = sqrt(n2 + eps * eps);
    const double sgn = (q[3] >= 0) ? 1.0 : -1.0;
    const double w = sgn * q[3];
    const double theta_2 = atan2(norm, w);
    const double y_x = norm / w;
    const double y_x_sq = n2 / (w * w);
    const double theta = (n2 < ts_prec) ? 2.0 * (1.0 - y_x_sq / 3.0) * y_x : 2.0 * theta_2;
    const double th2_2 = theta * theta / 4.0;
    const double inv_sinc = (n2 < ts_prec) ? 2.0 * (1.0 + th2_2 / 6.0 + 7.0 / 360.0 * th2_2 * th2_2) : theta / sin(theta_2);
    for (int k = 0; k < 3; ++k) out[k] = inv_sinc * (sgn * q[k]);
}

__device__ void quat_exp3(const double v[3], double q[4]) {
    // pinocchio::quaternion::exp3: Taylor branch when |v|^2 <= eps^(1/4)
    const double eps = 2.220446049250313e-16;
    const double ts_prec = 1.220703125e-04;         // eps^(1/4)
    const double t2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
    const double t = sqrt(t2 + eps * eps);
    if (t2 > ts_prec) {
        const double s = sin(t / 2), c = cos(t / 2);
        for (int k = 0; k < 3; ++k) q[k] = s * (v[k] / t);
        q[3] = c;
    } else {
        const double k_ = 0.5 - t2 / 48.0;
        for (int k = 0; k < 3; ++k) q[k] = k_ * v[k];
        q[3] = 1.0 - t2 / 8.0;
    }
}

__device__ void quat_first_order_normalize(double q[4]) {
    const double n2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
    const double a = (3.0 - n2) / 2.0;
    for (int k = 0; k < 4; ++k) q[k] *= a;
}

__device__ __forceinline__ void cross3(const double a[3], const double b[3], double o[3]) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}

__global__ void __launch_bounds__(128, 6) k_prepare(Params P, Instance* __restrict__ inst, const double* __restrict__ state_in,
                                                 const double* __restrict__ t0_in, const double* __restrict__ ee_start,
                                                 WsLayout L, char* __restrict__ ws_base) {
    const int b = blockIdx.x, tid = threadIdx.x, nth = blockDim.x;
    Instance& I = inst[b];
    char* ws = ws_base + static_cast<size_t>(b) * L.stride;
    WsHeader* H = reinterpret_cast<WsHeader*>(ws + L.hdr);
    NodeLin* nodes = reinterpret_cast<NodeLin*>(ws + L.nodes);
    Sample* samples = reinterpret_cast<Sample*>(ws + L.samples);
    EqRow* eqs = reinterpret_cast<EqRow*>(ws + L.eq);
    double* zprev = reinterpret_cast<double*>(ws + L.zprev);

    __shared__ FootSpline sf[kNumEE];
    __shared__ WsHeader sh;
    __shared__ double s_state[kMaxNodes + 1][kNxMan];
    __shared__ double st_lo[kNumEE][kMaxStances], st_hi[kNumEE][kMaxStances];
    __shared__ int st_n[kNumEE], st_base[kNumEE + 1];

    const int N = P.N;
    const double t0 = t0_in[b];
    {
        const double* src = reinterpret_cast<const double*>(I.foot);
        double* dst = reinterpret_cast<double*>(sf);
        for (int i = tid; i < static_cast<int>(sizeof(FootSpline) * kNumEE / 8); i += nth) dst[i] = src[i];
        for (int i = tid; i < (N + 1) * kNxMan; i += nth) {
            const int k = i / kNxMan, c = i % kNxMan;
            s_state[k][c] = (k == 0) ? state_in[b * kNxMan + c] : I.states[k][c];   // prev_traj_.SetState(0, state)
        }
    }
    __syncthreads();

    // ---- horizon maintenance, one thread per foot (mpc_single_rigid_body.cpp:36-38)
    if (tid < kNumEE) {
        FootSpline& s = sf[tid];
        add_polys_until(s, P.dt * N + t0);
        set_swing_pos_z(s, P.swing_height, P.foot_offset);
        remove_poly(s, t0);
        // stance segments: a TouchDown contact that is not the last contact (mpc.cpp:171-173)
        int ns = 0, prev = -1;
        for (int i = 0; i < s.n; ++i) {
            if (s.ttype[i] == kInter) continue;
            if (prev >= 0 && s.ttype[prev] == kTouchDown) {
                if (ns < kMaxStances) {
                    st_lo[tid][ns] = s.t[prev];
                    st_hi[tid][ns] = s.t[i];
                }
                ns++;   // counted beyond the cap: the instance is refused (error bit 2) rather than solved without those rows
            }
            prev = i;
        }
        st_n[tid] = ns;
    }
    __syncthreads();

    if (tid == 0) {
        int err = 0, fb = 0, pb = 0;
        st_base[0] = 0;
        for (int e = 0; e < kNumEE; ++e) {
            sh.nfv[e] = num_force_vars(sf[e]);
            sh.npv[e] = num_pos_vars(sf[e]);
            sh.fbase[e] = fb;
            sh.pbase[e] = pb;
            fb += 3 * sh.nfv[e];
            pb += 2 * sh.npv[e];
            if (st_n[e] > kMaxStances) {
                err |= 4;
                st_n[e] = kMaxStances;
            }
            st_base[e + 1] = st_base[e] + st_n[e] * kSamplesPerStance;
            if (sf[e].n + kNumForcePolys > kMaxKnots) err |= 1;
            // a touch-down row pair when the next touch-down is closer than td_fraction of the current swing
            // (mpc.cpp:1205-1214)
            sh.td_flag[e] = (next_touchdown_time(sf[e], t0) - t0 < P.td_fraction * swing_time(sf[e], t0)) ? 1 : 0;
        }
        sh.nf = fb;
        sh.np = pb;
        sh.nu = fb + pb;
        sh.n = kNx * (N + 1) + sh.nu;
        sh.n_samples = st_base[kNumEE];
        sh.n_eebox = (N - (kEENodeStart - 1)) * kNumEE * 2;
        sh.n_td = 2 * (sh.td_flag[0] + sh.td_flag[1] + sh.td_flag[2] + sh.td_flag[3]);
        sh.n_eq = sh.n_td + 2 * kNumEE;
        sh.m_ineq = 6 * sh.n_samples + 2 * sh.n_eebox;
        if (sh.nu > L.max_nu) err |= 2;
        if (sh.n_samples > kMaxSamples) err |= 4;
        sh.error = err;
        sh.t0 = t0;
        sh.status = kUnsolved;
        sh.iters = 0;
        sh.ls_iters = 0;
        sh.no_iterate = 0;
        sh.refined_iters = 0;
        sh.pass_state = 0;
        sh.cost = 0.0;      // k_finish writes these; a refused instance (error != 0) must not carry stale or uninitialised values
        sh.qp_cost = 0.0;
        sh.alpha = 0.0;
        sh.ee_box[0] = I.ee_box[0];
        sh.ee_box[1] = I.ee_box[1];
    }
    __syncthreads();
    const int nf = sh.nf, nu = sh.nu, ustart = kNx * (N + 1);
    if (sh.error) {
        if (tid == 0) *H = sh;
        return;
    }

    // ---- z_prev = ConvertTrajToQPVec(prev_traj_)
    for (int k = tid; k <= N; k += nth) {
        double* x = zprev + k * kNx;
        for (int i = 0; i < 6; ++i) x[i] = s_state[k][i];
        quat_log3(&s_state[k][6], x + 6);
        for (int i = 0; i < 3; ++i) x[9 + i] = s_state[k][10 + i];
    }
    for (int i = tid; i < kNumEE * 5; i += nth) {
        const int e = i / 5, c = i % 5;
        if (c < 3) get_force_vars(sf[e], c, zprev + ustart + sh.fbase[e] + c * sh.nfv[e]);
        else get_pos_vars(sf[e], c - 3, zprev + ustart + nf + sh.pbase[e] + (c - 3) * sh.npv[e]);
    }
    __syncthreads();

    // ---- spline queries at every node time t_k = k dt + t0 (MPC::GetTime, mpc.cpp:779-781)
    for (int i = tid; i < (N + 1) * kNumEE; i += nth) {
        const int k = i / kNumEE, e = i % kNumEE;
        const double tk = k * P.dt + t0;
        NodeLin& nl = nodes[k];
        const FootSpline& s = sf[e];
        for (int c = 0; c < 3; ++c) {
            nl.f[e][c] = value_at(s, true, c, tk);
            nl.rel[e][c] = value_at(s, false, c, tk) - s_state[k][c];
        }
        double w[4] = {0, 0, 0, 0};
        int off = 0;
        const int cnt = force_lin(s, tk, w, &off);
        nl.fcnt[e] = cnt;
        nl.foff[e] = off;
        for (int j = 0; j < 4; ++j) nl.fw[e][j] = (j < cnt) ? w[j] : 0.0;
        double pw[2] = {0, 0};
        const int pc = pos_lin(s, tk, pw, &off);
        nl.pcnt[e] = pc;
        nl.poff[e] = off;
        nl.pw[e][0] = pw[0];
        nl.pw[e][1] = (pc > 1) ? pw[1] : 0.0;
    }
    __syncthreads();

    // ---- per-node linearisation and Euler discretisation
    for (int k = tid; k < N; k += nth) {
        NodeLin& nl = nodes[k];
        const double* x = zprev + k * kNx;
        const double* u = zprev + ustart;
        const double* om = &s_state[k][10];
        double A[kNx][kNx];
        for (int i = 0; i < kNx; ++i)
            for (int j = 0; j < kNx; ++j) A[i][j] = 0.0;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                A[i][3 + j] = ((i == j) ? 1.0 : 0.0) / P.mass;
                A[6 + i][9 + j] = P.Ir_inv[3 * i + j];
            }
        double Iw[3];
        for (int i = 0; i < 3; ++i) Iw[i] = P.Ir[3 * i] * om[0] + P.Ir[3 * i + 1] * om[1] + P.Ir[3 * i + 2] * om[2];
        for (int i = 0; i < 3; ++i) {
            const double ei[3] = {i == 0 ? 1.0 : 0.0, i == 1 ? 1.0 : 0.0, i == 2 ? 1.0 : 0.0};
            const double Icol[3] = {P.Ir[i], P.Ir[3 + i], P.Ir[6 + i]};
            double c1[3], c2[3];
            cross3(ei, Iw, c1);
            cross3(om, Icol, c2);
            for (int r = 0; r < 3; ++r) A[9 + r][9 + i] = -c1[r] - c2[r];
            for (int e = 0; e < kNumEE; ++e) {
                double c3[3];
                cross3(ei, nl.f[e], c3);
                for (int r = 0; r < 3; ++r) A[9 + r][i] += -c3[r];
            }
        }
        // C = -A x - B u + f(x, t_k)
        double C[kNx];
        for (int i = 0; i < kNx; ++i) {
            double s = 0;
            for (int j = 0; j < kNx; ++j) s += -A[i][j] * x[j];
            C[i] = s;
        }
        double Bu[kNx];
        for (int i = 0; i < kNx; ++i) Bu[i] = 0.0;
        for (int e = 0; e < kNumEE; ++e) {
            for (int c = 0; c < 3; ++c) {
                const double ec[3] = {c == 0 ? 1.0 : 0.0, c == 1 ? 1.0 : 0.0, c == 2 ? 1.0 : 0.0};
                if (nl.fcnt[e] > 0) {
                    const double* uf = u + sh.fbase[e] + c * sh.nfv[e] + nl.foff[e];
                    double rc[3];
                    cross3(nl.rel[e], ec, rc);
                    for (int j = 0; j < nl.fcnt[e]; ++j) {
                        Bu[3 + c] += nl.fw[e][j] * uf[j];
                        for (int r = 0; r < 3; ++r) Bu[9 + r] += (rc[r] * nl.fw[e][j]) * uf[j];
                    }
                }
                if (c != 2) {
                    const double* up = u + nf + sh.pbase[e] + c * sh.npv[e] + nl.poff[e];
                    double ef[3];
                    cross3(ec, nl.f[e], ef);
                    for (int j = 0; j < nl.pcnt[e]; ++j)
                        for (int r = 0; r < 3; ++r) Bu[9 + r] += (ef[r] * nl.pw[e][j]) * up[j];
                }
            }
        }
        // CalcDynamics (single_rigid_body_model.cpp:222-256)
        double fd[kNx];
        for (int i = 0; i < 3; ++i) fd[i] = x[3 + i] / P.mass;
        for (int i = 0; i < 3; ++i) fd[3 + i] = P.mass * P.gravity[i];
        for (int i = 0; i < 3; ++i) fd[6 + i] = P.Ir_inv[3 * i] * om[0] + P.Ir_inv[3 * i + 1] * om[1] + P.Ir_inv[3 * i + 2] * om[2];
        double wxIw[3];
        cross3(om, Iw, wxIw);
        for (int i = 0; i < 3; ++i) fd[9 + i] = -wxIw[i];
        for (int e = 0; e < kNumEE; ++e) {
            double tq[3];
            cross3(nl.rel[e], nl.f[e], tq);
            for (int i = 0; i < 3; ++i) {
                fd[3 + i] += nl.f[e][i];
                fd[9 + i] += tq[i];
            }
        }
        for (int i = 0; i < kNx; ++i) {
            C[i] = (C[i] - Bu[i]) + fd[i];
            nl.cd[i] = P.dt * C[i];
            for (int j = 0; j < kNx; ++j) nl.Ad[i * kNx + j] = ((i == j) ? 1.0 : 0.0) + P.dt * A[i][j];
        }
    }

    // ---- force samples (10 per stance): tau = (i/10)(t_LO - t_TD) + t_TD
    for (int i = tid; i < sh.n_samples; i += nth) {
        int e = 0;
        while (i >= st_base[e + 1]) ++e;
        const int loc = i - st_base[e];
        const int j = loc / kSamplesPerStance, q = loc % kSamplesPerStance;
        const double lower = st_lo[e][j], upper = st_hi[e][j];
        const double time = (static_cast<double>(q) / static_cast<double>(kSamplesPerStance)) * (upper - lower) + lower;
        Sample& sp = samples[i];
        double w[4] = {0, 0, 0, 0};
        int off = 0;
        const int cnt = force_lin(sf[e], time, w, &off);
        sp.ee = e;
        sp.off = off;
        sp.cnt = cnt;
        sp.time = time;
        int act = 0;
        for (int k = 0; k < 4; ++k) {
            sp.w[k] = (k < cnt) ? w[k] : 0.0;
            act |= (sp.w[k] != 0.0);
        }
        sp.active = act;
    }

    // ---- equality rows: touch-down rows first, then foot-start rows (constraint order, single_rigid_body_model.cpp:22-29)
    if (tid == 0) {
        int r = 0;
        for (int e = 0; e < kNumEE; ++e) {
            if (!sh.td_flag[e]) continue;
            const double td = next_touchdown_time(sf[e], t0);
            double w[2];
            int off;
            const int cnt = pos_lin(sf[e], td, w, &off);
            for (int c = 0; c < 2; ++c) {
                EqRow& q = eqs[r++];
                q.cnt = cnt;
                q.pad = e * 2 + c;   // rows of one (foot, coord) only touch that pair's position variables
                q.rhs = value_at(sf[e], false, c, td);
                for (int j = 0; j < 2; ++j) {
                    q.w[j] = (j < cnt) ? w[j] : 0.0;
                    q.col[j] = nf + sh.pbase[e] + c * sh.npv[e] + off + j;
                }
            }
        }
        for (int e = 0; e < kNumEE; ++e) {
            double w[2];
            int off;
            const int cnt = pos_lin(sf[e], 0 * P.dt + t0, w, &off);
            for (int c = 0; c < 2; ++c) {
                EqRow& q = eqs[r++];
                q.cnt = cnt;
                q.pad = e * 2 + c;
                q.rhs = ee_start[(b * kNumEE + e) * 3 + c];
                for (int j = 0; j < 2; ++j) {
                    q.w[j] = (j < cnt) ? w[j] : 0.0;
                    q.col[j] = nf + sh.pbase[e] + c * sh.npv[e] + off + j;
                }
            }
        }
        *H = sh;
        I.init_time = t0;
    }
    __syncthreads();
    // ---- write the maintained splines and the new initial state back to the instance
    {
        double* dst = reinterpret_cast<double*>(I.foot);
        const double* src = reinterpret_cast<const double*>(sf);
        for (int i = tid; i < static_cast<int>(sizeof(FootSpline) * kNumEE / 8); i += nth) dst[i] = src[i];
        if (tid < kNxMan) I.states[0][tid] = s_state[0][tid];
    }
    (void)nu;
}

void launch_prepare(const Params& P, Instance* inst, const double* state, const double* t0, const double* ee_start,
                    const WsLayout& L, char* ws, int B, cudaStream_t stream) {
    k_prepare<<<B, 128, 0, stream>>>(P, inst, state, t0, ee_start, L, ws);
}

}  // namespace bgg
