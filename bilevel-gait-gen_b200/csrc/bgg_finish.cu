// bilevel-gait-gen_b200 -- kernel 5: state recovery, the l1-merit backtracking line search and the trajectory
// update, one CTA per MPC instance (steps 8-11 of the reference's Solve(), mpc_single_rigid_body.cpp:131-184).
//
//   sol      = [x_0 .. x_N | u*]  with the states rolled out through the linearised dynamics (what the QP's
//              equality rows enforce)
//   p        = sol - prev_qp_sol ;  alpha by MPC::LineSearch (mpc.cpp:730-747): merit phi(z) = mu |d(z)|_1 + cost(z),
//              d_k = x_{k+1} - (x_k + dt f(x_k, t_k)) on the trajectory rebuilt from z (mpc.cpp:749-776,
//              rk_integrator.cpp:14-30), Armijo constant 1e-5, at most 10 halvings
//   prev_traj_ = ConvertQPSolToTrajectory(prev + alpha p)  (mpc_single_rigid_body.cpp:275-321), foot-box adaptation
//              (:136-144, 929-937) and the per-solve statistics of RecordStats (mpc.cpp:804-816).
#include "bgg_kernels.cuh"

namespace bgg {

__device__ __forceinline__ void cross3f(const double a[3], const double b[3], double o[3]) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}

size_t finish_smem_bytes(const WsLayout& L) {
    const size_t n_max = static_cast<size_t>(kNx) * (L.N + 1) + L.max_nu;
    return 2 * sizeof(FootSpline) * kNumEE + 8 * (4 * n_max + static_cast<size_t>(kNx) * (L.N + 1) + 64 + static_cast<size_t>(kNumEE) * 6 * L.N);
}

__global__ void __launch_bounds__(128, 5) k_finish(Params P, Instance* __restrict__ inst, WsLayout L, char* __restrict__ ws_base, int want) {
    const int b = blockIdx.x, tid = threadIdx.x, nth = blockDim.x;
    Instance& I = inst[b];
    char* ws = ws_base + static_cast<size_t>(b) * L.stride;
    WsHeader* Hd = reinterpret_cast<WsHeader*>(ws + L.hdr);
    if (Hd->error || Hd->pass_state != want) return;
    const NodeLin* nodes = reinterpret_cast<const NodeLin*>(ws + L.nodes);
    double* zprev_g = reinterpret_cast<double*>(ws + L.zprev);
    double* zqp_g = reinterpret_cast<double*>(ws + L.zqp);
    const double* ustar = reinterpret_cast<const double*>(ws + L.u);
    const double* xoff = reinterpret_cast<const double*>(ws + L.xoff);

    const int N = P.N, nu = Hd->nu, nf = Hd->nf, n = Hd->n, ustart = kNx * (N + 1);
    const double t0 = Hd->t0;
    const int n_max = kNx * (L.N + 1) + L.max_nu;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    FootSpline* sfb = reinterpret_cast<FootSpline*>(smem_raw);   // the previous trajectory's splines
    FootSpline* sf = sfb + kNumEE;                               // trial splines
    double* zp = reinterpret_cast<double*>(sf + kNumEE);         // prev_qp_sol
    double* zq = zp + n_max;                                     // QP solution
    double* pd = zq + n_max;                                     // direction
    double* zt = pd + n_max;                                     // trial point
    double* xt = zt + n_max;                                     // round-tripped tangent states of the trial [N+1][12]
    double* red = xt + kNx * (L.N + 1);
    double* fp = red + 64;                                       // [N][4][6]: foot force and position at the node times (defects_of)
    __shared__ int s_fbase[kNumEE], s_pbase[kNumEE], s_nfv[kNumEE], s_npv[kNumEE];

    {
        const double* src = reinterpret_cast<const double*>(I.foot);
        double* dst = reinterpret_cast<double*>(sfb);
        for (int i = tid; i < static_cast<int>(sizeof(FootSpline) * kNumEE / 8); i += nth) dst[i] = src[i];
        for (int i = tid; i < n; i += nth) zp[i] = zprev_g[i];
        for (int i = tid; i < nu; i += nth) zq[ustart + i] = ustar[i];
        if (tid < kNumEE) {
            s_fbase[tid] = Hd->fbase[tid];
            s_pbase[tid] = Hd->pbase[tid];
            s_nfv[tid] = Hd->nfv[tid];
            s_npv[tid] = Hd->npv[tid];
        }
    }
    __syncthreads();
    const int status = Hd->status;

    // ---- sol: roll the states out through the linearised dynamics; Bd_k u from the structured pieces
    for (int k = tid; k < N; k += nth) {
        const NodeLin& nl = nodes[k];
        const double* u = zq + ustart;
        double bu[kNx];
        for (int i = 0; i < kNx; ++i) bu[i] = nl.cd[i];
        for (int e = 0; e < kNumEE; ++e)
            for (int c = 0; c < 3; ++c) {
                const double ec[3] = {c == 0 ? 1.0 : 0.0, c == 1 ? 1.0 : 0.0, c == 2 ? 1.0 : 0.0};
                if (nl.fcnt[e] > 0) {
                    const double* uf = u + s_fbase[e] + c * s_nfv[e] + nl.foff[e];
                    double fv = 0;
                    for (int j = 0; j < nl.fcnt[e]; ++j) fv += nl.fw[e][j] * uf[j];
                    double rc[3];
                    cross3f(nl.rel[e], ec, rc);
                    bu[3 + c] += P.dt * fv;
                    for (int r = 0; r < 3; ++r) bu[9 + r] += P.dt * rc[r] * fv;
                }
                if (c != 2) {
                    const double* up = u + nf + s_pbase[e] + c * s_npv[e] + nl.poff[e];
                    double pv = 0;
                    for (int j = 0; j < nl.pcnt[e]; ++j) pv += nl.pw[e][j] * up[j];
                    double ef[3];
                    cross3f(ec, nl.f[e], ef);
                    for (int r = 0; r < 3; ++r) bu[9 + r] += P.dt * ef[r] * pv;
                }
            }
        for (int i = 0; i < kNx; ++i) xt[k * kNx + i] = bu[i];   // stash Bd_k u + cd_k
    }
    if (tid < kNx) zq[tid] = xoff[tid];
    __syncthreads();
    if (tid < 32) {   // 12 lanes, one state row each; the next node's row of Ad travels from L2 while this node is multiplied
        double a_cur[kNx], a_next[kNx];
        if (tid < kNx)
            for (int q = 0; q < kNx; ++q) a_cur[q] = nodes[0].Ad[tid * kNx + q];
        for (int k = 0; k < N; ++k) {
            if (tid < kNx) {
                if (k + 1 < N)
                    for (int q = 0; q < kNx; ++q) a_next[q] = nodes[k + 1].Ad[tid * kNx + q];
                double s = xt[k * kNx + tid];
                for (int q = 0; q < kNx; ++q) s += a_cur[q] * zq[k * kNx + q];
                zq[(k + 1) * kNx + tid] = s;
                for (int q = 0; q < kNx; ++q) a_cur[q] = a_next[q];
            }
            __syncwarp();
        }
    }
    __syncthreads();
    // "Primal infeasible." is thrown and the previous solution reused (:115-129); the same when the solver broke down
    // before it had an iterate at all (NaN input, refused instance): there is nothing to step towards
    if (status == kPrimalInfeasible || Hd->no_iterate)
        for (int i = tid; i < n; i += nth) zq[i] = zp[i];
    __syncthreads();
    double sn = 0;
    for (int i = tid; i < n; i += nth) {
        pd[i] = zq[i] - zp[i];
        zqp_g[i] = zq[i];
        sn += pd[i] * pd[i];
    }
    const double step_norm = sqrt(block_reduce<kSum>(sn, red));

    // cost 1/2 z'Pz + q'z with the diagonal P of the assembly, evaluated at `z` with states taken from `xs`
    auto cost_of = [&](const double* z, const double* xs) -> double {
        double c = 0;
        for (int i = tid; i < n; i += nth) {
            double pii, qi, v;
            if (i < ustart) {
                const int k = i / kNx, r = i % kNx;
                pii = ((k < N) ? P.Q[r] : P.Phi[r]) + 1e-3;
                qi = (k < N) ? P.w[r] : P.Phi_w[r];
                v = xs[i];
            } else {
                pii = ((i - ustart < nf) ? P.force_cost : 0.0) + 1e-3;
                qi = 0.0;
                v = z[i];
            }
            c += 0.5 * pii * v * v + qi * v;
        }
        return block_reduce<kSum>(c, red);
    };
    // ConvertQPSolToTrajectory(z) into (sf, xt) and the l1 norm of the Euler defects
    auto defects_of = [&](const double* z, bool keep_manifold, double* man_out) -> double {
        {
            const double* src = reinterpret_cast<const double*>(sfb);
            double* dst = reinterpret_cast<double*>(sf);
            for (int i = tid; i < static_cast<int>(sizeof(FootSpline) * kNumEE / 8); i += nth) dst[i] = src[i];
        }
        __syncthreads();
        if (tid < kNumEE * 5) {
            const int e = tid / 5, c = tid % 5;
            if (c < 3) set_force_vars(sf[e], c, z + ustart + s_fbase[e] + c * s_nfv[e]);
            else set_pos_vars(sf[e], c - 3, z + ustart + nf + s_pbase[e] + (c - 3) * s_npv[e]);
        }
        for (int k = tid; k <= N; k += nth) {
            double q[4];
            quat_exp3(z + k * kNx + 6, q);
            quat_first_order_normalize(q);
            double* x = xt + k * kNx;
            for (int i = 0; i < 6; ++i) x[i] = z[k * kNx + i];
            quat_log3(q, x + 6);
            for (int i = 0; i < 3; ++i) x[9 + i] = z[k * kNx + 9 + i];
            if (keep_manifold) {
                double* mo = man_out + k * kNxMan;
                for (int i = 0; i < 6; ++i) mo[i] = z[k * kNx + i];
                for (int i = 0; i < 4; ++i) mo[6 + i] = q[i];
                for (int i = 0; i < 3; ++i) mo[10 + i] = z[k * kNx + 9 + i];
            }
        }
        __syncthreads();
        // the 24 spline evaluations of a node (4 feet x 3 coordinates x force / position) are spread over the CTA: one
        // work item per (node, foot, coordinate); the defects are then formed by one thread per node in the same order
        // of operations as before
        for (int it = tid; it < N * kNumEE * 3; it += nth) {
            const int k = it / (kNumEE * 3), e = (it / 3) % kNumEE, c = it % 3;
            const double tk = k * P.dt + t0;
            fp[(k * kNumEE + e) * 6 + c] = value_at(sf[e], true, c, tk);
            fp[(k * kNumEE + e) * 6 + 3 + c] = value_at(sf[e], false, c, tk);
        }
        __syncthreads();
        double l1 = 0;
        for (int k = tid; k < N; k += nth) {
            const double* x = xt + k * kNx;
            const double* xn = xt + (k + 1) * kNx;
            const double* om = x + 9;
            double fd[kNx];
            for (int i = 0; i < 3; ++i) fd[i] = x[3 + i] / P.mass;
            for (int i = 0; i < 3; ++i) fd[3 + i] = P.mass * P.gravity[i];
            for (int i = 0; i < 3; ++i) fd[6 + i] = P.Ir_inv[3 * i] * om[0] + P.Ir_inv[3 * i + 1] * om[1] + P.Ir_inv[3 * i + 2] * om[2];
            double Iw[3], wx[3];
            for (int i = 0; i < 3; ++i) Iw[i] = P.Ir[3 * i] * om[0] + P.Ir[3 * i + 1] * om[1] + P.Ir[3 * i + 2] * om[2];
            cross3f(om, Iw, wx);
            for (int i = 0; i < 3; ++i) fd[9 + i] = -wx[i];
            for (int e = 0; e < kNumEE; ++e) {
                double f[3], rel[3], tq[3];
                for (int c = 0; c < 3; ++c) {
                    f[c] = fp[(k * kNumEE + e) * 6 + c];
                    rel[c] = fp[(k * kNumEE + e) * 6 + 3 + c] - x[c];
                }
                cross3f(rel, f, tq);
                for (int i = 0; i < 3; ++i) {
                    fd[3 + i] += f[i];
                    fd[9 + i] += tq[i];
                }
            }
            for (int i = 0; i < kNx; ++i) l1 += fabs(xn[i] - (x[i] + P.dt * fd[i]));
        }
        return block_reduce<kSum>(l1, red);
    };

    // ---- MPC::LineSearch
    const double d0 = defects_of(zp, false, nullptr);
    const double merit0 = P.merit_mu * d0 + cost_of(zp, zp);   // GetCostValue(x) on the raw vector
    double gp = 0;   // (P z + q) . p at z = prev
    for (int i = tid; i < n; i += nth) {
        double pii, qi;
        if (i < ustart) {
            const int k = i / kNx, r = i % kNx;
            pii = ((k < N) ? P.Q[r] : P.Phi[r]) + 1e-3;
            qi = (k < N) ? P.w[r] : P.Phi_w[r];
        } else {
            pii = ((i - ustart < nf) ? P.force_cost : 0.0) + 1e-3;
            qi = 0.0;
        }
        gp += (pii * zp[i] + qi) * pd[i];
    }
    gp = block_reduce<kSum>(gp, red);
    const double merit_dd = gp - P.merit_mu * d0;
    double alpha = 1.0;
    int ls = 0;
    for (int i = tid; i < n; i += nth) zt[i] = alpha * pd[i] + zp[i];
    __syncthreads();
    double merit_step = P.merit_mu * defects_of(zt, false, nullptr);
    merit_step += cost_of(zt, zt);
    while ((merit0 - merit_step) < -0.00001 * alpha * merit_dd && ls < 10) {
        alpha *= 0.5;
        __syncthreads();
        for (int i = tid; i < n; i += nth) zt[i] = (alpha * pd[i]) + zp[i];
        __syncthreads();
        merit_step = P.merit_mu * defects_of(zt, false, nullptr);
        merit_step += cost_of(zt, zt);
        ls++;
    }

    // ---- prev_qp_sol += alpha p ; prev_traj_ = ConvertQPSolToTrajectory(prev_qp_sol)
    __syncthreads();
    for (int i = tid; i < n; i += nth) zt[i] = (alpha * pd[i]) + zp[i];
    __syncthreads();
    double* man = pd;   // pd is free now; 13(N+1) <= 12(N+1)+max_nu because max_nu >= N+1 (checked at bgg_create)
    const double dfin = defects_of(zt, true, man);
    const double cost_new = cost_of(zt, zt);      // GetCostValue(prev_qp_sol)
    const double cost_rt = cost_of(zt, xt);       // cost of ConvertTrajToQPVec(prev_traj_) used by the merit statistic
    for (int i = tid; i < (N + 1) * kNxMan; i += nth) I.states[i / kNxMan][i % kNxMan] = man[i];
    {
        double* dst = reinterpret_cast<double*>(I.foot);
        const double* src = reinterpret_cast<const double*>(sf);
        for (int i = tid; i < static_cast<int>(sizeof(FootSpline) * kNumEE / 8); i += nth) dst[i] = src[i];
    }
    for (int i = tid; i < n; i += nth) zprev_g[i] = zt[i];
    if (tid == 0) {
        // IncreaseEEBox / DecreaseEEBox
        if (status != kSolvedInacc && status != kSolved && status != kMaxIter) {
            I.ee_box[0] += 0.05;
            I.ee_box[1] += 0.05;
        } else {
            I.ee_box[0] = fmax(I.ee_box[0] - 0.05, P.ee_box_nominal[0]);
            I.ee_box[1] = fmax(I.ee_box[1] - 0.05, P.ee_box_nominal[1]);
        }
        I.run_count += 1;
        Hd->alpha = alpha;
        Hd->ls_iters = ls;
        Hd->eq_violation = dfin;
        Hd->step_norm = step_norm;
        Hd->cost = cost_new;
        Hd->merit = P.merit_mu * dfin + cost_rt;
        Hd->merit_dd = merit_dd;
        Hd->pass_state = 1;   // done: the second pass of this solve leaves the instance alone
    }
}

void launch_finish(const Params& P, Instance* inst, const WsLayout& L, char* ws, int B, int want, cudaStream_t stream) {
    const size_t smem = finish_smem_bytes(L);
    // per device and context, so set on every launch (see launch_ipm)
    cudaFuncSetAttribute(k_finish, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    k_finish<<<B, 128, smem, stream>>>(P, inst, L, ws, want);
}

}  // namespace bgg
