// bilevel-gait-gen_b200 -- kernel 6: derivative of the MPC cost with respect to the contact times, one CTA per MPC
// instance.  Stands in for the reference's whole derivative chain of MPCController::GaitOpt
// (controllers/mpc_controller.cpp:518-573):
//   MPC::ComputeDerivativeTerms -> ClarabelInterface::Computedx / SetupDerivativeCalcs   (mpc.cpp:1047-1057,
//                                                                      clarabel_interface.cpp:604-612, 262-602)
//   MPC::GetQPPartials -> CalcDerivativeWrtMats / CalcDerivativeWrtVecs                   (clarabel_interface.cpp:182-260)
//   MPCSingleRigidBody::ComputeParamPartialsClarabel for every (foot, contact time)        (mpc_single_rigid_body.cpp:642-792,
//                                                                      mpc.cpp:240-350, 416-531; single_rigid_body_model.cpp:458-555)
//   GaitOptimizer::ModifyQPPartials / ComputeCostFcnDerivWrtContactTimes                   (gait_optimizer.cpp:536-539, 92-179)
//
// (a) Adjoint.  The reference solves the (n+m) x (n+m) sparse system
//        [ P   G'D(lam)  A' ] [dz  ]     [P z + q]
//        [ G   D(s)      0  ] [dlam] = - [   0   ]        z = prev_qp_sol, (lam, s) = the solver's duals / slacks
//        [ A   0         0  ] [dnu ]     [   0   ]
//     by sparse LU (1384^2 at N = 20).  Here dlam = -G dz / s is eliminated, the dynamics rows A dz = 0 make the state
//     part of dz the linearised rollout of its spline part du, and what is left is the nu x nu system
//        (H - C'WC) du + E' dnu_E = -Phi~' (P z + q),  E du = 0,   W = lam / s,
//     assembled by the same owner-computes routine the interior-point kernel uses (csrc/bgg_kkt.cuh, sign = -1) and
//     solved by an in-shared-memory LU with partial pivoting (the matrix is indefinite: the reference's D(s) block has
//     the sign of the slack, not of G z - h, and that is kept).  Products with Phi~' and the dynamics multipliers dnu, nu
//     come from 12-wide adjoint recursions through Ad_k instead of from stored matrices.
// (b) Contraction.  The reference forms the dense rank-2 matrices dA = dnu z*' + nu dz' (260 x 372) and
//     dG = D(lam) dlam z*' + lam dz' (752 x 372) and multiplies them entry-wise with each contact time's sparse parameter
//     partial.  <dA, X> = dnu'(X z*) + nu'(X dz), so each partial is applied to the two vectors as it is generated and
//     never stored: one warp per contact time, lanes over MPC nodes and constraint samples.
#include "bgg_kernels.cuh"
#include "bgg_kkt.cuh"

namespace bgg {

namespace {

__device__ __forceinline__ void cross3g(const double a[3], const double b[3], double o[3]) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}

struct GradCaps {
    int nu, rows;
};

size_t grad_smem_for(int N, const GradCaps& c) {
    const size_t eb = 4 * (N - 3), nkc = 2 * (N - 3);
    return 8 * (static_cast<size_t>(c.nu) * c.nu        // K, dense
                + 4 * static_cast<size_t>(c.nu)          // rhs / du, ustar, utraj, scratch
                + static_cast<size_t>(c.rows)            // wv -> lam * dlam
                + 3 * static_cast<size_t>(kNx) * (N + 1) // adjoint recursions: mu / dnu, nu, dxs
                + 2 * eb + nkc + 2 * kMaxEq + 64 + (1 + kMaxEq) * static_cast<size_t>(c.nu) + kMaxEq * (kMaxEq + 1) + kMaxEq * kMaxEq + kMaxEq + static_cast<size_t>(c.rows) + static_cast<size_t>(c.nu)) +
           4 * (2 * eb + static_cast<size_t>(c.nu)) + sizeof(ColInfo) * static_cast<size_t>(c.nu) + 256;
}

}  // namespace

__global__ void __launch_bounds__(256, 1) k_gradient(Params P, const Instance* __restrict__ inst, WsLayout L, char* __restrict__ ws_base,
                                                      int cap_nu, int cap_rows) {
    const int b = blockIdx.x, tid = threadIdx.x, nth = blockDim.x;
    const int lane = tid & 31, wid = tid >> 5, nwarp = nth >> 5;
    const Instance& I = inst[b];
    char* ws = ws_base + static_cast<size_t>(b) * L.stride;
    const WsHeader* Hd = reinterpret_cast<const WsHeader*>(ws + L.hdr);
    GradInfo* gi = reinterpret_cast<GradInfo*>(ws + L.ginfo);
    double* gdH = reinterpret_cast<double*>(ws + L.gdH);
    if (Hd->error || Hd->status != kSolved) {   // MPC::ComputeDerivativeTerms returns false unless the last solve was Solved
        if (tid == 0) {
            gi->status = 1;
            gi->n_theta = 0;
            for (int e = 0; e < kNumEE; ++e) gi->nct[e] = 0;
        }
        for (int i = tid; i < kNumEE * kMaxContacts; i += nth) gdH[i] = 0.0;
        return;
    }
    if (Hd->nu > cap_nu || 6 * Hd->n_samples + 2 * Hd->n_eebox > cap_rows) {   // larger than the shared memory this launch was sized for
        if (tid == 0) {
            gi->status = 2;
            gi->n_theta = 0;
            for (int e = 0; e < kNumEE; ++e) gi->nct[e] = 0;
        }
        for (int i = tid; i < kNumEE * kMaxContacts; i += nth) gdH[i] = 0.0;
        return;
    }
    const NodeLin* nodes = reinterpret_cast<const NodeLin*>(ws + L.nodes);
    const Sample* samples = reinterpret_cast<const Sample*>(ws + L.samples);
    const EqRow* eqs = reinterpret_cast<const EqRow*>(ws + L.eq);
    const double* Hg = reinterpret_cast<const double*>(ws + L.H);
    const double* phipos = reinterpret_cast<const double*>(ws + L.phipos);
    const double* znew = reinterpret_cast<const double*>(ws + L.zprev);   // prev_qp_sol after the line search
    const double* zqp = reinterpret_cast<const double*>(ws + L.zqp);      // the QP optimum (ClarabelInterface::primal_)
    const double* lam_g = reinterpret_cast<const double*>(ws + L.lam);
    const double* slack_g = reinterpret_cast<const double*>(ws + L.slack);
    const double* nueq_g = reinterpret_cast<const double*>(ws + L.nueq);
    double* gdx = reinterpret_cast<double*>(ws + L.gdx);
    double* gdz = reinterpret_cast<double*>(ws + L.gdz);
    double* gdlam = reinterpret_cast<double*>(ws + L.gdlam);
    double* gdnu = reinterpret_cast<double*>(ws + L.gdnu);
    double* gdnue = reinterpret_cast<double*>(ws + L.gdnue);
    double* gnu = reinterpret_cast<double*>(ws + L.dualx);

    const int N = P.N, nu = Hd->nu, nf = Hd->nf, ns = Hd->n_samples, ne = Hd->n_eebox, neq = Hd->n_eq, ntd = Hd->n_td;
    const int m = 6 * ns + 2 * ne, m_force = 6 * ns, nkc = 2 * (N - 3), ustart = kNx * (N + 1), n = Hd->n;
    const double dt = P.dt, t0 = Hd->t0, mu_f = P.friction_coef;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* p = reinterpret_cast<double*>(smem_raw);
    double* K = p; p += static_cast<size_t>(cap_nu) * cap_nu;
    double* du = p; p += cap_nu;          // right-hand side, then the solution
    double* us = p; p += cap_nu;          // u of the QP optimum
    double* ut = p; p += cap_nu;          // SplinesAsVec of the updated trajectory
    double* rhs0 = p; p += cap_nu;        // right-hand side kept for the refinement step
    double* wv = p; p += cap_rows;        // W = lam / s, later y = lam * dlam
    double* rec = p; p += kNx * (N + 1);  // mu_k, later dnu_k
    double* nus = p; p += kNx * (N + 1);  // nu_k
    double* dxs = p; p += kNx * (N + 1);  // state part of dz
    double* s_pw = p; p += 2 * 4 * (N - 3);
    double* ckc = p; p += nkc;
    double* dnue = p; p += kMaxEq;
    double* nue = p; p += kMaxEq;
    double* red = p; p += 64;
    double* R = p; p += static_cast<size_t>(1 + kMaxEq) * cap_nu;   // right-hand sides [1 + neq][nu]
    double* Sm = p; p += kMaxEq * (kMaxEq + 1);
    double* Sinv = p; p += kMaxEq * kMaxEq;
    double* sig = p; p += kMaxEq;
    double* gv = p; p += cap_rows;        // G dz
    double* tv = p; p += cap_nu;
    int* s_pcnt = reinterpret_cast<int*>(p);
    int* s_poff = s_pcnt + 4 * (N - 3);
    int* piv = s_poff + 4 * (N - 3);
    ColInfo* col = reinterpret_cast<ColInfo*>(piv + cap_nu + (cap_nu & 1));
    __shared__ FootSpline sf[kNumEE];
    __shared__ int s_fbase[kNumEE], s_pbase[kNumEE], s_nfv[kNumEE], s_npv[kNumEE], s_sb[kNumEE + 1], s_nct[kNumEE + 1];
    __shared__ int s_ctk[kNumEE][kMaxContacts];   // knot index of each contact time
    __shared__ int s_flag;
    __shared__ double s_pivmin;

    // ------------------------------------------------------------------------------------------------ staging
    {
        const double* src = reinterpret_cast<const double*>(I.foot);
        double* dst = reinterpret_cast<double*>(sf);
        for (int i = tid; i < static_cast<int>(sizeof(FootSpline) * kNumEE / 8); i += nth) dst[i] = src[i];
    }
    if (tid < kNumEE) {
        s_fbase[tid] = Hd->fbase[tid];
        s_pbase[tid] = Hd->pbase[tid];
        s_nfv[tid] = Hd->nfv[tid];
        s_npv[tid] = Hd->npv[tid];
    }
    if (tid == 0) {
        int e = 0;
        s_sb[0] = 0;
        for (int j = 0; j < ns; ++j)
            while (samples[j].ee > e) s_sb[++e] = j;
        while (e < kNumEE) s_sb[++e] = ns;
        s_flag = 0;
        s_pivmin = 1e300;
    }
    for (int i = tid; i < 4 * (N - 3); i += nth) {
        const NodeLin& nl = nodes[i / 4 + kEENodeStart];
        const int foot = i & 3;
        s_pcnt[i] = nl.pcnt[foot];
        s_poff[i] = nl.poff[foot];
        s_pw[2 * i] = nl.pw[foot][0];
        s_pw[2 * i + 1] = nl.pw[foot][1];
    }
    for (int i = tid; i < m; i += nth) wv[i] = (lam_g[i] > 0.0) ? lam_g[i] / slack_g[i] : 0.0;   // rows kept out of the solve: lam == 0
    for (int i = tid; i < nu; i += nth) us[i] = zqp[ustart + i];
    if (tid < kMaxEq) nue[tid] = (tid < neq) ? nueq_g[tid] : 0.0;
    // dx = P z + q at prev_qp_sol (Computedx).  P is diagonal.
    for (int j = tid; j < n; j += nth) {
        const int k = j / kNx, i = j % kNx;
        double pv, qv;
        if (k < N) { pv = P.Q[i] + 1e-3; qv = P.w[i]; }
        else if (k == N) { pv = P.Phi[i] + 1e-3; qv = P.Phi_w[i]; }
        else { pv = ((j - ustart < nf) ? P.force_cost : 0.0) + 1e-3; qv = 0.0; }
        gdx[j] = pv * znew[j] + qv;
    }
    __syncthreads();
    if (tid < kNumEE) {   // contact times of the updated trajectory and SplinesAsVec
        const FootSpline& s = sf[tid];
        int c = 0;
        for (int i = 0; i < s.n && c < kMaxContacts; ++i)
            if (s.ttype[i] != kInter) s_ctk[tid][c++] = i;
        s_nct[tid] = c;
    }
    for (int i = tid; i < kNumEE * 5; i += nth) {
        const int e = i / 5, c = i % 5;
        if (c < 3) get_force_vars(sf[e], c, ut + s_fbase[e] + c * s_nfv[e]);
        else get_pos_vars(sf[e], c - 3, ut + nf + s_pbase[e] + (c - 3) * s_npv[e]);
    }
    kkt_build_colinfo(col, nu, nf, N, s_fbase, s_pbase, s_nfv, s_npv, s_sb, samples, s_pcnt, s_poff);
    __syncthreads();

    // y += Ad_k' x, 12 lanes of one warp (lane i owns component i); Ad row-major in HBM / L2
    auto adT = [&](const NodeLin& nl, const double* x, int i) {
        double s = 0;
#pragma unroll
        for (int r = 0; r < kNx; ++r) s += nl.Ad[r * kNx + i] * x[r];
        return s;
    };
    // (Bd_k' x)[column i of u]  (single_rigid_body_model.cpp:113-148 transposed)
    auto bdT_col = [&](int i, const double* xall /* [N+1][12], uses x_{k+1} */) {
        const ColInfo ci = col[i];
        const int e = ci.foot, c = ci.coord;
        const double ec[3] = {c == 0 ? 1.0 : 0.0, c == 1 ? 1.0 : 0.0, c == 2 ? 1.0 : 0.0};
        double s = 0;
        for (int k = 0; k < N; ++k) {
            const NodeLin& nl = nodes[k];
            const double* x = xall + (k + 1) * kNx;
            if (i < nf) {
                const int a = ci.var - nl.foff[e];
                if (a < 0 || a >= nl.fcnt[e]) continue;
                double rc[3];
                cross3g(nl.rel[e], ec, rc);
                s += dt * nl.fw[e][a] * (x[3 + c] + rc[0] * x[9] + rc[1] * x[10] + rc[2] * x[11]);
            } else {
                const int a = ci.var - nl.poff[e];
                if (a < 0 || a >= nl.pcnt[e]) continue;
                double ef[3];
                cross3g(ec, nl.f[e], ef);
                s += dt * nl.pw[e][a] * (ef[0] * x[9] + ef[1] * x[10] + ef[2] * x[11]);
            }
        }
        return s;
    };
    // (Bd_k v)[12] for a spline vector v
    auto bd_apply = [&](const NodeLin& nl, const double* v, double out[kNx]) {
        for (int i = 0; i < kNx; ++i) out[i] = 0.0;
        for (int e = 0; e < kNumEE; ++e)
            for (int c = 0; c < 3; ++c) {
                const double ec[3] = {c == 0 ? 1.0 : 0.0, c == 1 ? 1.0 : 0.0, c == 2 ? 1.0 : 0.0};
                if (nl.fcnt[e] > 0) {
                    const double* uf = v + s_fbase[e] + c * s_nfv[e] + nl.foff[e];
                    double fv = 0;
                    for (int j = 0; j < nl.fcnt[e]; ++j) fv += nl.fw[e][j] * uf[j];
                    double rc[3];
                    cross3g(nl.rel[e], ec, rc);
                    out[3 + c] += dt * fv;
                    for (int r = 0; r < 3; ++r) out[9 + r] += dt * rc[r] * fv;
                }
                if (c != 2) {
                    const double* up = v + nf + s_pbase[e] + c * s_npv[e] + nl.poff[e];
                    double pv = 0;
                    for (int j = 0; j < nl.pcnt[e]; ++j) pv += nl.pw[e][j] * up[j];
                    double ef[3];
                    cross3g(ec, nl.f[e], ef);
                    for (int r = 0; r < 3; ++r) out[9 + r] += dt * ef[r] * pv;
                }
            }
    };

    // ------------------------------------------------------------------------------------------------ (a) adjoint
    // mu_N = dx_N, mu_k = dx_k + Ad_k' mu_{k+1};  Phi~' dx = dx_u + sum_k Bd_k' mu_{k+1}
    if (wid == 0) {
        if (lane < kNx) rec[N * kNx + lane] = gdx[N * kNx + lane];
        __syncwarp();
        for (int k = N - 1; k >= 0; --k) {
            double s = 0;
            if (lane < kNx) s = gdx[k * kNx + lane] + adT(nodes[k], rec + (k + 1) * kNx, lane);
            __syncwarp();
            if (lane < kNx) rec[k * kNx + lane] = s;
            __syncwarp();
        }
    }
    __syncthreads();
    for (int i = tid; i < nu; i += nth) {
        const double r = gdx[ustart + i] + bdT_col(i, rec);
        du[i] = -r;
        rhs0[i] = -r;
    }
    // K0 = H - C'WC, dense
    KktView kv;
    kv.K = K; kv.ld = nu; kv.Hg = Hg; kv.nu = nu; kv.nf = nf; kv.N = N; kv.ns = ns; kv.ne = ne; kv.neq = neq; kv.nkc = nkc;
    kv.wv = wv; kv.phi = phipos; kv.phi_stride = L.max_nu; kv.pw = s_pw; kv.pcnt = s_pcnt; kv.poff = s_poff;
    kv.smp = samples; kv.eq = eqs; kv.col = col; kv.ckc = ckc; kv.mu_f = mu_f; kv.inv_delta = 0.0; kv.sign = -1.0; kv.tile = nullptr;
    kkt_assemble<false>(kv, s_fbase, s_nfv);
    for (int idx = tid; idx < nu * nu; idx += nth) {   // mirror the lower triangle
        const int i = idx / nu, j = idx % nu;
        if (j > i) K[idx] = K[j * nu + i];
    }
    __syncthreads();
    // The touch-down / foot-start rows E du = 0 are imposed exactly (a delta-penalty on them is 14x off at delta = 1e-8:
    // their multipliers are large and the cost is flat in most spline directions):  with K0 = H - C'WC,
    //   K0 X = [rhs | E']  ->  S = E X_E,  S dnu_E = E x_0,  du = x_0 - X_E dnu_E.
    // LU with partial pivoting, right-looking, in place; the 1 + neq right-hand sides are permuted and eliminated along.
    const int nrhs = 1 + neq;
    for (int idx = tid; idx < neq * nu; idx += nth) R[nu + idx] = 0.0;
    for (int i = tid; i < nu; i += nth) R[i] = du[i];
    __syncthreads();
    if (tid < neq) {
        const EqRow& q = eqs[tid];
        for (int i = 0; i < q.cnt; ++i) R[(1 + tid) * nu + q.col[i]] = q.w[i];
    }
    __syncthreads();
    for (int k = 0; k < nu; ++k) {
        if (wid == 0) {
            double best = -1.0;
            int bi = k;
            for (int i = k + lane; i < nu; i += 32) {
                const double a = fabs(K[i * nu + k]);
                if (a > best) { best = a; bi = i; }
            }
            for (int o = 16; o > 0; o >>= 1) {
                const double ob = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
            }
            if (lane == 0) {
                piv[k] = bi;
                if (!(best > 0.0)) s_flag = 1;
                if (best < s_pivmin) s_pivmin = best;
            }
        }
        __syncthreads();
        const int pk_ = piv[k];
        if (pk_ != k) {
            for (int j = tid; j < nu; j += nth) {
                const double t = K[k * nu + j];
                K[k * nu + j] = K[pk_ * nu + j];
                K[pk_ * nu + j] = t;
            }
            if (tid < nrhs) {
                const double t = R[tid * nu + k];
                R[tid * nu + k] = R[tid * nu + pk_];
                R[tid * nu + pk_] = t;
            }
        }
        __syncthreads();
        const double dkk = K[k * nu + k];
        const double inv = (dkk != 0.0) ? 1.0 / dkk : 0.0;
        for (int i = k + 1 + tid; i < nu; i += nth) K[i * nu + k] *= inv;
        __syncthreads();
        const int rem = nu - k - 1;
        if (rem > 0) {
            // trailing update: thread -> (row, 8-column strip) so the multiplier and the index arithmetic are shared
            const int strips = (rem + 7) >> 3;
            for (int idx = tid; idx < rem * strips; idx += nth) {
                const int i = k + 1 + idx / strips, j0 = k + 1 + 8 * (idx % strips);
                const double l = K[i * nu + k];
                const int j1 = (j0 + 8 < nu) ? j0 + 8 : nu;
                for (int j = j0; j < j1; ++j) K[i * nu + j] -= l * K[k * nu + j];
            }
            for (int idx = tid; idx < rem * nrhs; idx += nth) {
                const int c = idx / rem, i = k + 1 + idx % rem;
                R[c * nu + i] -= K[i * nu + k] * R[c * nu + k];
            }
        }
        __syncthreads();
    }
    // back substitution U x = y, one warp per right-hand side
    for (int c = wid; c < nrhs; c += nwarp) {
        double* x = R + c * nu;
        for (int k = nu - 1; k >= 0; --k) {
            double sacc = 0;
            for (int j = k + 1 + lane; j < nu; j += 32) sacc += K[k * nu + j] * x[j];
            sacc = warp_sum(sacc);
            if (lane == 0) x[k] = (x[k] - sacc) / K[k * nu + k];
            __syncwarp();
        }
    }
    __syncthreads();
    // Schur complement of the equality rows: S [neq][neq], right-hand side t = E x_0
    if (tid < neq * nrhs) {
        const int r = tid / nrhs, c = tid % nrhs;
        const EqRow& q = eqs[r];
        double v = 0;
        for (int i = 0; i < q.cnt; ++i) v += q.w[i] * R[c * nu + q.col[i]];
        Sm[r * (kMaxEq + 1) + c] = v;   // column 0: t, columns 1..neq: S
    }
    __syncthreads();
    if (tid == 0) {   // Gauss-Jordan with partial pivoting: Sinv = S^-1 (kept for the refinement steps), dnu_E = Sinv t
        const int ld = kMaxEq + 1;
        for (int r = 0; r < neq; ++r)
            for (int c = 0; c < neq; ++c) Sinv[r * kMaxEq + c] = (r == c) ? 1.0 : 0.0;
        for (int k = 0; k < neq; ++k) {
            int pr = k;
            double best = fabs(Sm[k * ld + 1 + k]);
            for (int r = k + 1; r < neq; ++r)
                if (fabs(Sm[r * ld + 1 + k]) > best) { best = fabs(Sm[r * ld + 1 + k]); pr = r; }
            if (!(best > 0.0)) s_flag = 1;
            if (pr != k) {
                for (int c = 0; c <= neq; ++c) { const double t = Sm[k * ld + c]; Sm[k * ld + c] = Sm[pr * ld + c]; Sm[pr * ld + c] = t; }
                for (int c = 0; c < neq; ++c) { const double t = Sinv[k * kMaxEq + c]; Sinv[k * kMaxEq + c] = Sinv[pr * kMaxEq + c]; Sinv[pr * kMaxEq + c] = t; }
            }
            const double inv = 1.0 / Sm[k * ld + 1 + k];
            for (int c = 0; c <= neq; ++c) Sm[k * ld + c] *= inv;
            for (int c = 0; c < neq; ++c) Sinv[k * kMaxEq + c] *= inv;
            for (int r = 0; r < neq; ++r) {
                if (r == k) continue;
                const double l = Sm[r * ld + 1 + k];
                if (l == 0.0) continue;
                for (int c = 0; c <= neq; ++c) Sm[r * ld + c] -= l * Sm[k * ld + c];
                for (int c = 0; c < neq; ++c) Sinv[r * kMaxEq + c] -= l * Sinv[k * kMaxEq + c];
            }
        }
        for (int k = 0; k < kMaxEq; ++k) dnue[k] = (k < neq) ? Sm[k * ld] : 0.0;
    }
    __syncthreads();
    for (int i = tid; i < nu; i += nth) {
        double v = R[i];
        for (int r = 0; r < neq; ++r) v -= R[(1 + r) * nu + i] * dnue[r];
        du[i] = v;
    }
    if (tid < neq) gdnue[tid] = dnue[tid];
    __syncthreads();

    // ---- operators of the condensed problem (every thread calls them; each ends with a barrier)
    // xs = Phi v : xs_0 = 0, xs_{k+1} = Ad_k xs_k + Bd_k v
    auto rollout = [&](const double* v, double* xs) {
        for (int k = tid; k < N; k += nth) {
            double o[kNx];
            bd_apply(nodes[k], v, o);
            for (int i = 0; i < kNx; ++i) xs[(k + 1) * kNx + i] = o[i];   // stash Bd_k v
        }
        if (tid < kNx) xs[tid] = 0.0;
        __syncthreads();
        if (wid == 0) {
            for (int k = 0; k < N; ++k) {
                double sacc = 0;
                if (lane < kNx) {
                    sacc = xs[(k + 1) * kNx + lane];
                    for (int q = 0; q < kNx; ++q) sacc += nodes[k].Ad[lane * kNx + q] * xs[k * kNx + q];
                }
                __syncwarp();
                if (lane < kNx) xs[(k + 1) * kNx + lane] = sacc;
                __syncwarp();
            }
        }
        __syncthreads();
    };
    // g = G [xs ; v], inequality rows in the kernel's order
    auto apply_G = [&](const double* v, const double* xs, double* g) {
        for (int j = tid; j < ns; j += nth) {
            const Sample& sp = samples[j];
            double fv[3];
            for (int c = 0; c < 3; ++c) {
                const double* vv = v + s_fbase[sp.ee] + c * s_nfv[sp.ee] + sp.off;
                double sacc = 0;
                for (int i = 0; i < sp.cnt; ++i) sacc += sp.w[i] * vv[i];
                fv[c] = sacc;
            }
            double* o = g + 6 * j;
            o[0] = fv[2];
            o[1] = -fv[2];
            o[2] = fv[0] - mu_f * fv[2];
            o[3] = -fv[0] - mu_f * fv[2];
            o[4] = fv[1] - mu_f * fv[2];
            o[5] = -fv[1] - mu_f * fv[2];
        }
        for (int e = tid; e < ne; e += nth) {
            const int c = e & 1, foot = (e >> 1) & 3, kk = e >> 3, kf = kk * 4 + foot;
            const double* vv = v + nf + s_pbase[foot] + c * s_npv[foot] + s_poff[kf];
            double sacc = -xs[(kk + kEENodeStart) * kNx + c];
            for (int i = 0; i < s_pcnt[kf]; ++i) sacc += s_pw[2 * kf + i] * vv[i];
            g[m_force + 2 * e] = sacc;
            g[m_force + 2 * e + 1] = -sacc;
        }
        __syncthreads();
    };
    // out = Phi~' G' y : state columns through the adjoint recursion (scratch mu [N+1][12]), spline columns directly
    auto apply_Gt = [&](const double* y, double* mu, double* out) {
        if (wid == 0) {
            for (int k = N; k >= 0; --k) {
                double sacc = 0;
                if (lane < kNx) {
                    if (lane < 2 && k >= kEENodeStart)
                        for (int foot = 0; foot < kNumEE; ++foot) {
                            const int row = m_force + 2 * (((k - kEENodeStart) * 4 + foot) * 2 + lane);
                            sacc += -y[row] + y[row + 1];
                        }
                    if (k < N) sacc += adT(nodes[k], mu + (k + 1) * kNx, lane);
                }
                __syncwarp();
                if (lane < kNx) mu[k * kNx + lane] = sacc;
                __syncwarp();
            }
        }
        __syncthreads();
        for (int i = tid; i < nu; i += nth) {
            double sacc = bdT_col(i, mu);
            const ColInfo ci = col[i];
            if (i < nf) {
                for (int sidx = ci.lo; sidx < ci.hi; ++sidx) {
                    const Sample& sp = samples[sidx];
                    const double* yy = y + 6 * sidx;
                    double coef;
                    if (ci.coord == 2) coef = (yy[0] - yy[1]) - mu_f * (yy[2] + yy[3] + yy[4] + yy[5]);
                    else if (ci.coord == 0) coef = yy[2] - yy[3];
                    else coef = yy[4] - yy[5];
                    sacc += coef * sp.w[ci.var - sp.off];
                }
            } else {
                for (int kk = ci.lo; kk < ci.hi; ++kk) {
                    const int kf = kk * kNumEE + ci.foot, e = kf * 2 + ci.coord;
                    sacc += (y[m_force + 2 * e] - y[m_force + 2 * e + 1]) * s_pw[2 * kf + (ci.var - s_poff[kf])];
                }
            }
            out[i] = sacc;
        }
        __syncthreads();
    };

    // ---- iterative refinement of [K0 E'; E 0] [du; dnu_E] = [rhs; 0] against the matrix-free operator
    // K0 v = H v - Phi~' G' (W o (G Phi~ v)) (the factorised normal-equations matrix mixes entries of 1e-3 and 1e8)
    for (int rf = 0; rf < 2; ++rf) {
        rollout(du, dxs);
        apply_G(du, dxs, gv);
        for (int i = tid; i < m; i += nth) gv[i] *= wv[i];
        __syncthreads();
        apply_Gt(gv, nus, tv);
        double* rho = R;   // column 0 of R is free once du has been formed
        double rn = 0, bn = 0;
        for (int i = tid; i < nu; i += nth) {
            double hv = 0;
            for (int j = 0; j < nu; ++j) hv += Hg[static_cast<size_t>(j) * nu + i] * du[j];
            double r = rhs0[i] - hv + tv[i];
            const ColInfo ci = col[i];
            if (i >= nf)
                for (int q = 0; q < neq; ++q) {
                    const EqRow& er = eqs[q];
                    const int a = i - er.col[0];
                    if (er.pad == ci.foot * 2 + ci.coord && a >= 0 && a < er.cnt) r -= er.w[a] * dnue[q];
                }
            rho[i] = r;
            rn = fmax(rn, fabs(r));
            bn = fmax(bn, fabs(rhs0[i]));
        }
        rn = block_reduce<kMax>(rn, red);
        bn = block_reduce<kMax>(bn, red);
        if (tid == 0) gi->resid = rn / fmax(bn, 1e-300);
        if (tid < kMaxEq) {   // sigma = -E du
            double v = 0;
            if (tid < neq) {
                const EqRow& q = eqs[tid];
                for (int i = 0; i < q.cnt; ++i) v -= q.w[i] * du[q.col[i]];
            }
            sig[tid] = v;
        }
        __syncthreads();
        if (wid == 0) {   // x = K0^-1 rho with the stored factors: P, L, then U
            if (lane == 0)   // all row interchanges first (the stored multipliers are in fully permuted order), then L, then U
                for (int k = 0; k < nu; ++k) {
                    const int pk_ = piv[k];
                    if (pk_ != k) {
                        const double t = rho[k];
                        rho[k] = rho[pk_];
                        rho[pk_] = t;
                    }
                }
            __syncwarp();
            for (int k = 0; k < nu; ++k) {
                const double xk = rho[k];
                for (int i = k + 1 + lane; i < nu; i += 32) rho[i] -= K[i * nu + k] * xk;
                __syncwarp();
            }
            for (int k = nu - 1; k >= 0; --k) {
                double sacc = 0;
                for (int j = k + 1 + lane; j < nu; j += 32) sacc += K[k * nu + j] * rho[j];
                sacc = warp_sum(sacc);
                if (lane == 0) rho[k] = (rho[k] - sacc) / K[k * nu + k];
                __syncwarp();
            }
        }
        __syncthreads();
        if (tid < kMaxEq) {   // d(dnu_E) = Sinv (E x - sigma)
            double v = 0;
            if (tid < neq) {
                const EqRow& q = eqs[tid];
                for (int i = 0; i < q.cnt; ++i) v += q.w[i] * rho[q.col[i]];
                v -= sig[tid];
            }
            Sm[tid] = v;
        }
        __syncthreads();
        if (tid < kMaxEq) {
            double v = 0;
            if (tid < neq)
                for (int c = 0; c < neq; ++c) v += Sinv[tid * kMaxEq + c] * Sm[c];
            sig[tid] = v;
        }
        __syncthreads();
        for (int i = tid; i < nu; i += nth) {
            double v = rho[i];
            for (int r = 0; r < neq; ++r) v -= R[(1 + r) * nu + i] * sig[r];
            du[i] += v;
        }
        if (tid < neq) dnue[tid] += sig[tid];
        __syncthreads();
    }
    if (tid < neq) gdnue[tid] = dnue[tid];

    // state part of dz, dlam = -(G dz) / s and y = lam * dlam (kept in wv)
    rollout(du, dxs);
    apply_G(du, dxs, gv);
    for (int i = tid; i < ustart; i += nth) gdz[i] = dxs[i];
    for (int i = tid; i < nu; i += nth) gdz[ustart + i] = du[i];
    for (int i = tid; i < m; i += nth) {
        const double dl = (lam_g[i] > 0.0) ? -gv[i] / slack_g[i] : 0.0;
        gdlam[i] = dl;
        wv[i] = lam_g[i] * dl;
    }
    __syncthreads();
    // dnu_N = r_N, dnu_k = r_k + Ad_k' dnu_{k+1},  r_k = dx_k + P_k dxs_k + (G'(lam dlam))_xk   (first block row)
    // nu_N  = rho_N, nu_k = rho_k + Ad_k' nu_{k+1}, rho_k = P_k z*_k + q_k + (G' lam)_xk          (stationarity of the QP)
    if (wid < 2) {
        double* out = wid == 0 ? rec : nus;
        for (int k = N; k >= 0; --k) {
            double s = 0;
            if (lane < kNx) {
                const double pv = ((k < N) ? P.Q[lane] : P.Phi[lane]) + 1e-3;
                if (wid == 0) s = gdx[k * kNx + lane] + pv * dxs[k * kNx + lane];
                else s = pv * zqp[k * kNx + lane] + ((k < N) ? P.w[lane] : P.Phi_w[lane]);
                if (lane < 2 && k >= kEENodeStart)
                    for (int foot = 0; foot < kNumEE; ++foot) {
                        const int row = m_force + 2 * (((k - kEENodeStart) * 4 + foot) * 2 + lane);
                        s += (wid == 0) ? (-wv[row] + wv[row + 1]) : (-lam_g[row] + lam_g[row + 1]);
                    }
                if (k < N) s += adT(nodes[k], out + (k + 1) * kNx, lane);
            }
            __syncwarp();
            if (lane < kNx) out[k * kNx + lane] = s;
            __syncwarp();
        }
    }
    __syncthreads();
    for (int i = tid; i < ustart; i += nth) {
        gdnu[i] = rec[i];
        gnu[i] = nus[i];
    }

    // ------------------------------------------------------------------------------------------------ (b) contraction
    // theta = (foot ee, contact idx), foot-major (GaitOptimizer::GetNumTimeNodes); one warp per theta
    if (tid == 0) {
        int tot = 0;
        for (int e = 0; e < kNumEE; ++e) tot += s_nct[e];
        gi->n_theta = tot;
        for (int e = 0; e < kNumEE; ++e) gi->nct[e] = s_nct[e];
        gi->status = s_flag ? 2 : 0;
        gi->pivot_min = s_pivmin;
    }
    const int n_theta = s_nct[0] + s_nct[1] + s_nct[2] + s_nct[3];
    for (int th = wid; th < n_theta; th += nwarp) {
        int ee = 0, idx = th;
        while (idx >= s_nct[ee]) { idx -= s_nct[ee]; ++ee; }
        const FootSpline& s = sf[ee];
        const int fb = s_fbase[ee], nv = s_nfv[ee], pb = nf + s_pbase[ee], npv = s_npv[ee];
        double acc = 0.0;
        // --- dynamics rows (model partials, single_rigid_body_model.cpp:458-555) and foot-box rows, lanes over nodes
        for (int k = lane; k <= N; k += 32) {
            const double tk = k * dt + t0;
            double dwp[2], wp[2];
            int poff_;
            const int pcnt_ = pos_coef_partial(s, tk, idx, dwp);
            pos_lin(s, tk, wp, &poff_);
            if (pcnt_ < 2) wp[1] = 0.0;
            if (k < N) {
                double fp[3], pp[3] = {0, 0, 0}, f[3], rel[3];
                for (int c = 0; c < 3; ++c) {
                    fp[c] = partial_wrt_time(s, true, c, tk, idx);
                    f[c] = value_at(s, true, c, tk);
                    rel[c] = value_at(s, false, c, tk) - I.states[k][c];
                }
                for (int c = 0; c < 2; ++c) pp[c] = partial_wrt_time(s, false, c, tk, idx);
                double dw[4] = {0, 0, 0, 0}, w[4] = {0, 0, 0, 0};
                int foff_ = 0, fcnt_ = 0;
                if (is_force_mutable(s, tk)) {
                    force_coef_partial(s, tk, idx, 0.0, dw);
                    fcnt_ = force_lin(s, tk, w, &foff_);
                }
                double xt[kNx];   // tangent state of the updated trajectory at node k
                for (int i = 0; i < 6; ++i) xt[i] = I.states[k][i];
                quat_log3(&I.states[k][6], xt + 6);
                for (int i = 0; i < 3; ++i) xt[9 + i] = I.states[k][10 + i];
                // out = dA_k vx + dB_k vu for the three vector pairs (z*, dz, updated trajectory); rows 3..5, 9..11
                double o_s[6], o_d[6], o_t[6];
                for (int i = 0; i < 6; ++i) o_s[i] = o_d[i] = o_t[i] = 0.0;
                for (int c = 0; c < 3; ++c) {
                    const double ec[3] = {c == 0 ? 1.0 : 0.0, c == 1 ? 1.0 : 0.0, c == 2 ? 1.0 : 0.0};
                    double ca[3];   // column c of dA[9:12, 0:3] = -(e_c x fp)
                    cross3g(ec, fp, ca);
                    for (int r = 0; r < 3; ++r) {
                        o_s[3 + r] += -ca[r] * zqp[k * kNx + c];
                        o_d[3 + r] += -ca[r] * dxs[k * kNx + c];
                        o_t[3 + r] += -ca[r] * xt[c];
                    }
                    if (fcnt_ > 0) {
                        double rc[3], pc[3];
                        cross3g(rel, ec, rc);
                        cross3g(pp, ec, pc);
                        const int c0 = fb + c * nv + foff_;
                        for (int a = 0; a < fcnt_; ++a) {
                            const double vs = us[c0 + a], vd = du[c0 + a], vt = ut[c0 + a];
                            o_s[c] += dw[a] * vs;
                            o_d[c] += dw[a] * vd;
                            o_t[c] += dw[a] * vt;
                            for (int r = 0; r < 3; ++r) {
                                const double cf = rc[r] * dw[a] + pc[r] * w[a];
                                o_s[3 + r] += cf * vs;
                                o_d[3 + r] += cf * vd;
                                o_t[3 + r] += cf * vt;
                            }
                        }
                    }
                    if (c != 2) {
                        double ef[3], efp[3];
                        cross3g(ec, f, ef);
                        cross3g(ec, fp, efp);
                        const int c0 = pb + c * npv + poff_;
                        for (int a = 0; a < pcnt_; ++a) {
                            const double vs = us[c0 + a], vd = du[c0 + a], vt = ut[c0 + a];
                            for (int r = 0; r < 3; ++r) {
                                const double cf = ef[r] * dwp[a] + efp[r] * wp[a];
                                o_s[3 + r] += cf * vs;
                                o_d[3 + r] += cf * vd;
                                o_t[3 + r] += cf * vt;
                            }
                        }
                    }
                }
                // dC = -(dA x~ + dB u~) + [0; fp; 0; rel x fp + pp x f]
                double c1[3], c2[3], dC[6];
                cross3g(rel, fp, c1);
                cross3g(pp, f, c2);
                for (int i = 0; i < 3; ++i) {
                    dC[i] = -o_t[i] + fp[i];
                    dC[3 + i] = -o_t[3 + i] + c1[i] + c2[i];
                }
                const double* dn = rec + (k + 1) * kNx;
                const double* nn = nus + (k + 1) * kNx;
                double v = 0;
                for (int i = 0; i < 3; ++i) {
                    v += dn[3 + i] * (o_s[i] + dC[i]) + nn[3 + i] * o_d[i];
                    v += dn[9 + i] * (o_s[3 + i] + dC[3 + i]) + nn[9 + i] * o_d[3 + i];
                }
                acc += dt * v;
            }
            if (k >= kEENodeStart) {   // foot-box rows of this foot at node k (mpc_single_rigid_body.cpp:705-733)
                for (int c = 0; c < 2; ++c) {
                    const int c0 = pb + c * npv + poff_;
                    double ps = 0, pd_ = 0;
                    for (int a = 0; a < pcnt_; ++a) {
                        ps += dwp[a] * us[c0 + a];
                        pd_ += dwp[a] * du[c0 + a];
                    }
                    const int row = m_force + 2 * (((k - kEENodeStart) * 4 + ee) * 2 + c);
                    acc += ps * (wv[row] - wv[row + 1]) + pd_ * (lam_g[row] - lam_g[row + 1]);
                }
            }
        }
        // --- force-box and friction-pyramid rows of the stance this contact time bounds (mpc.cpp:240-350, 416-531)
        {
            const int nct = s_nct[ee];
            const int kn = s_ctk[ee][idx];
            const bool is_td = s.ttype[kn] == kTouchDown && idx < nct - 1;
            const bool is_lo = s.ttype[kn] == kLiftOff && idx > 0;
            if ((is_td || is_lo) && lane < kSamplesPerStance) {
                const int i_lo = is_td ? idx : idx - 1;
                const double lower = s.t[s_ctk[ee][i_lo]], upper = s.t[s_ctk[ee][i_lo + 1]];
                int ord = 0;   // stances of this foot before this one
                for (int t = 0; t < i_lo; ++t)
                    if (s.ttype[s_ctk[ee][t]] == kTouchDown) ord++;
                const double frac = static_cast<double>(lane) / static_cast<double>(kSamplesPerStance);
                const double time = frac * (upper - lower) + lower;
                const double dtimedth = is_td ? -frac + 1.0 : frac;
                double dw[4], w[4];
                int off;
                force_coef_partial(s, time, idx, dtimedth, dw);
                const int cnt = force_lin(s, time, w, &off);
                const int srow = s_sb[ee] + ord * kSamplesPerStance + lane;
                if (srow < s_sb[ee + 1]) {
                    double Ps[3], Pd[3];
                    for (int c = 0; c < 3; ++c) {
                        const int c0 = fb + c * nv + off;
                        double a1 = 0, a2 = 0;
                        for (int a = 0; a < cnt; ++a) {
                            a1 += dw[a] * us[c0 + a];
                            a2 += dw[a] * du[c0 + a];
                        }
                        Ps[c] = a1;
                        Pd[c] = a2;
                    }
                    const double* y6 = wv + 6 * srow;
                    const double* l6 = lam_g + 6 * srow;
                    acc += (y6[0] - y6[1]) * Ps[2] + (l6[0] - l6[1]) * Pd[2];
                    const double gs[4] = {Ps[0] - mu_f * Ps[2], -Ps[0] - mu_f * Ps[2], Ps[1] - mu_f * Ps[2], -Ps[1] - mu_f * Ps[2]};
                    const double gd[4] = {Pd[0] - mu_f * Pd[2], -Pd[0] - mu_f * Pd[2], Pd[1] - mu_f * Pd[2], -Pd[1] - mu_f * Pd[2]};
                    for (int fc = 0; fc < 4; ++fc) acc += y6[2 + fc] * gs[fc] + l6[2 + fc] * gd[fc];
                }
            }
        }
        // --- touch-down rows (mpc_single_rigid_body.cpp:889-927) and foot-start rows (:733-752), one lane
        if (lane == 0) {
            if (next_touchdown_time(s, t0) - t0 < swing_time(s, t0) / 2) {
                int row = 0;
                for (int e2 = 0; e2 < ee; ++e2) row += 2 * Hd->td_flag[e2];
                const double td = next_touchdown_time(s, t0);
                double lin[2], wtmp[2];
                int off;
                const int cnt = pos_coef_partial(s, td, idx, lin);
                pos_lin(s, td, wtmp, &off);
                for (int c = 0; c < 2; ++c) {
                    const int c0 = pb + c * npv + off;
                    double ps = 0, pd_ = 0;
                    for (int a = 0; a < cnt; ++a) {
                        ps += lin[a] * us[c0 + a];
                        pd_ += lin[a] * du[c0 + a];
                    }
                    const double ppc = partial_wrt_time(s, false, c, td, idx);
                    if (row + c < ntd) acc += dnue[row + c] * ps + nue[row + c] * pd_ - dnue[row + c] * ppc;
                }
            }
            {   // the reference writes every foot's start-row partial into rows 0-1 of the block (kept)
                double lin[2], wtmp[2];
                int off;
                const int cnt = pos_coef_partial(s, t0, idx, lin);
                pos_lin(s, t0, wtmp, &off);
                for (int c = 0; c < 2; ++c) {
                    const int c0 = pb + c * npv + off;
                    double ps = 0, pd_ = 0;
                    for (int a = 0; a < cnt; ++a) {
                        ps += lin[a] * us[c0 + a];
                        pd_ += lin[a] * du[c0 + a];
                    }
                    acc += dnue[ntd + c] * ps + nue[ntd + c] * pd_;
                }
            }
        }
        acc = warp_sum(acc);
        if (lane == 0) gdH[ee * kMaxContacts + idx] = acc;
    }
}

size_t gradient_smem_bytes(const WsLayout& L, int nu_max, int ns_max) {
    GradCaps c;
    c.nu = (nu_max + 7) / 8 * 8;
    if (c.nu > L.max_nu) c.nu = L.max_nu;
    c.rows = 6 * (ns_max < 1 ? 1 : ns_max) + 2 * (L.N - 3) * 8;
    return grad_smem_for(L.N, c);
}

int launch_gradient(const Params& P, const Instance* inst, const WsLayout& L, char* ws, int B, int nu_max, int ns_max, int max_smem,
                    cudaStream_t stream) {
    GradCaps c;
    c.nu = (nu_max + 7) / 8 * 8;
    if (c.nu > L.max_nu) c.nu = L.max_nu;
    c.rows = 6 * (ns_max < 1 ? 1 : ns_max) + 2 * (L.N - 3) * 8;
    const size_t smem = grad_smem_for(L.N, c);
    if (smem > static_cast<size_t>(max_smem)) return -1;
    // per device and context, so set on every launch (see launch_ipm)
    cudaFuncSetAttribute(k_gradient, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    k_gradient<<<B, 256, smem, stream>>>(P, inst, L, ws, c.nu, c.rows);
    return 0;
}

}  // namespace bgg
