// bilevel-gait-gen_b200 -- kernel 7: the gait optimiser's outer step on the device.
//
//   k_gait_lp    GaitOptimizer::OptimizeContactTimes (mpc/gait_optimizer.cpp:185-364) with its constraint builders
//                CreatePolytopeConstraint / CreateStartConstraint / CreateTrustRegionConstraint /
//                CreateNextNodeConstraints (:410-534).  BFGS is disabled in the reference (Bk = 0, :276), so the
//                problem is the LP  min dH/dtheta . s  over the step s of all contact times.  Every row couples at most
//                two neighbouring contact times of ONE foot, so the LP separates into four chain LPs whose Newton
//                matrix is tridiagonal: one thread per (instance, foot) runs a Mehrotra predictor-corrector with a
//                Thomas solve per step.  (The reference runs OSQP at eps 1e-10 with polishing on the stacked 20 x 52
//                problem; both end on the LP's optimal vertex when it is unique.)
//   k_ls_expand  GaitOptimizer::LineSearch (:671-753): LS_SIZE copies of the MPC, copy i with the contact times
//                GetContactTimes(i / LS_SIZE) = ConvertQPVecToContactTimes(x_k + alpha_i step) (:645-669).  The copies
//                are then solved as ONE batch by the regular RTI kernels (the reference spawns 10 OpenMP threads).
//   k_ls_select  arg-min of cost / num_decision_vars over the copies that are not primal infeasible (:723-741) and
//                mpc.SetWarmStartTrajectory(best) (:743).
#include "bgg_kernels.cuh"

namespace bgg {

namespace {

constexpr int kLpMaxRows = 2 * kMaxContacts;
constexpr double kMinTime = 0.2;   // MIN_TIME, gait_optimizer.cpp:412

struct ChainLp {
    int n, nrows;
    int ra[kLpMaxRows], rb[kLpMaxRows];   // row = x[ra] - x[rb] (index -1: absent)
    double l[kLpMaxRows], u[kLpMaxRows];
    bool pinned[kMaxContacts];
    double g[kMaxContacts];
};

// rows of one foot, in the reference's order: chain rows 1..n-1, the last-node row, the trust-region rows
__device__ void build_chain_lp(ChainLp& lp, const double* t, const int* type, int n, const double* grad, double time, double trust) {
    lp.n = n;
    int next = -1;
    for (int j = 1; j < n; ++j)
        if (t[j] >= time) {
            next = j;
            break;
        }
    const bool next_td = next >= 0 && type[next] == kTouchDown;
    for (int i = 0; i < n; ++i) {
        lp.pinned[i] = (i == 0) || (next_td && (i == next || i == next - 1));   // CreateStartConstraint, CreateNextNodeConstraints
        lp.g[i] = grad[i];
    }
    int r = 0;
    for (int i = 1; i < n; ++i) {   // CreatePolytopeConstraint: each time between its neighbours, MIN_TIME apart
        lp.ra[r] = i - 1;
        lp.rb[r] = i;
        if (i != next || !next_td) {
            lp.u[r] = t[i] - t[i - 1] - kMinTime;
            lp.l[r] = -2;
        } else {
            lp.u[r] = t[next] - t[next - 1];
            lp.l[r] = -3;
        }
        ++r;
    }
    lp.ra[r] = n - 1;
    lp.rb[r] = -1;
    lp.l[r] = 0;
    lp.u[r] = 1;
    ++r;
    for (int i = 0; i < n; ++i) {   // CreateTrustRegionConstraint (infinity norm)
        lp.ra[r] = i;
        lp.rb[r] = -1;
        lp.l[r] = -trust;
        lp.u[r] = trust;
        ++r;
    }
    lp.nrows = r;
}

// min g.x  s.t.  l <= Bx <= u, x_i = 0 for pinned i.  Infeasible-start Mehrotra predictor-corrector on
// Bx + su = u, Bx - sl = l.  Returns 0 when converged, 2 at the iteration limit.
__device__ int solve_chain_lp(const ChainLp& lp, double* x, int* iters_out) {
    const int n = lp.n, R = lp.nrows;
    double su[kLpMaxRows], sl[kLpMaxRows], yu[kLpMaxRows], yl[kLpMaxRows];
    double dsu[kLpMaxRows], dsl[kLpMaxRows], dyu[kLpMaxRows], dyl[kLpMaxRows];
    double rpu[kLpMaxRows], rpl[kLpMaxRows], rcu[kLpMaxRows], rcl[kLpMaxRows];
    bool live[kLpMaxRows];
    double dx[kMaxContacts], rd[kMaxContacts];
    for (int i = 0; i < n; ++i) x[i] = 0.0;
    int nlive = 0;
    for (int r = 0; r < R; ++r) {
        const bool fa = lp.ra[r] >= 0 && !lp.pinned[lp.ra[r]], fb = lp.rb[r] >= 0 && !lp.pinned[lp.rb[r]];
        live[r] = fa || fb;   // a row between pinned times is a constant
        nlive += live[r];
        su[r] = fmax(lp.u[r], 0.1);
        sl[r] = fmax(-lp.l[r], 0.1);
        yu[r] = yl[r] = 1.0;
    }
    auto Bx = [&](const double* v, int r) {
        double s = 0;
        if (lp.ra[r] >= 0) s += v[lp.ra[r]];
        if (lp.rb[r] >= 0) s -= v[lp.rb[r]];
        return s;
    };
    // one Newton solve for complementarity targets (rcu, rcl)
    auto newton = [&]() {
        double dg[kMaxContacts], lo[kMaxContacts], rhs[kMaxContacts];   // tridiagonal: dg diagonal, lo[i] = M[i][i-1]
        for (int i = 0; i < n; ++i) {
            dg[i] = 0;
            lo[i] = 0;
            rhs[i] = -rd[i];
        }
        for (int r = 0; r < R; ++r) {
            if (!live[r]) continue;
            const double D = yu[r] / su[r] + yl[r] / sl[r];
            const double c = (rcu[r] + yu[r] * rpu[r]) / su[r] - (rcl[r] - yl[r] * rpl[r]) / sl[r];
            const int a = lp.ra[r], b = lp.rb[r];
            if (a >= 0) {
                dg[a] += D;
                rhs[a] -= c;
            }
            if (b >= 0) {
                dg[b] += D;
                rhs[b] += c;
            }
            if (a >= 0 && b >= 0) lo[b] -= D;   // b = a + 1
        }
        for (int i = 0; i < n; ++i)
            if (lp.pinned[i]) {
                dg[i] = 1.0;
                rhs[i] = 0.0;
                lo[i] = 0.0;
                if (i + 1 < n) lo[i + 1] = 0.0;
            }
        // Thomas algorithm (M symmetric tridiagonal, positive definite on the free times)
        double cp[kMaxContacts];
        for (int i = 0; i < n; ++i) {
            const double up = (i + 1 < n) ? lo[i + 1] : 0.0;
            double den = dg[i];
            if (i > 0) {
                den -= lo[i] * cp[i - 1];
                rhs[i] -= lo[i] * rhs[i - 1];
            }
            cp[i] = up / den;
            rhs[i] /= den;
        }
        for (int i = n - 1; i >= 0; --i) dx[i] = rhs[i] - ((i + 1 < n) ? cp[i] * dx[i + 1] : 0.0);
        for (int r = 0; r < R; ++r) {
            if (!live[r]) {
                dsu[r] = dsl[r] = dyu[r] = dyl[r] = 0;
                continue;
            }
            const double bdx = Bx(dx, r);
            dsu[r] = -rpu[r] - bdx;
            dsl[r] = rpl[r] + bdx;
            dyu[r] = (rcu[r] - yu[r] * dsu[r]) / su[r];
            dyl[r] = (rcl[r] - yl[r] * dsl[r]) / sl[r];
        }
    };
    auto max_step = [&]() {
        double a = 1e300;
        for (int r = 0; r < R; ++r) {
            if (!live[r]) continue;
            if (dsu[r] < 0) a = fmin(a, -su[r] / dsu[r]);
            if (dsl[r] < 0) a = fmin(a, -sl[r] / dsl[r]);
            if (dyu[r] < 0) a = fmin(a, -yu[r] / dyu[r]);
            if (dyl[r] < 0) a = fmin(a, -yl[r] / dyl[r]);
        }
        return a;
    };
    int it = 0, status = 2;
    double gs = 1.0;
    for (int i = 0; i < n; ++i) gs = fmax(gs, fabs(lp.g[i]));
    for (it = 0; it < 80; ++it) {
        double mu = 0, nrp = 0, nrd = 0;
        for (int i = 0; i < n; ++i) rd[i] = lp.pinned[i] ? 0.0 : lp.g[i];
        for (int r = 0; r < R; ++r) {
            if (!live[r]) continue;
            const double bx = Bx(x, r);
            rpu[r] = bx + su[r] - lp.u[r];
            rpl[r] = bx - sl[r] - lp.l[r];
            nrp = fmax(nrp, fmax(fabs(rpu[r]), fabs(rpl[r])));
            mu += su[r] * yu[r] + sl[r] * yl[r];
            const double y = yu[r] - yl[r];
            if (lp.ra[r] >= 0 && !lp.pinned[lp.ra[r]]) rd[lp.ra[r]] += y;
            if (lp.rb[r] >= 0 && !lp.pinned[lp.rb[r]]) rd[lp.rb[r]] -= y;
        }
        mu /= fmax(1.0, 2.0 * nlive);
        for (int i = 0; i < n; ++i) nrd = fmax(nrd, fabs(rd[i]));
        // the complementarity target is absolute: a time whose gradient entry is 1e-5 sits mu / 1e-5 off its vertex
        if (nrp < 1e-11 && nrd < 1e-11 * gs && mu < 1e-16) {
            status = 0;
            break;
        }
        for (int r = 0; r < R; ++r) {
            rcu[r] = -su[r] * yu[r];
            rcl[r] = -sl[r] * yl[r];
        }
        newton();
        const double aa = fmin(1.0, max_step());
        double mu_aff = 0;
        for (int r = 0; r < R; ++r)
            if (live[r]) mu_aff += (su[r] + aa * dsu[r]) * (yu[r] + aa * dyu[r]) + (sl[r] + aa * dsl[r]) * (yl[r] + aa * dyl[r]);
        mu_aff /= fmax(1.0, 2.0 * nlive);
        const double sr = mu_aff / mu, sigma = sr * sr * sr;
        for (int r = 0; r < R; ++r) {
            rcu[r] = -su[r] * yu[r] - dsu[r] * dyu[r] + sigma * mu;
            rcl[r] = -sl[r] * yl[r] - dsl[r] * dyl[r] + sigma * mu;
        }
        newton();
        const double al = fmin(1.0, 0.995 * max_step());
        for (int i = 0; i < n; ++i)
            if (!lp.pinned[i]) x[i] += al * dx[i];
        for (int r = 0; r < R; ++r) {
            if (!live[r]) continue;
            su[r] += al * dsu[r];
            sl[r] += al * dsl[r];
            yu[r] += al * dyu[r];
            yl[r] += al * dyl[r];
        }
    }
    *iters_out = it;
    return status;
}

// ConvertQPVecToContactTimes for one foot, gait_optimizer.cpp:651-669
__device__ void convert_times(const double* vec, int n, double* out) {
    for (int i = 0; i < n; ++i) {
        out[i] = vec[i];
        if (i > 0) {
            const double d = out[i - 1] - out[i];
            if (d <= 1e-3 && d > 0) out[i] = out[i - 1];
        }
    }
}

__device__ int foot_contacts(const FootSpline& s, double* t, int* type) {
    int c = 0;
    for (int i = 0; i < s.n && c < kMaxContacts; ++i)
        if (s.ttype[i] != kInter) {
            t[c] = s.t[i];
            type[c] = s.ttype[i];
            c++;
        }
    return c;
}

}  // namespace

// grad: [B][4][kMaxContacts] or nullptr (use the instance's last gradient in the workspace)
__global__ void k_gait_lp(const Instance* __restrict__ inst, WsLayout L, const char* __restrict__ ws_base, int B, const double* __restrict__ grad,
                          const double* __restrict__ time, double trust, double alpha, double* __restrict__ step_out,
                          double* __restrict__ xk_out, double* __restrict__ times_out, int32_t* __restrict__ status_out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * kNumEE) return;
    const int b = idx / kNumEE, e = idx % kNumEE;
    const FootSpline& s = inst[b].foot[e];
    double t[kMaxContacts], x[kMaxContacts], xn[kMaxContacts], tn[kMaxContacts];
    int type[kMaxContacts];
    const int n = foot_contacts(s, t, type);
    const double* g = grad ? grad + static_cast<size_t>(idx) * kMaxContacts
                           : reinterpret_cast<const double*>(ws_base + static_cast<size_t>(b) * L.stride + L.gdH) + e * kMaxContacts;
    ChainLp lp;
    build_chain_lp(lp, t, type, n, g, time[b], trust);
    int iters = 0;
    const int st = solve_chain_lp(lp, x, &iters);
    for (int i = 0; i < n; ++i) {
        x[i] = alpha * x[i];     // step_ = alpha * step_ (:339)
        xn[i] = t[i] + x[i];     // xkp1_ = xk_ + step_
    }
    convert_times(xn, n, tn);
    double* so = step_out + static_cast<size_t>(idx) * kMaxContacts;
    double* xo = xk_out + static_cast<size_t>(idx) * kMaxContacts;
    double* to = times_out + static_cast<size_t>(idx) * kMaxContacts;
    for (int i = 0; i < kMaxContacts; ++i) {
        so[i] = (i < n) ? x[i] : 0.0;
        xo[i] = (i < n) ? t[i] : 0.0;
        to[i] = (i < n) ? tn[i] : 0.0;
    }
    status_out[idx] = st;
}

// child (b, k) = copy of parent b with the contact times of alpha_k = k / K; inputs replicated for the batch solve
__global__ void __launch_bounds__(256) k_ls_expand(const Instance* __restrict__ parent, Instance* __restrict__ child, int K,
                                                   const double* __restrict__ xk, const double* __restrict__ step,
                                                   const double* __restrict__ state, const double* __restrict__ t0,
                                                   const double* __restrict__ ee, double* __restrict__ c_state,
                                                   double* __restrict__ c_t0, double* __restrict__ c_ee) {
    const int c = blockIdx.x, b = c / K, k = c % K, tid = threadIdx.x;
    const double* src = reinterpret_cast<const double*>(parent + b);
    double* dst = reinterpret_cast<double*>(child + c);
    for (int i = tid; i < static_cast<int>(sizeof(Instance) / 8); i += blockDim.x) dst[i] = src[i];
    if (tid < kNxMan) c_state[c * kNxMan + tid] = state[b * kNxMan + tid];
    if (tid < 12) c_ee[c * 12 + tid] = ee[b * 12 + tid];
    if (tid == 0) c_t0[c] = t0[b];
    __syncthreads();
    if (tid < kNumEE) {
        FootSpline& s = child[c].foot[tid];
        const double a = static_cast<double>(k) / static_cast<double>(K);
        double vec[kMaxContacts], tn[kMaxContacts];
        const int n = num_contacts(s) < kMaxContacts ? num_contacts(s) : kMaxContacts;
        for (int i = 0; i < n; ++i)
            vec[i] = xk[(b * kNumEE + tid) * kMaxContacts + i] + a * step[(b * kNumEE + tid) * kMaxContacts + i];
        convert_times(vec, n, tn);
        set_contact_times(s, tn, n);   // MPC::UpdateContactTimes -> EndEffectorSplines::SetContactTimes
    }
}

__global__ void __launch_bounds__(256) k_ls_select(Instance* __restrict__ parent, const Instance* __restrict__ child, WsLayout L,
                                                   const char* __restrict__ child_ws, int K, int32_t* __restrict__ best_out,
                                                   double* __restrict__ costs, int32_t* __restrict__ quality) {
    const int b = blockIdx.x, tid = threadIdx.x;
    __shared__ int s_best;
    if (tid == 0) {
        int best = -1;
        double cmin = 1e10;
        for (int k = 0; k < K; ++k) {
            const WsHeader* h = reinterpret_cast<const WsHeader*>(child_ws + static_cast<size_t>(b * K + k) * L.stride + L.hdr);
            const int st = h->error ? static_cast<int>(kOther) : h->status;
            const double c = h->error ? 1e300 : h->cost / h->n;   // GetCost() / GetNumDecisionVars(), :716
            costs[b * K + k] = c;
            quality[b * K + k] = st;
            // a child that k_prepare refused was never solved (the reference's copy would have thrown): it cannot win
            if (!h->error && c == c && c < cmin && st != kPrimalInfeasible) {
                cmin = c;
                best = k;
            }
        }
        best_out[b] = best;          // -1: "no valid trajectories... using the current one" -> copy 0 (:737-741)
        s_best = best < 0 ? 0 : best;
    }
    __syncthreads();
    // MPC::SetWarmStartTrajectory (mpc.cpp:110-119) assigns prev_traj_ and init_time_ only: the parent's adaptive foot box
    // and run count are its own (the child's solve changed the child's copies through Increase / DecreaseEEBox)
    const Instance* src = child + b * K + s_best;
    Instance* dst = parent + b;
    const double* ssrc = reinterpret_cast<const double*>(src->states);
    double* sdst = reinterpret_cast<double*>(dst->states);
    for (int i = tid; i < static_cast<int>(sizeof(src->states) / 8); i += blockDim.x) sdst[i] = ssrc[i];
    const double* fsrc = reinterpret_cast<const double*>(src->foot);
    double* fdst = reinterpret_cast<double*>(dst->foot);
    for (int i = tid; i < static_cast<int>(sizeof(src->foot) / 8); i += blockDim.x) fdst[i] = fsrc[i];
    if (tid == 0) dst->init_time = src->init_time;
}

void launch_gait_lp(const Instance* inst, const WsLayout& L, const char* ws, int B, const double* grad, const double* time, double trust,
                    double alpha, double* step, double* xk, double* times, int32_t* status, cudaStream_t stream) {
    const int tot = B * kNumEE;
    k_gait_lp<<<(tot + 63) / 64, 64, 0, stream>>>(inst, L, ws, B, grad, time, trust, alpha, step, xk, times, status);
}
void launch_ls_expand(const Instance* parent, Instance* child, int B, int K, const double* xk, const double* step, const double* state,
                      const double* t0, const double* ee, double* c_state, double* c_t0, double* c_ee, cudaStream_t stream) {
    k_ls_expand<<<B * K, 256, 0, stream>>>(parent, child, K, xk, step, state, t0, ee, c_state, c_t0, c_ee);
}
void launch_ls_select(Instance* parent, const Instance* child, const WsLayout& L, const char* child_ws, int B, int K, int32_t* best,
                      double* costs, int32_t* quality, cudaStream_t stream) {
    k_ls_select<<<B, 256, 0, stream>>>(parent, child, L, child_ws, K, best, costs, quality);
}

}  // namespace bgg
