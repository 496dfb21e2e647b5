// Export of one contact time's parameter partials of the QP -- what MPCSingleRigidBody::ComputeParamPartialsClarabel
// (mpc/mpc_single_rigid_body.cpp:642-792) writes into QPPartials::dA / dG / db: the model partials of every node
// (SingleRigidBodyModel::ComputeLinearizationPartialWrtContactTimes, single_rigid_body_model.cpp:458-555), the force-box and
// friction-pyramid row partials (mpc.cpp:416-531, 240-350), the foot-box, foot-start and touch-down row partials (:705-752,
// :889-927).  On the gait-optimisation path these are generated and contracted inside k_gradient without ever being stored
// (csrc/bgg_gradient.cu); this kernel writes them out, on request, for callers that want the matrices themselves
// (test/mpc_test.cpp:140-236 compares them with finite differences of the assembled QP).  One CTA, one (instance, foot, contact
// time); the host side of the C ABI turns the blocks into triplets in the reference's row / column numbering.
#include <cuda_runtime.h>

#include "bgg_kernels.cuh"
#include "bgg_spline.cuh"
#include "bgg_ws.cuh"

namespace bgg {
namespace {

__device__ __forceinline__ void cross3p(const double a[3], const double b[3], double o[3]) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}

}  // namespace

// Layout of `out` (doubles): header [kPartHdr] | per node k < N: dA[9] (rows 9..11 x cols 0..2), dB[6][nu_cap] (rows 3..5, 9..11),
// dC[6] | force samples [10][3][6] (off, cnt, p[4]) | foot-box rows [(N+1)][2][4] (off, cnt, p[2]) | foot-start [2][4] | touch-down
// [2][5] (off, cnt, p[2], db).  Nothing is scaled by dt here.
__host__ __device__ size_t param_partials_doubles(int N, int nu_cap) {
    return kPartHdr + static_cast<size_t>(N) * (9 + 6 * nu_cap + 6) + 10 * 3 * 6 + static_cast<size_t>(N + 1) * 2 * 4 + 2 * 4 + 2 * 5;
}

__global__ void __launch_bounds__(128) k_param_partials(Params P, const Instance* __restrict__ inst, WsLayout L, const char* __restrict__ ws_base, int b,
                                                        int ee, int idx, int nu_cap, double* __restrict__ out, double* __restrict__ ut) {
    const int tid = threadIdx.x, nth = blockDim.x;
    const Instance& I = inst[b];
    const WsHeader* Hd = reinterpret_cast<const WsHeader*>(ws_base + static_cast<size_t>(b) * L.stride + L.hdr);
    const FootSpline& s = I.foot[ee];
    const int N = P.N, nf = Hd->nf, nu = Hd->nu;
    const double dt = P.dt, t0 = I.init_time;
    __shared__ int s_ctk[kMaxContacts], s_nct;
    double* dyn = out + kPartHdr;
    const size_t dyn_stride = 9 + 6 * static_cast<size_t>(nu_cap) + 6;
    double* frc = dyn + N * dyn_stride;
    double* loc = frc + 10 * 3 * 6;
    double* start = loc + static_cast<size_t>(N + 1) * 2 * 4;
    double* tdp = start + 2 * 4;
    const size_t total = param_partials_doubles(N, nu_cap);
    for (size_t i = tid; i < total; i += nth) out[i] = 0.0;
    if (tid == 0) {
        int c = 0;
        for (int i = 0; i < s.n && c < kMaxContacts; ++i)
            if (s.ttype[i] != kInter) s_ctk[c++] = i;
        s_nct = c;
    }
    for (int i = tid; i < kNumEE * 5; i += nth) {   // SplinesAsVec of the updated trajectory
        const int e = i / 5, c = i % 5;
        if (c < 3) get_force_vars(I.foot[e], c, ut + Hd->fbase[e] + c * Hd->nfv[e]);
        else get_pos_vars(I.foot[e], c - 3, ut + nf + Hd->pbase[e] + (c - 3) * Hd->npv[e]);
    }
    __syncthreads();
    const int fb = Hd->fbase[ee], nv = Hd->nfv[ee], pb = nf + Hd->pbase[ee], npv = Hd->npv[ee];
    if (tid == 0) {
        out[0] = nu;
        out[1] = nf;
        for (int e = 0; e < kNumEE; ++e) {
            out[2 + e] = Hd->fbase[e];
            out[6 + e] = Hd->nfv[e];
            out[10 + e] = Hd->pbase[e];
            out[14 + e] = Hd->npv[e];
        }
        // rows of the force-box / friction blocks (mpc.cpp:424-441, 248-265): stances of the feet before this one, stances of this
        // foot before this contact time, one back for a lift-off
        int row = 0;
        for (int e = 0; e < ee; ++e) {
            const FootSpline& o = I.foot[e];
            int last = -1, cnt = 0;
            for (int i = 0; i < o.n; ++i)
                if (o.ttype[i] != kInter) last = i;
            for (int i = 0; i < o.n; ++i)
                if (o.ttype[i] == kTouchDown && i != last) cnt++;
            row += cnt;
        }
        for (int t = 0; t < idx; ++t)
            if (s.ttype[s_ctk[t]] == kTouchDown) row++;
        const int kn = s_ctk[idx];
        const bool is_td = s.ttype[kn] == kTouchDown && idx < s_nct - 1;
        const bool is_lo = s.ttype[kn] == kLiftOff && idx > 0;
        if (is_lo) row--;
        out[18] = (is_td || is_lo) ? 1.0 : 0.0;
        out[19] = row * kSamplesPerStance;
        // touch-down rows (mpc_single_rigid_body.cpp:889-927): the row offset counts the feet before with the constraint's own
        // td_fraction test, the partial itself is written when this foot passes the swing_time / 2 test
        int trow = 0;
        for (int e = 0; e < ee; ++e) trow += 2 * Hd->td_flag[e];
        out[20] = (next_touchdown_time(s, t0) - t0 < swing_time(s, t0) / 2) ? 1.0 : 0.0;
        out[21] = trow;
        out[22] = Hd->n_samples;
        out[23] = Hd->n_td;
    }
    // ---- model partials per node (single_rigid_body_model.cpp:458-555)
    for (int k = tid; k < N; k += nth) {
        double* dA = dyn + k * dyn_stride;
        double* dB = dA + 9;
        double* dC = dB + 6 * static_cast<size_t>(nu_cap);
        const double tk = k * dt + t0;
        double dwp[2], wp[2];
        int poff;
        const int pcnt = pos_coef_partial(s, tk, idx, dwp);
        pos_lin(s, tk, wp, &poff);
        if (pcnt < 2) wp[1] = 0.0;
        double fp[3], pp[3] = {0, 0, 0}, f[3], rel[3];
        for (int c = 0; c < 3; ++c) {
            fp[c] = partial_wrt_time(s, true, c, tk, idx);
            f[c] = value_at(s, true, c, tk);
            rel[c] = value_at(s, false, c, tk) - I.states[k][c];
        }
        for (int c = 0; c < 2; ++c) pp[c] = partial_wrt_time(s, false, c, tk, idx);
        double dw[4] = {0, 0, 0, 0}, w[4] = {0, 0, 0, 0};
        int foff = 0, fcnt = 0;
        if (is_force_mutable(s, tk)) {
            force_coef_partial(s, tk, idx, 0.0, dw);
            fcnt = force_lin(s, tk, w, &foff);
        }
        double xt[kNx];   // tangent state of the updated trajectory at node k
        for (int i = 0; i < 6; ++i) xt[i] = I.states[k][i];
        quat_log3(&I.states[k][6], xt + 6);
        for (int i = 0; i < 3; ++i) xt[9 + i] = I.states[k][10 + i];
        double o_t[6] = {0, 0, 0, 0, 0, 0};
        for (int c = 0; c < 3; ++c) {
            const double ec[3] = {c == 0 ? 1.0 : 0.0, c == 1 ? 1.0 : 0.0, c == 2 ? 1.0 : 0.0};
            double ca[3];   // column c of dA[9:12, 0:3] = -(e_c x fp)
            cross3p(ec, fp, ca);
            for (int r = 0; r < 3; ++r) {
                dA[3 * r + c] = -ca[r];
                o_t[3 + r] += -ca[r] * xt[c];
            }
            if (fcnt > 0) {
                double rc[3], pc[3];
                cross3p(rel, ec, rc);
                cross3p(pp, ec, pc);
                const int c0 = fb + c * nv + foff;
                for (int a = 0; a < fcnt; ++a) {
                    dB[static_cast<size_t>(c) * nu_cap + c0 + a] += dw[a];
                    o_t[c] += dw[a] * ut[c0 + a];
                    for (int r = 0; r < 3; ++r) {
                        const double cf = rc[r] * dw[a] + pc[r] * w[a];
                        dB[static_cast<size_t>(3 + r) * nu_cap + c0 + a] += cf;
                        o_t[3 + r] += cf * ut[c0 + a];
                    }
                }
            }
            if (c != 2) {
                double ef[3], efp[3];
                cross3p(ec, f, ef);
                cross3p(ec, fp, efp);
                const int c0 = pb + c * npv + poff;
                for (int a = 0; a < pcnt; ++a)
                    for (int r = 0; r < 3; ++r) {
                        const double cf = ef[r] * dwp[a] + efp[r] * wp[a];
                        dB[static_cast<size_t>(3 + r) * nu_cap + c0 + a] += cf;
                        o_t[3 + r] += cf * ut[c0 + a];
                    }
            }
        }
        double c1[3], c2[3];   // dC = -(dA x~ + dB u~) + [0; fp; 0; rel x fp + pp x f]
        cross3p(rel, fp, c1);
        cross3p(pp, f, c2);
        for (int i = 0; i < 3; ++i) {
            dC[i] = -o_t[i] + fp[i];
            dC[3 + i] = -o_t[3 + i] + c1[i] + c2[i];
        }
    }
    // ---- foot-box rows of this foot at nodes 4 .. N (mpc_single_rigid_body.cpp:705-733)
    for (int i = tid; i < 2 * (N + 1); i += nth) {
        const int k = i >> 1, c = i & 1;
        if (k < kEENodeStart) continue;
        const double tk = k * dt + t0;
        double dwp[2], wp[2];
        int poff;
        const int pcnt = pos_coef_partial(s, tk, idx, dwp);
        pos_lin(s, tk, wp, &poff);
        double* o = loc + (static_cast<size_t>(k) * 2 + c) * 4;
        o[0] = Hd->pbase[ee] + c * npv + poff;
        o[1] = pcnt;
        o[2] = dwp[0];
        o[3] = pcnt > 1 ? dwp[1] : 0.0;
    }
    // ---- foot-start rows (:733-752) and touch-down rows (:889-927)
    if (tid < 2) {
        const int c = tid;
        double lin[2], wtmp[2];
        int off;
        int cnt = pos_coef_partial(s, t0, idx, lin);
        pos_lin(s, t0, wtmp, &off);
        start[4 * c] = Hd->pbase[ee] + c * npv + off;
        start[4 * c + 1] = cnt;
        start[4 * c + 2] = lin[0];
        start[4 * c + 3] = cnt > 1 ? lin[1] : 0.0;
        const double td = next_touchdown_time(s, t0);
        cnt = pos_coef_partial(s, td, idx, lin);
        pos_lin(s, td, wtmp, &off);
        tdp[5 * c] = Hd->pbase[ee] + c * npv + off;
        tdp[5 * c + 1] = cnt;
        tdp[5 * c + 2] = lin[0];
        tdp[5 * c + 3] = cnt > 1 ? lin[1] : 0.0;
        tdp[5 * c + 4] = partial_wrt_time(s, false, c, td, idx);
    }
    // ---- force-box and friction-pyramid samples of the stance this contact time bounds (mpc.cpp:416-531, 240-350)
    if (tid < kSamplesPerStance * 3) {
        const int i = tid / 3, c = tid % 3;
        const int kn = s_ctk[idx];
        const bool is_td = s.ttype[kn] == kTouchDown && idx < s_nct - 1;
        const bool is_lo = s.ttype[kn] == kLiftOff && idx > 0;
        if (is_td || is_lo) {
            const int i_lo = is_td ? idx : idx - 1;
            const double lower = s.t[s_ctk[i_lo]], upper = s.t[s_ctk[i_lo + 1]];
            const double frac = static_cast<double>(i) / static_cast<double>(kSamplesPerStance);
            const double time = frac * (upper - lower) + lower;
            const double dtimedth = is_td ? -frac + 1.0 : frac;
            double dw[4], w[4];
            int off;
            force_coef_partial(s, time, idx, dtimedth, dw);
            const int cnt = force_lin(s, time, w, &off);
            double* o = frc + (static_cast<size_t>(i) * 3 + c) * 6;
            o[0] = fb + c * nv + off;
            o[1] = cnt;
            for (int a = 0; a < 4; ++a) o[2 + a] = a < cnt ? dw[a] : 0.0;
        }
    }
}

void launch_param_partials(const Params& P, const Instance* inst, const WsLayout& L, const char* ws, int b, int ee, int idx, int nu_cap, double* out,
                           double* ut, cudaStream_t stream) {
    k_param_partials<<<1, 128, 0, stream>>>(P, inst, L, ws, b, ee, idx, nu_cap, out, ut);
}

}  // namespace bgg
