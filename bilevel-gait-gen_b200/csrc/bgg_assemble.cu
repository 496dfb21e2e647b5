// bilevel-gait-gen_b200 -- kernel 3c: the reference's sparse QP, bit-identical in sparsity, from the structured form.
//
// The reference pushes triplets into utils::SparseMatrixBuilder (entries that are exactly 0.0 are skipped,
// utils/sparse_matrix_builder.cpp:22-30) and compresses them with Eigen's setFromTriplets (qp_data.cpp:169-178):
// column-major, rows ascending inside a column.  The solver kernels never need that matrix -- they work on the
// structured rows of csrc/bgg_ws.cuh -- but callers of MPC::GetQPData() (test/mpc_test.cpp:125,140-171) and the
// parity tests do.  One CTA per instance walks every column twice (count, exclusive scan, fill); a column's rows come
// out ascending by construction because the constraint blocks are stacked in the order of
// SingleRigidBodyModel's constraint list (single_rigid_body_model.cpp:22-29):
//   Dynamics | ForceBox (+ rows, then - rows) | FrictionCone | EndEffectorLocation (+, then -) | TDPosition | EndEffectorStart
// Right-hand side: QPData::ConstructVectors, Clarabel form A z + s = ub (qp_data.cpp:200-289).
// Compiled with -fmad=false (exact zeros must stay exact).
#include "bgg_kernels.cuh"

namespace bgg {

namespace {

struct ColCtx {
    const Params* P;
    const WsHeader* Hd;
    const NodeLin* nodes;
    const Sample* samples;
    const EqRow* eqs;
    const int* sb;   // per-foot sample ranges [kNumEE + 1]
    int N, nf, ns, ne;
    int r_fb, r_cone, r_ee, r_td;
};

__device__ __forceinline__ void cross3a(const double a[3], const double b[3], double o[3]) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}

// Visits the structural entries of column j in ascending row order; emit(row, value) drops exact zeros itself.
template <class Emit>
__device__ void visit_column(const ColCtx& c, int j, Emit&& emit) {
    const int N = c.N, nstate = kNx * (N + 1);
    const double dt = c.P->dt;
    if (j < nstate) {
        // state column x_k[i]: -1 of row block k, Ad_k[:, i] in row block k + 1 (mpc_single_rigid_body.cpp:222,250-256),
        // foot-box rows -p_c(node) / +p_c(node) for i < 2, k >= 4 (:408-431)
        const int k = j / kNx, i = j % kNx;
        emit(kNx * k + i, -1.0);
        if (k < N) {
            const NodeLin& nl = c.nodes[k];
            for (int r = 0; r < kNx; ++r) emit(kNx * (k + 1) + r, nl.Ad[r * kNx + i]);
        }
        if (i < 2 && k >= kEENodeStart) {
            for (int foot = 0; foot < kNumEE; ++foot) emit(c.r_ee + ((k - kEENodeStart) * kNumEE + foot) * 2 + i, -1.0);
            for (int foot = 0; foot < kNumEE; ++foot)
                emit(c.r_ee + c.ne + ((k - kEENodeStart) * kNumEE + foot) * 2 + i, 1.0);
        }
        return;
    }
    const int ju = j - nstate;
    if (ju < c.nf) {
        int e = 0;
        while (e < kNumEE - 1 && ju >= c.Hd->fbase[e + 1]) ++e;
        const int nv = c.Hd->nfv[e], loc = ju - c.Hd->fbase[e], cc = loc / nv, i = loc % nv;
        const double ec[3] = {cc == 0 ? 1.0 : 0.0, cc == 1 ? 1.0 : 0.0, cc == 2 ? 1.0 : 0.0};
        for (int k = 0; k < N; ++k) {   // Bd_k (single_rigid_body_model.cpp:113-135)
            const NodeLin& nl = c.nodes[k];
            const int a = i - nl.foff[e];
            if (a < 0 || a >= nl.fcnt[e]) continue;
            double rc[3];
            cross3a(nl.rel[e], ec, rc);
            emit(kNx * (k + 1) + 3 + cc, dt * nl.fw[e][a]);
            for (int r = 0; r < 3; ++r) emit(kNx * (k + 1) + 9 + r, dt * (rc[r] * nl.fw[e][a]));
        }
        if (cc == 2)   // force box: + rows of every sample, then the - rows (mpc.cpp:352-414)
            for (int pass = 0; pass < 2; ++pass)
                for (int s = c.sb[e]; s < c.sb[e + 1]; ++s) {
                    const Sample& sp = c.samples[s];
                    const int a = i - sp.off;
                    if (a < 0 || a >= sp.cnt) continue;
                    emit(c.r_fb + pass * c.ns + s, pass == 0 ? 1.0 * sp.w[a] : -1.0 * sp.w[a]);
                }
        // friction pyramid rows (1,0,-mu) (-1,0,-mu) (0,1,-mu) (0,-1,-mu) per sample (mpc.cpp:153-209)
        const double mu = c.P->friction_coef;
        for (int s = c.sb[e]; s < c.sb[e + 1]; ++s) {
            const Sample& sp = c.samples[s];
            const int a = i - sp.off;
            if (a < 0 || a >= sp.cnt) continue;
            for (int fc = 0; fc < 4; ++fc) {
                double coef;
                if (cc == 2) coef = -mu;
                else if (cc == 0) coef = (fc == 0) ? 1.0 : (fc == 1 ? -1.0 : 0.0);
                else coef = (fc == 2) ? 1.0 : (fc == 3 ? -1.0 : 0.0);
                emit(c.r_cone + 4 * s + fc, coef * sp.w[a]);
            }
        }
        return;
    }
    // position column
    const int jp = ju - c.nf;
    int e = 0;
    while (e < kNumEE - 1 && jp >= c.Hd->pbase[e + 1]) ++e;
    const int nv = c.Hd->npv[e], loc = jp - c.Hd->pbase[e], cc = loc / nv, v = loc % nv;
    const double ec[3] = {cc == 0 ? 1.0 : 0.0, cc == 1 ? 1.0 : 0.0, 0.0};
    for (int k = 0; k < N; ++k) {   // Bd_k position part (single_rigid_body_model.cpp:137-148)
        const NodeLin& nl = c.nodes[k];
        const int a = v - nl.poff[e];
        if (a < 0 || a >= nl.pcnt[e]) continue;
        double ef[3];
        cross3a(ec, nl.f[e], ef);
        for (int r = 0; r < 3; ++r) emit(kNx * (k + 1) + 9 + r, dt * (ef[r] * nl.pw[e][a]));
    }
    for (int pass = 0; pass < 2; ++pass)   // foot box (mpc_single_rigid_body.cpp:381-443)
        for (int k = kEENodeStart; k <= N; ++k) {
            const NodeLin& nl = c.nodes[k];
            const int a = v - nl.poff[e];
            if (a < 0 || a >= nl.pcnt[e]) continue;
            emit(c.r_ee + pass * c.ne + ((k - kEENodeStart) * kNumEE + e) * 2 + cc, pass == 0 ? nl.pw[e][a] : -nl.pw[e][a]);
        }
    for (int r = 0; r < c.Hd->n_eq; ++r) {   // touch-down rows, then foot-start rows (:849-887, 445-475)
        const EqRow& q = c.eqs[r];
        if (q.pad != e * 2 + cc) continue;
        const int a = ju - q.col[0];
        if (a < 0 || a >= q.cnt) continue;
        emit(c.r_td + r, q.w[a]);
    }
}

}  // namespace

// dims[b] = {n, m, nnz, num_eq_rows_total, num_ineq_rows, error}
__global__ void __launch_bounds__(256) k_export_csc(Params P, WsLayout L, const char* __restrict__ ws_base, int32_t* __restrict__ colptr,
                                                    int32_t* __restrict__ rowidx, double* __restrict__ val,
                                                    double* __restrict__ p_diag, double* __restrict__ q_out,
                                                    double* __restrict__ ub_out, int32_t* __restrict__ dims, int n_stride,
                                                    int m_stride, int nnz_cap) {
    const int b = blockIdx.x, tid = threadIdx.x, nth = blockDim.x;
    const char* ws = ws_base + static_cast<size_t>(b) * L.stride;
    const WsHeader* Hd = reinterpret_cast<const WsHeader*>(ws + L.hdr);
    int32_t* cp = colptr + static_cast<size_t>(b) * (n_stride + 1);
    int32_t* ri = rowidx + static_cast<size_t>(b) * nnz_cap;
    double* va = val + static_cast<size_t>(b) * nnz_cap;
    int32_t* dm = dims + static_cast<size_t>(b) * 6;
    if (Hd->error) {
        if (tid == 0) {
            for (int i = 0; i < 5; ++i) dm[i] = 0;
            dm[5] = Hd->error;
        }
        return;
    }
    extern __shared__ int s_cnt[];   // n_stride + 1 column counters / offsets
    __shared__ int s_sb[kNumEE + 1];
    const int N = P.N, n = Hd->n, nf = Hd->nf, ns = Hd->n_samples, ne = Hd->n_eebox;
    ColCtx c;
    c.P = &P;
    c.Hd = Hd;
    c.nodes = reinterpret_cast<const NodeLin*>(ws + L.nodes);
    c.samples = reinterpret_cast<const Sample*>(ws + L.samples);
    c.eqs = reinterpret_cast<const EqRow*>(ws + L.eq);
    c.sb = s_sb;
    c.N = N;
    c.nf = nf;
    c.ns = ns;
    c.ne = ne;
    c.r_fb = kNx * (N + 1);
    c.r_cone = c.r_fb + 2 * ns;
    c.r_ee = c.r_cone + 4 * ns;
    c.r_td = c.r_ee + 2 * ne;
    const int m = c.r_td + Hd->n_eq;
    if (tid == 0) {
        int e = 0;
        s_sb[0] = 0;
        for (int j = 0; j < ns; ++j)
            while (c.samples[j].ee > e) s_sb[++e] = j;
        while (e < kNumEE) s_sb[++e] = ns;
    }
    __syncthreads();
    for (int j = tid; j < n; j += nth) {
        int cnt = 0;
        visit_column(c, j, [&](int, double v) { cnt += (v != 0.0); });
        s_cnt[j] = cnt;
    }
    __syncthreads();
    if (tid == 0) {   // exclusive scan (n <= 940; this is an export path, not the solve path)
        int acc = 0;
        for (int j = 0; j < n; ++j) {
            const int t = s_cnt[j];
            s_cnt[j] = acc;
            acc += t;
        }
        s_cnt[n] = acc;
        dm[0] = n;
        dm[1] = m;
        dm[2] = acc;
        dm[3] = kNx * (N + 1) + Hd->n_eq;
        dm[4] = Hd->m_ineq;
        dm[5] = (acc > nnz_cap) ? 8 : 0;
    }
    __syncthreads();
    const int nnz = s_cnt[n];
    for (int j = tid; j <= n; j += nth) cp[j] = s_cnt[j];
    if (nnz <= nnz_cap)
        for (int j = tid; j < n; j += nth) {
            int pos = s_cnt[j];
            visit_column(c, j, [&](int row, double v) {
                if (v != 0.0) {
                    ri[pos] = row;
                    va[pos] = v;
                    ++pos;
                }
            });
        }
    // cost: P = blkdiag(Q .. Q, Phi, force_cost I, 0) + 1e-3 I, q = [w .. w, Phi_w, 0]  (mpc.cpp:542-564,791-802,1090-1095)
    double* pd = p_diag + static_cast<size_t>(b) * n_stride;
    double* qo = q_out + static_cast<size_t>(b) * n_stride;
    for (int j = tid; j < n; j += nth) {
        const int k = j / kNx, i = j % kNx;
        double pv, qv;
        if (k < N) { pv = P.Q[i] + 1e-3; qv = P.w[i]; }
        else if (k == N) { pv = P.Phi[i] + 1e-3; qv = P.Phi_w[i]; }
        else { pv = ((j - kNx * (N + 1) < nf) ? P.force_cost : 0.0) + 1e-3; qv = 0.0; }
        pd[j] = pv;
        qo[j] = qv;
    }
    // right-hand side
    double* ub = ub_out + static_cast<size_t>(b) * m_stride;
    const double* x0 = reinterpret_cast<const double*>(ws + L.xoff);   // phi_0 = tangent(state)
    const double bx[2] = {Hd->ee_box[0] / 2, Hd->ee_box[1] / 2};
    for (int r = tid; r < m; r += nth) {
        double v;
        if (r < kNx) v = -x0[r];
        else if (r < c.r_fb) v = -c.nodes[r / kNx - 1].cd[r % kNx];
        else if (r < c.r_fb + ns) v = P.force_bound;
        else if (r < c.r_ee) v = 0.0;
        else if (r < c.r_td) {
            const int q = r - c.r_ee, neg = q >= ne, e2 = neg ? q - ne : q, cc = e2 & 1, foot = (e2 >> 1) & 3;
            v = neg ? -1 * (-bx[cc] + P.hip_xy[foot][cc]) : bx[cc] + P.hip_xy[foot][cc];
        } else v = c.eqs[r - c.r_td].rhs;
        ub[r] = v;
    }
}

void launch_export_csc(const Params& P, const WsLayout& L, const char* ws, int B, int32_t* colptr, int32_t* rowidx, double* val,
                       double* p_diag, double* q, double* ub, int32_t* dims, int n_stride, int m_stride, int nnz_cap,
                       cudaStream_t stream) {
    k_export_csc<<<B, 256, sizeof(int) * (n_stride + 1), stream>>>(P, L, ws, colptr, rowidx, val, p_diag, q, ub, dims, n_stride, m_stride, nnz_cap);
}

}  // namespace bgg
