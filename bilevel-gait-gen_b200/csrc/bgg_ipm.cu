// bilevel-gait-gen_b200 -- kernel 4: the QP solve, one CTA per MPC instance, a primal-dual interior-point method
// (Mehrotra predictor-corrector) on the condensed QP
//      min 1/2 u'Hu + g'u   s.t.  C u <= d  (force box, friction pyramid, foot box),   E u = e  (touch-down, foot start)
// The reference's live solver is Clarabel, an interior-point method run to 1e-8 (mpc.h:264, clarabel_interface.cpp:
// 18-27,72-155); an ADMM at OSQP tolerances leaves this ill-conditioned QP 1e-2 away from the optimum,
// so the kernel follows the live path.  Per iteration: K = H + C'WC + E'E/delta is assembled from the *structured*
// rows (never a sparse matrix), factorised by an in-shared-memory packed Cholesky, and used for the predictor and the
// corrector solve.  Equality rows are handled by the proximal (static-regularisation) term E'E/delta with their
// multipliers accumulated, as Clarabel does with its static KKT regularisation.
//
// Inequality rows, internal order (m = 6 ns + 2 ne):
//   6 j + 0 :  f_z(tau_j) <= force_bound          6 j + 1 : -f_z(tau_j) <= 0               (mpc.cpp:352-414)
//   6 j + 2..5 : (+-e_x - mu e_z).f <= 0, (+-e_y - mu e_z).f <= 0                          (mpc.cpp:153-209)
//   6 ns + 2 e + 0 : -p_c(k) + w.u_pos <=  hip_c + box_c/2     e = ((k-4)*4 + foot)*2 + c   (mpc_single_rigid_body.cpp:381-443)
//   6 ns + 2 e + 1 :  p_c(k) - w.u_pos <= -(hip_c - box_c/2)
#include "bgg_kernels.cuh"
#include "bgg_chol.cuh"
#include "bgg_kkt.cuh"
#include "bgg_kkt_mma.cuh"
#include "bgg_l2ops.cuh"

namespace bgg {

// Phase clocks (tools/profile_phases.py): compiled in only with -DBGG_IPM_PROF into a separate library, never into
// libbgg_b200.so.  Thread 0 of CTA 0 accumulates clock64() deltas between barrier-delimited phases.
#ifdef BGG_IPM_PROF
__device__ long long g_ipm_prof[32];
#define PROF_DECL __shared__ long long s_prof[32]; long long prof_t = 0; if (threadIdx.x == 0) { for (int i_ = 0; i_ < 32; ++i_) s_prof[i_] = 0; prof_t = clock64(); }
#define PROF(k) do { if (threadIdx.x == 0) { const long long t_ = clock64(); s_prof[k] += t_ - prof_t; prof_t = t_; } } while (0)
#define PROF_DUMP do { if (threadIdx.x == 0 && blockIdx.x == 0) for (int i_ = 0; i_ < 32; ++i_) g_ipm_prof[i_] = s_prof[i_]; } while (0)
#else
#define PROF_DECL
#define PROF(k) do { } while (0)
#define PROF_DUMP do { } while (0)
#endif

namespace {

struct Smem {
    double* K;       // lower triangle in 8 x 8 blocks (csrc/bgg_chol.cuh)
    double *u, *du, *rd, *rhs, *g, *tmpn;            // nu
    double *s, *lam, *ds, *dl, *rp, *wv, *d;         // m
    double *tkc, *ckc;                               // 2(N-3)
    double *nueq, *re, *dnu;                         // kMaxEq
    double* red;                                     // 72
    double* pw;                                      // [(N-3)*4][2] foot-box position weights
    int *pcnt, *poff;                                // [(N-3)*4]
    Sample* smp;                                     // staged force samples
    ColInfo* col;                                    // [nu] per-variable tables of the K assembly
    const double* phi;                               // position rows: shared memory when they fit, else HBM/L2
    int phi_stride;
};

}  // namespace

// What the operators below need.  They are force-inlined.  Measured on B200 (4096 solves): inlined 37.2 ms; as
// __noinline__ functions (one copy of each in the binary: ncu shows the 0.4 MB kernel stalling on instruction fetch) with
// the context by reference and every field copied to locals 41.1 ms, the same plus __isShared assumptions 39.2 ms,
// context by value 83 ms -- the call boundary costs more than the instruction-cache misses it saves.
struct IpmCtx {
    Smem S;
    const double* Hg;
    const EqRow* eq;
    const int *fbase, *pbase, *nfv, *npv;
    int N, nu, nf, ns, ne, neq, nkc;
    double mu_f;
};

// out[0..m) = C v   (v: nu-vector in shared memory)
static __device__ __forceinline__ void ipm_apply_C(const IpmCtx& c, const double* v, double* out) {
    const Smem S = c.S;
    const int *fbase = c.fbase, *pbase = c.pbase, *nfv = c.nfv, *npv = c.npv;
    __builtin_assume(__isShared(v)); __builtin_assume(__isShared(out)); __builtin_assume(__isShared(S.tkc));
    __builtin_assume(__isShared(S.smp)); __builtin_assume(__isShared(S.pw)); __builtin_assume(__isShared(S.pcnt));
    __builtin_assume(__isShared(S.poff)); __builtin_assume(__isShared(fbase)); __builtin_assume(__isShared(pbase));
    __builtin_assume(__isShared(nfv)); __builtin_assume(__isShared(npv));
    const int tid = threadIdx.x, nth = blockDim.x;
    const int nf = c.nf, ns = c.ns, ne = c.ne, nkc = c.nkc;
    const double mu_f = c.mu_f;
    l2_phi_rows_dot(S.phi, S.phi_stride, nkc, nf, smem_addr(v), smem_addr(S.tkc));   // dense position rows (csrc/bgg_l2ops.cuh)
    #pragma unroll 1
    for (int j = tid; j < ns; j += nth) {
        const Sample& sp = S.smp[j];
        double fv[3];
        for (int cc = 0; cc < 3; ++cc) {
            const double* vv = v + fbase[sp.ee] + cc * nfv[sp.ee] + sp.off;
            double s = 0;
            for (int i = 0; i < sp.cnt; ++i) s += sp.w[i] * vv[i];
            fv[cc] = s;
        }
        double* o = out + 6 * j;
        o[0] = fv[2];
        o[1] = -fv[2];
        o[2] = fv[0] - mu_f * fv[2];
        o[3] = -fv[0] - mu_f * fv[2];
        o[4] = fv[1] - mu_f * fv[2];
        o[5] = -fv[1] - mu_f * fv[2];
    }
    __syncthreads();
    #pragma unroll 1
    for (int e = tid; e < ne; e += nth) {
        const int cc = e & 1, foot = (e >> 1) & 3, kk = e >> 3, kf = kk * 4 + foot;
        const double* vv = v + nf + pbase[foot] + cc * npv[foot] + S.poff[kf];
        double s = -S.tkc[kk * 2 + cc];
        for (int i = 0; i < S.pcnt[kf]; ++i) s += S.pw[2 * kf + i] * vv[i];
        out[6 * ns + 2 * e] = s;
        out[6 * ns + 2 * e + 1] = -s;
    }
    __syncthreads();
}

// out[0..nu) += C' y   (y: m-vector; inactive rows carry y == 0).  Every output entry is owned by one thread.
static __device__ __forceinline__ void ipm_add_Ct(const IpmCtx& c, const double* y, double* out) {
    const Smem S = c.S;
    __builtin_assume(__isShared(y)); __builtin_assume(__isShared(out)); __builtin_assume(__isShared(S.ckc));
    __builtin_assume(__isShared(S.smp)); __builtin_assume(__isShared(S.pw)); __builtin_assume(__isShared(S.col));
    __builtin_assume(__isShared(S.poff));
    const int tid = threadIdx.x, nth = blockDim.x;
    const int nu = c.nu, nf = c.nf, ns = c.ns, nkc = c.nkc;
    const double mu_f = c.mu_f;
    #pragma unroll 1
    for (int q = tid; q < nkc; q += nth) {
        const int kk = q >> 1, cc = q & 1;
        double s = 0;
        for (int foot = 0; foot < kNumEE; ++foot) {
            const int e = (kk * 4 + foot) * 2 + cc;
            s += y[6 * ns + 2 * e] - y[6 * ns + 2 * e + 1];
        }
        S.ckc[q] = -s;
    }
    __syncthreads();
    l2_phi_cols_dot(S.phi, S.phi_stride, nkc, nf, smem_addr(S.ckc), smem_addr(out));   // dense foot-box rows (csrc/bgg_l2ops.cuh)
    // Two threads per column (nu <= 160 < blockDim / 2 ... else one): the first takes the dense foot-box rows (force
    // column) or the first half of the nodes (position column), the second the sample rows / the second half; the
    // loops run over the column's own sample / node range only (ColInfo, csrc/bgg_kkt.cuh).
    {
        const int half = (2 * nu <= nth) ? 2 : 1;
        #pragma unroll 1
        for (int base = 0; base < half * nu; base += nth) {   // whole warps iterate together: the partner exchange is a shuffle
            const int it = base + tid;
            const bool act = it < half * nu;
            const int col = act ? it / half : 0, part = it % half;
            const ColInfo ci = S.col[col];
            double s = 0;
            if (!act) {
            } else if (col < nf) {
                if (part == 1 || half == 1)
                    #pragma unroll 1
                    for (int j = ci.lo; j < ci.hi; ++j) {
                        const Sample& sp = S.smp[j];
                        const double* yy = y + 6 * j;
                        double coef;
                        if (ci.coord == 2) coef = (yy[0] - yy[1]) - mu_f * (yy[2] + yy[3] + yy[4] + yy[5]);
                        else if (ci.coord == 0) coef = yy[2] - yy[3];
                        else coef = yy[4] - yy[5];
                        s += coef * sp.w[ci.var - sp.off];
                    }
            } else {
                const int mid = (half == 2) ? (ci.lo + ci.hi + 1) / 2 : ci.hi;
                const int k0 = (part == 0) ? ci.lo : mid, k1 = (part == 0) ? mid : ci.hi;
                #pragma unroll 1
                for (int kk = k0; kk < k1; ++kk) {
                    const int kf = kk * 4 + ci.foot, e = kf * 2 + ci.coord;
                    s += (y[6 * ns + 2 * e] - y[6 * ns + 2 * e + 1]) * S.pw[2 * kf + (ci.var - S.poff[kf])];
                }
            }
            if (half == 2) s += __shfl_xor_sync(0xffffffffu, s, 1);   // partner thread: adjacent lane
            if (act && part == 0) out[col] += s;
        }
    }
    __syncthreads();
}

// out[0..nu) = H v, H full symmetric in HBM/L2, read column-wise (coalesced across threads)
static __device__ __forceinline__ void ipm_apply_H(const IpmCtx& c, const double* v, double* out) {
    l2_apply_H(c.Hg, c.nu, smem_addr(v), smem_addr(out));   // csrc/bgg_l2ops.cuh
}

// out = E v - e (or E v when with_rhs == false).  The two equality-row operators take scalars only and are out of line:
// nine call sites per Newton iteration, one copy in the binary.
static __device__ __noinline__ void ipm_apply_E(const EqRow* eq, int neq, const double* v, double* out, bool with_rhs) {
    const int tid = threadIdx.x;
    __builtin_assume(__isShared(v)); __builtin_assume(__isShared(out)); __builtin_assume(__isShared(eq));
    if (tid < neq) {
        const EqRow& q = eq[tid];
        double s = with_rhs ? -q.rhs : 0.0;
        for (int i = 0; i < q.cnt; ++i) s += q.w[i] * v[q.col[i]];
        out[tid] = s;
    }
    __syncthreads();
}

// out += scale E' y
static __device__ __noinline__ void ipm_add_Et(const EqRow* eq, int neq, const double* y, double* out, double scale) {
    const int tid = threadIdx.x;
    __builtin_assume(__isShared(y)); __builtin_assume(__isShared(out)); __builtin_assume(__isShared(eq));
    if (tid < kNumEE * 2)   // one thread per (foot, coord): rows of different groups touch different columns
#pragma unroll 1
        for (int r = 0; r < neq; ++r) {
            const EqRow& q = eq[r];
            if (q.pad != tid) continue;
            for (int i = 0; i < q.cnt; ++i) out[q.col[i]] += scale * y[r] * q.w[i];
        }
    __syncthreads();
}

struct IpmCaps {          // per-launch shared-memory sizing, from the actual maxima over the batch
    int nu, rows, ns, stage_phi;
};
static size_t ipm_smem_core(int N, int nu, int rows, int ns) {
    const size_t kc = 2 * (N - 3), eb = 4 * (N - 3);
    return 8 * (chol::doubles(nu / 8) + 6 * nu + 6 * static_cast<size_t>(rows) + 2 * eb * 2 + 2 * kc + 3 * kMaxEq + 72 + 2 * eb) +
           8 * eb + sizeof(Sample) * static_cast<size_t>(ns) + sizeof(ColInfo) * static_cast<size_t>(nu) + 64;
}
static IpmCaps ipm_caps(const WsLayout& L, int nu_max, int ns_max) {
    IpmCaps c;
    c.nu = (nu_max + 7) / 8 * 8;
    if (c.nu > L.max_nu) c.nu = L.max_nu;
    c.ns = ns_max < 1 ? 1 : ns_max;
    c.rows = 6 * c.ns + 2 * (L.N - 3) * 8;
    // the KKT assembly stages phi chunks and a per-sample table in the ds / dl vectors (csrc/bgg_kkt_mma.cuh)
    const int scratch = kkt_scratch_doubles(c.nu / 8, c.ns);
    if (2 * c.rows < scratch) c.rows = (scratch + 1) / 2;
    const size_t core = ipm_smem_core(L.N, c.nu, c.rows, c.ns);
    const size_t phi = 8 * static_cast<size_t>(2 * (L.N - 3)) * c.nu;
    // two CTAs per SM when the core fits twice into the 227 KB; the dense position rows are staged on chip only
    // if that does not cost the second CTA
    const size_t half = 113 * 1024;
    if (core <= half) c.stage_phi = (core + phi <= half) ? 1 : 0;
    else c.stage_phi = (core + phi <= 225 * 1024) ? 1 : 0;
    return c;
}
static size_t ipm_smem_for(const WsLayout& L, const IpmCaps& c) {
    return ipm_smem_core(L.N, c.nu, c.rows, c.ns) + (c.stage_phi ? 8 * static_cast<size_t>(2 * (L.N - 3)) * c.nu : 0);
}
size_t ipm_smem_bytes(const WsLayout& L) {   // worst case for the configured caps (bgg_create's feasibility check)
    return ipm_smem_core(L.N, L.max_nu, L.max_rows, kMaxSamples);
}

__global__ void __launch_bounds__(256, 2) k_ipm(Params P, WsLayout L, char* __restrict__ ws_base, int stage_phi, int cap_nu, int cap_rows) {
    const int b = blockIdx.x, tid = threadIdx.x, nth = blockDim.x;
    const int lane = tid & 31, wid = tid >> 5, nwarp = nth >> 5;
    char* ws = ws_base + static_cast<size_t>(b) * L.stride;
    WsHeader* Hd = reinterpret_cast<WsHeader*>(ws + L.hdr);
    if (Hd->error) {
        if (tid == 0) Hd->status = kOther;
        return;
    }
    const NodeLin* nodes = reinterpret_cast<const NodeLin*>(ws + L.nodes);
    const Sample* samples = reinterpret_cast<const Sample*>(ws + L.samples);
    const EqRow* eqs = reinterpret_cast<const EqRow*>(ws + L.eq);
    const double* Hg = reinterpret_cast<const double*>(ws + L.H);       // full symmetric nu x nu
    const double* gg = reinterpret_cast<const double*>(ws + L.g);
    const double* phipos = reinterpret_cast<const double*>(ws + L.phipos);
    const double* xoff = reinterpret_cast<const double*>(ws + L.xoff);

    const int N = P.N, nu = Hd->nu, nf = Hd->nf, ns = Hd->n_samples, ne = Hd->n_eebox, neq = Hd->n_eq;
    const int m = 6 * ns + 2 * ne, nkc = 2 * (N - 3), npk = nu * (nu + 1) / 2;
    const double cost_const = Hd->cost_const;
    const double mu_f = P.friction_coef, delta = P.ipm_eq_delta, inv_delta = 1.0 / P.ipm_eq_delta;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    PROF_DECL
    Smem S;
    {
        double* p = reinterpret_cast<double*>(smem_raw);
        S.K = p; p += chol::doubles(cap_nu >> 3);
        S.u = p; p += cap_nu; S.du = p; p += cap_nu; S.rd = p; p += cap_nu;
        S.rhs = p; p += cap_nu; S.g = p; p += cap_nu; S.tmpn = p; p += cap_nu;
        S.s = p; p += cap_rows; S.lam = p; p += cap_rows; S.ds = p; p += cap_rows; S.dl = p; p += cap_rows;
        S.rp = p; p += cap_rows; S.wv = p; p += cap_rows;
        S.d = p; p += 2 * 8 * (N - 3);   // right-hand sides of the foot-box rows only (force rows: see rhs_of)
        S.tkc = p; p += nkc; S.ckc = p; p += nkc;
        S.nueq = p; p += kMaxEq; S.re = p; p += kMaxEq; S.dnu = p; p += kMaxEq;
        S.red = p; p += 72;   // block_reduce scratch (33) / chol::solve scratch (64)
        S.pw = p; p += 2 * 4 * (N - 3);
        S.pcnt = reinterpret_cast<int*>(p);
        S.poff = S.pcnt + 4 * (N - 3);
        p += 4 * (N - 3);
        S.smp = reinterpret_cast<Sample*>(p);
        p += (sizeof(Sample) * static_cast<size_t>(ns) + 7) / 8;
        S.col = reinterpret_cast<ColInfo*>(p);
        p += (sizeof(ColInfo) * static_cast<size_t>(nu) + 7) / 8;
        if (stage_phi) {
            S.phi = p;
            S.phi_stride = nf;
            for (int i = tid; i < nkc * nf; i += nth) p[i] = phipos[static_cast<size_t>(i / nf) * L.max_nu + (i % nf)];
        } else {
            S.phi = phipos;
            S.phi_stride = L.max_nu;
        }
    }
    #pragma unroll 1
    for (int i = tid; i < ns * static_cast<int>(sizeof(Sample) / 8); i += nth)
        reinterpret_cast<double*>(S.smp)[i] = reinterpret_cast<const double*>(samples)[i];
    for (int i = tid; i < 4 * (N - 3); i += nth) {
        const NodeLin& nl = nodes[i / 4 + kEENodeStart];
        const int foot = i & 3;
        S.pcnt[i] = nl.pcnt[foot];
        S.poff[i] = nl.poff[foot];
        S.pw[2 * i] = nl.pw[foot][0];
        S.pw[2 * i + 1] = nl.pw[foot][1];
    }
    __shared__ int s_fbase[kNumEE], s_pbase[kNumEE], s_nfv[kNumEE], s_npv[kNumEE];
    __shared__ int s_flag, s_nitems, s_nwork, s_npos;
    __shared__ KktWork s_work[kMaxKktWork];
    __shared__ int s_sb[kNumEE + 1];   // per-foot sample ranges (samples are stored foot-major)
    __shared__ EqRow s_eq[kMaxEq];
    if (tid < neq) s_eq[tid] = eqs[tid];
    if (tid < kNumEE) {
        s_fbase[tid] = Hd->fbase[tid];
        s_pbase[tid] = Hd->pbase[tid];
        s_nfv[tid] = Hd->nfv[tid];
        s_npv[tid] = Hd->npv[tid];
    }
    #pragma unroll 1
    for (int i = tid; i < nu; i += nth) S.g[i] = gg[i];
    #pragma unroll 1
    for (int i = tid; i < 6 * cap_nu; i += nth)   // u du rd rhs g tmpn: the padding up to 8 nb stays zero (chol::solve)
        if (i % cap_nu >= nu) S.u[i] = 0.0;
    if (tid == 0) {
        int e = 0;
        s_sb[0] = 0;
        for (int j = 0; j < ns; ++j)
            while (samples[j].ee > e) s_sb[++e] = j;
        while (e < kNumEE) s_sb[++e] = ns;
    }

    // ---- right-hand sides d and the active mask (wv = 1 / 0 while setting up).  Force rows have the fixed pattern
    // (force_bound, 0, 0, 0, 0, 0) per sample; only the foot-box right-hand sides are stored.
    const double box0 = Hd->ee_box[0] / 2, box1 = Hd->ee_box[1] / 2;
    const double fbound = P.force_bound;
    const int m_force = 6 * ns;
    #pragma unroll 1
    for (int j = tid; j < ns; j += nth) {
        const bool act = samples[j].active != 0;
        for (int r = 0; r < 6; ++r) S.wv[6 * j + r] = act ? 1.0 : 0.0;
    }
    #pragma unroll 1
    for (int e = tid; e < ne; e += nth) {
        const int c = e & 1, foot = (e >> 1) & 3, kk = e >> 3;   // node k = kk + 4
        const double bx = c ? box1 : box0;
        const double off = xoff[(kk + kEENodeStart) * kNx + c];
        S.d[2 * e + 0] = (bx + P.hip_xy[foot][c]) + off;
        S.d[2 * e + 1] = -(-bx + P.hip_xy[foot][c]) - off;
        S.wv[m_force + 2 * e + 0] = 1.0;
        S.wv[m_force + 2 * e + 1] = 1.0;
    }
    auto rhs_of = [&](int i) -> double { return (i < m_force) ? ((i % 6 == 0) ? fbound : 0.0) : S.d[i - m_force]; };
    __syncthreads();
    kkt_build_colinfo(S.col, nu, nf, N, s_fbase, s_pbase, s_nfv, s_npv, s_sb, S.smp, S.pcnt, S.poff);
    __syncthreads();
    KktItem* kitems = reinterpret_cast<KktItem*>(ws + L.ktab);
    KktPos* kpos = reinterpret_cast<KktPos*>(ws + L.ktab + sizeof(KktItem) * kMaxKktItems);
    kkt_mma_setup(s_work, &s_nwork, (nu + 7) >> 3, kitems, &s_nitems, s_fbase, s_nfv, S.col, S.smp, kpos, &s_npos, nu, nf, s_eq, neq);
    if (s_nitems > kMaxKktItems || s_npos > kMaxKktPos) {   // cannot happen within max_spline_vars = 160; refuse rather than drop terms
        if (tid == 0) Hd->status = kOther;
        return;
    }
    __syncthreads();
    PROF(0);

    // ------------------------------------------------------------------------------------------------ operators
    IpmCtx ctx;
    ctx.S = S; ctx.Hg = Hg; ctx.eq = s_eq; ctx.fbase = s_fbase; ctx.pbase = s_pbase; ctx.nfv = s_nfv; ctx.npv = s_npv;
    ctx.N = N; ctx.nu = nu; ctx.nf = nf; ctx.ns = ns; ctx.ne = ne; ctx.neq = neq; ctx.nkc = nkc; ctx.mu_f = mu_f;
    auto apply_C = [&](const double* v, double* out) { PROF(10); ipm_apply_C(ctx, v, out); PROF(6); };
    auto add_Ct = [&](const double* y, double* out) { PROF(10); ipm_add_Ct(ctx, y, out); PROF(7); };
    auto apply_H = [&](const double* v, double* out) { PROF(10); ipm_apply_H(ctx, v, out); PROF(8); };
    auto apply_E = [&](const double* v, double* out, bool with_rhs) { PROF(10); ipm_apply_E(s_eq, neq, v, out, with_rhs); PROF(9); };
    auto add_Et = [&](const double* y, double* out, double scale) { PROF(10); ipm_add_Et(s_eq, neq, y, out, scale); PROF(9); };

    // K = H + C' diag(wv) C + E'E/delta in 8 x 8 blocks in shared memory (csrc/bgg_kkt_mma.cuh), then chol::factor in place.
    const int nb = (nu + 7) >> 3;   // 8 x 8 blocks per side; rows nu .. 8 nb - 1 are padded with the identity
    KktMma km;
    km.K = S.K; km.Hg = Hg; km.phig = phipos; km.phi_ld = L.max_nu; km.nu = nu; km.nf = nf; km.nb = nb; km.ns = ns; km.ne = ne;
    km.neq = neq; km.nkc = nkc; km.wv = S.wv; km.pw = S.pw; km.pcnt = S.pcnt; km.poff = S.poff; km.smp = S.smp; km.eq = s_eq;
    km.col = S.col; km.ckc = S.ckc; km.scratch = S.ds; km.work = s_work; km.nwork = s_nwork; km.items = kitems; km.nitems = s_nitems; km.pos = kpos; km.npos = s_npos;
    km.mu_f = mu_f; km.inv_delta = inv_delta;
    auto build_and_factor = [&]() -> bool {
        PROF(10);
        kkt_assemble_mma(km);   // csrc/bgg_kkt_mma.cuh (also writes the identity padding)
        PROF(1);
        chol::factor(S.K, nb, &s_flag);   // csrc/bgg_chol.cuh: DMMA block Cholesky, diagonal super-blocks inverted
        PROF(3);
        return s_flag == 0;
    };
    // Solve K x = v in place (v padded with zeros to 8 nb entries)
    auto chol_solve = [&](double* v) {
        PROF(10);
        chol::solve(S.K, nb, v, S.red);
        PROF(5);
    };

    // ------------------------------------------------------------------------------------------------ start point
    // u0 = argmin of the equality/inequality-penalised quadratic: (H + C'C + E'E/delta) u = -g + C'd + E'e/delta
    bool ok = build_and_factor();
    #pragma unroll 1
    for (int i = tid; i < nu; i += nth) S.rhs[i] = -S.g[i];
    #pragma unroll 1
    for (int i = tid; i < m; i += nth) S.rp[i] = (S.wv[i] != 0.0) ? rhs_of(i) : 0.0;
    if (tid < neq) S.re[tid] = s_eq[tid].rhs;
    __syncthreads();
    add_Ct(S.rp, S.rhs);
    add_Et(S.re, S.rhs, inv_delta);
    #pragma unroll 1
    for (int i = tid; i < nu; i += nth) S.u[i] = S.rhs[i];
    chol_solve(S.u);
    apply_C(S.u, S.ds);
    double mn = 1e300;
    #pragma unroll 1
    for (int i = tid; i < m; i += nth)
        if (S.wv[i] != 0.0) {
            S.s[i] = rhs_of(i) - S.ds[i];
            mn = fmin(mn, S.s[i]);
        }
    mn = block_reduce<kMin>(mn, S.red);
    const double shift = fmax(0.0, -1.5 * mn);
    double sl = 0, ss = 0, xi = 0;
    #pragma unroll 1
    for (int i = tid; i < m; i += nth) {
        if (S.wv[i] != 0.0) {
            const double v = fmax(S.s[i] + shift, 1e-2);
            S.s[i] = v;
            S.lam[i] = v;
            xi += v * v;
            sl += v;
        } else {
            S.s[i] = 1.0;
            S.lam[i] = 0.0;
        }
    }
    xi = block_reduce<kSum>(xi, S.red);
    sl = block_reduce<kSum>(sl, S.red);
    #pragma unroll 1
    for (int i = tid; i < m; i += nth)
        if (S.wv[i] != 0.0) {
            S.s[i] += 0.5 * xi / sl;
            ss += S.s[i];
        }
    ss = block_reduce<kSum>(ss, S.red);
    int m_act = 0;
    #pragma unroll 1
    for (int i = tid; i < m; i += nth)
        if (S.wv[i] != 0.0) {
            S.lam[i] += 0.5 * xi / ss;
            m_act++;
        }
    m_act = static_cast<int>(block_reduce<kSum>(static_cast<double>(m_act), S.red) + 0.5);
    if (tid < kMaxEq) S.nueq[tid] = 0.0;
    __syncthreads();

    double nrm_q = 1.0, nrm_d = 1.0;
    for (int r = 0; r < kNx; ++r) nrm_q = fmax(nrm_q, fmax(fabs(P.w[r]), fabs(P.Phi_w[r])));
    {
        double v = 0;
        #pragma unroll 1
        for (int i = tid; i < m; i += nth)
            if (S.wv[i] != 0.0) v = fmax(v, fabs(rhs_of(i)));
        nrm_d = fmax(1.0, block_reduce<kMax>(v, S.red));
    }

    // ------------------------------------------------------------------------------------------------ main loop
    int it = 0, status = kMaxIter;
    double mu_first = 0.0, rp_ref = 0.0;
    double n_rd = 0, n_rp = 0, n_re = 0, mu = 0, gap_scale = 1, qp_obj = 0, last_rp = 0, last_re = 0, last_rd = 0, last_mu = 0;
    for (it = 0; it <= P.ipm_max_iter; ++it) {
        // residuals: rd = H u + g + C'lam + E'nu ; rp = C u + s - d ; re = E u - e
        apply_H(S.u, S.rd);
        double pobj = 0;
        #pragma unroll 1
        for (int i = tid; i < nu; i += nth) {
            pobj += S.u[i] * (0.5 * S.rd[i] + S.g[i]);
            S.rd[i] += S.g[i];
        }
        pobj = block_reduce<kSum>(pobj, S.red);
        qp_obj = pobj;
        add_Ct(S.lam, S.rd);
        add_Et(S.nueq, S.rd, 1.0);
        // The primal residuals rp = C u + s - d and re = E u - e are linear in the iterate and the Newton step satisfies
        // C du + ds = -rp, E du = delta dnu - re by construction: after a step of length alpha they are (1 - alpha) rp and
        // re + alpha (delta dnu - re) to rounding, so they are updated with the step and recomputed from scratch only at the
        // first iteration and to confirm convergence (two structured products and their barriers less per iteration).
        if (it == 0) {
            apply_C(S.u, S.rp);
            apply_E(S.u, S.re, true);
        }
        double a = 0, c = 0, dsum = 0;
        #pragma unroll 1
        for (int i = tid; i < nu; i += nth) a = fmax(a, fabs(S.rd[i]));
        #pragma unroll 1
        for (int i = tid; i < m; i += nth) {
            if (S.lam[i] > 0.0) {   // active row (inactive rows keep lam == 0 exactly)
                if (it == 0) S.rp[i] = S.rp[i] + S.s[i] - rhs_of(i);
                c = fmax(c, fabs(S.rp[i]));
                dsum += S.s[i] * S.lam[i];
            } else {
                S.rp[i] = 0.0;
            }
        }
        {   // one pass for the three reductions: two maxima and a sum
            a = warp_max(a);
            c = warp_max(c);
            dsum = warp_sum(dsum);
            __syncthreads();
            if (lane == 0) {
                S.red[wid] = a;
                S.red[8 + wid] = c;
                S.red[16 + wid] = dsum;
            }
            __syncthreads();
            a = c = dsum = 0;
            for (int w = 0; w < nwarp; ++w) {
                a = fmax(a, S.red[w]);
                c = fmax(c, S.red[8 + w]);
                dsum += S.red[16 + w];
            }
            n_rd = a;
            n_rp = c;
        }
        mu = dsum / m_act;
        n_re = 0;
        for (int r = 0; r < neq; ++r) n_re = fmax(n_re, fabs(S.re[r]));
        gap_scale = fmax(1.0, fabs(pobj + cost_const));   // full objective, as Clarabel's relative gap
        const bool nan_seen = !(n_rd == n_rd) || !(n_rp == n_rp) || !(mu == mu);
        if (nan_seen || !ok) {
            status = kOther;
            n_rp = last_rp;
            n_re = last_re;
            n_rd = last_rd;
            mu = last_mu;
            break;
        }
        last_rp = n_rp;
        last_re = n_re;
        last_rd = n_rd;
        last_mu = mu;
        if (n_rd <= P.ipm_tol_feas * nrm_q && n_rp <= P.ipm_tol_feas * nrm_d && n_re <= P.ipm_tol_feas * nrm_d &&
            dsum <= P.ipm_tol_gap * gap_scale) {
            bool confirmed = true;
            if (it > 0) {   // confirm with primal residuals computed from scratch; they replace the updated ones either way
                apply_C(S.u, S.rp);
                apply_E(S.u, S.re, true);
                double cf = 0;
#pragma unroll 1
                for (int i = tid; i < m; i += nth) {
                    if (S.lam[i] > 0.0) {
                        S.rp[i] = S.rp[i] + S.s[i] - rhs_of(i);
                        cf = fmax(cf, fabs(S.rp[i]));
                    } else {
                        S.rp[i] = 0.0;
                    }
                }
                n_rp = block_reduce<kMax>(cf, S.red);
                n_re = 0;
                for (int r = 0; r < neq; ++r) n_re = fmax(n_re, fabs(S.re[r]));
                last_rp = n_rp;
                last_re = n_re;
                confirmed = n_rp <= P.ipm_tol_feas * nrm_d && n_re <= P.ipm_tol_feas * nrm_d;
            }
            if (confirmed) {
                status = kSolved;
                break;
            }
        }
        // Early exit on a stalled primal residual (what the infeasibility certificate of Clarabel's homogeneous embedding does
        // for the reference: an infeasible QP -- a quarter of the first solves of a random batch, a fifth of the line-search
        // candidates -- otherwise runs to the iteration limit and holds its SM for twice the time of a solved one).  The
        // primal residual shrinks by exactly (1 - alpha) per step: less than 10 % over ten iterations while still far from
        // feasible means the steps have collapsed.  The oracle applies the same rule (oracle/qp_ipm.cpp).
        {
            const double prim = fmax(n_rp, n_re);
            if (it == 0) rp_ref = prim;
            if (it > 0 && it % 10 == 0) {
                if (prim > 1e3 * P.ipm_tol_feas * nrm_d && prim >= 0.9 * rp_ref) break;   // classified at exit (PrimalInfeasible)
                rp_ref = prim;
            }
        }
        if (it == P.ipm_max_iter) break;

        // Iterative refinement of the corrector pays for itself only once W = lam / s has spread over many decades: it starts
        // when mu has fallen to ipm_refine_mu_frac of its first value (measured on 4096 instances: always 32.0 ms, never
        // 27.7 ms with 0.8 % more unsolved instances)
        if (it == 0) mu_first = mu;
        const bool refine_now = mu <= P.ipm_refine_mu_frac * mu_first;
        // scaling W = lam / s (inactive rows keep 0), factorisation
        #pragma unroll 1
        for (int i = tid; i < m; i += nth) S.wv[i] = (S.lam[i] > 0.0) ? S.lam[i] / S.s[i] : 0.0;
        __syncthreads();
        ok = build_and_factor();
        if (!ok) {   // residuals of this iteration are valid; classify at exit
            status = kOther;
            break;
        }

        // one Newton solve for complementarity target rc (dl holds -rc on entry, see callers):
        //   K du = -rd - C'((-rc + lam rp)/s) - E' re / delta ; ds = -rp - C du ; dl = (-rc - lam ds)/s
        auto newton = [&](bool corrector, double sig_mu) {
            #pragma unroll 1
            for (int i = tid; i < m; i += nth) {
                if (S.wv[i] == 0.0) {
                    S.dl[i] = 0.0;
                    continue;
                }
                double rc = S.s[i] * S.lam[i];
                if (corrector) rc += S.ds[i] * S.dl[i] - sig_mu;
                S.dl[i] = rc;                                           // keep rc
            }
            __syncthreads();
            #pragma unroll 1
            for (int i = tid; i < m; i += nth)
                S.ds[i] = (S.wv[i] != 0.0) ? -(-S.dl[i] + S.lam[i] * S.rp[i]) / S.s[i] : 0.0;
            #pragma unroll 1
            for (int i = tid; i < nu; i += nth) S.rhs[i] = -S.rd[i];
            __syncthreads();
            add_Ct(S.ds, S.rhs);
            add_Et(S.re, S.rhs, -inv_delta);
            #pragma unroll 1
            for (int i = tid; i < nu; i += nth) S.du[i] = S.rhs[i];
            chol_solve(S.du);
            // The predictor only steers the centring parameter sigma: it is solved without refinement.
            for (int rf = 0; rf < ((corrector && refine_now) ? P.ipm_refine : 0); ++rf) {
                // iterative refinement against K = H + C'WC + E'E/delta applied matrix-free
                apply_H(S.du, S.tmpn);
                apply_C(S.du, S.ds);
                #pragma unroll 1
                for (int i = tid; i < m; i += nth) S.ds[i] *= S.wv[i];
                __syncthreads();
                add_Ct(S.ds, S.tmpn);
                apply_E(S.du, S.dnu, false);
                add_Et(S.dnu, S.tmpn, inv_delta);
                #pragma unroll 1
                for (int i = tid; i < nu; i += nth) S.tmpn[i] = S.rhs[i] - S.tmpn[i];
                chol_solve(S.tmpn);
                #pragma unroll 1
                for (int i = tid; i < nu; i += nth) S.du[i] += S.tmpn[i];
                __syncthreads();
            }
            apply_C(S.du, S.ds);
            #pragma unroll 1
            for (int i = tid; i < m; i += nth) {
                if (S.wv[i] == 0.0) {
                    S.ds[i] = 0.0;
                    S.dl[i] = 0.0;
                    continue;
                }
                const double dsi = -S.rp[i] - S.ds[i];
                S.dl[i] = (-S.dl[i] - S.lam[i] * dsi) / S.s[i];
                S.ds[i] = dsi;
            }
            apply_E(S.du, S.dnu, false);
            if (tid < neq) S.dnu[tid] = (S.dnu[tid] + S.re[tid]) * inv_delta;
            __syncthreads();
        };
        auto max_step = [&]() -> double {
            double al = 1e300;
            #pragma unroll 1
            for (int i = tid; i < m; i += nth)
                if (S.wv[i] != 0.0) {
                    if (S.ds[i] < 0.0) al = fmin(al, -S.s[i] / S.ds[i]);
                    if (S.dl[i] < 0.0) al = fmin(al, -S.lam[i] / S.dl[i]);
                }
            return block_reduce<kMin>(al, S.red);
        };
        // predictor, then corrector: one copy of the Newton step in the binary (instruction-cache footprint)
        double sig_mu = 0.0, alpha = 1.0;
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {
            newton(pass == 1, sig_mu);
            const double amax = max_step();
            if (pass == 0) {
                const double a_aff = fmin(1.0, amax);
                double mu_aff = 0;
#pragma unroll 1
                for (int i = tid; i < m; i += nth)
                    if (S.wv[i] != 0.0) mu_aff += (S.s[i] + a_aff * S.ds[i]) * (S.lam[i] + a_aff * S.dl[i]);
                mu_aff = block_reduce<kSum>(mu_aff, S.red) / m_act;
                const double sr = mu_aff / mu;
                sig_mu = sr * sr * sr * mu;
            } else {
                alpha = fmin(1.0, 0.99 * amax);
            }
        }
        #pragma unroll 1
        for (int i = tid; i < nu; i += nth) S.u[i] += alpha * S.du[i];
        #pragma unroll 1
        for (int i = tid; i < m; i += nth)
            if (S.wv[i] != 0.0) {
                S.s[i] += alpha * S.ds[i];
                S.lam[i] += alpha * S.dl[i];
                S.rp[i] *= 1.0 - alpha;
            }
        if (tid < neq) {
            S.nueq[tid] += alpha * S.dnu[tid];
            S.re[tid] += alpha * (delta * S.dnu[tid] - S.re[tid]);
        }
        __syncthreads();
    }
    // Exit classification when the iteration stopped without meeting the tolerances (iteration limit, or a
    // factorisation / NaN breakdown once the scaling W = lam/s has blown up):
    //  * residuals within 1e3 x the tolerances         -> SolvedInacc (Clarabel's "AlmostSolved")
    //  * primal residual still far from feasible        -> PrimalInfeasible (multipliers diverge on an infeasible QP;
    //                                                      Clarabel certifies this through its homogeneous embedding)
    if (status == kMaxIter || status == kOther) {
        const double loose = 1e3;
        if (n_rd <= loose * P.ipm_tol_feas * nrm_q && n_rp <= loose * P.ipm_tol_feas * nrm_d &&
            n_re <= loose * P.ipm_tol_feas * nrm_d && mu * m_act <= loose * P.ipm_tol_gap * gap_scale)
            status = kSolvedInacc;
        else if (n_rp > loose * P.ipm_tol_feas * nrm_d || n_re > loose * P.ipm_tol_feas * nrm_d)
            status = kPrimalInfeasible;
    }

    // ------------------------------------------------------------------------------------------------ outputs
    double* uo = reinterpret_cast<double*>(ws + L.u);
    double* lo = reinterpret_cast<double*>(ws + L.lam);
    double* so = reinterpret_cast<double*>(ws + L.slack);
    double* no = reinterpret_cast<double*>(ws + L.nueq);
    #pragma unroll 1
    for (int i = tid; i < nu; i += nth) uo[i] = S.u[i];
    #pragma unroll 1
    for (int i = tid; i < m; i += nth) {
        lo[i] = S.lam[i];
        so[i] = (S.lam[i] > 0.0 || S.s[i] != 1.0) ? S.s[i] : rhs_of(i);   // inactive rows: slack = d (row is 0 <= d)
    }
    if (tid < neq) no[tid] = S.nueq[tid];
    PROF(10);
    PROF_DUMP;
    if (tid == 0) {
        Hd->status = status;
        Hd->iters = it;
        Hd->prim_res = fmax(n_rp, n_re);
        Hd->dual_res = n_rd;
        Hd->gap = mu * m_act;
        Hd->qp_cost = qp_obj + cost_const;
    }
    (void)delta;
    (void)npk;
}

void launch_ipm(const Params& P, const WsLayout& L, char* ws, int B, int nu_max, int ns_max, cudaStream_t stream) {
    const IpmCaps c = ipm_caps(L, nu_max, ns_max);
    const size_t smem = ipm_smem_for(L, c);
    static size_t configured = 0;
    if (smem > configured) {
        cudaFuncSetAttribute(k_ipm, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        configured = smem;
    }
    k_ipm<<<B, 256, smem, stream>>>(P, L, ws, c.stage_phi, c.nu, c.rows);
}

}  // namespace bgg

#ifdef BGG_IPM_PROF
extern "C" int bgg_debug_ipm_prof(long long* out32) {
    int rc = static_cast<int>(cudaMemcpyFromSymbol(out32, bgg::g_ipm_prof, 16 * sizeof(long long)));
    if (rc == 0) rc = static_cast<int>(cudaMemcpyFromSymbol(out32 + 16, bgg::chol::g_chol_prof, 16 * sizeof(long long)));
    if (rc == 0) rc = static_cast<int>(cudaMemcpyFromSymbol(out32 + 24, bgg::g_kkt_prof, 8 * sizeof(long long)));
    const long long zero[16] = {0};
    cudaMemcpyToSymbol(bgg::g_kkt_prof, zero, 8 * sizeof(long long));
    cudaMemcpyToSymbol(bgg::chol::g_chol_prof, zero, sizeof(zero));
    return rc;
}
#endif
