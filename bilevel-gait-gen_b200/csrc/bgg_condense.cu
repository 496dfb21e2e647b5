// bilevel-gait-gen_b200 -- kernel 3b: state elimination ("condensing") of the arrow-shaped RTI QP, one CTA per
// instance, everything on chip except the outputs.
//
// The reference's decision vector is z = [x_0 .. x_N | u] with the spline coefficients u entering every node's
// dynamics row (mpc_single_rigid_body.cpp:259-273): a block-bidiagonal state chain with a dense border.  With
//   x_0 = x_init,   x_{k+1} = Ad_k x_k + Bd_k u + cd_k          (rows of AddDynamicsConstraints, :218-265)
// every state is affine in u:  x_k = Phi_k u + phi_k,  Phi_{k+1} = Ad_k Phi_k + Bd_k,  phi_{k+1} = Ad_k phi_k + cd_k.
// The cost (P diagonal: Q per node, Phi at node N, force weight, +1e-3 I; mpc.cpp:542-564,791-802,1090-1095) becomes
//   H = sum_k Phi_k' P_k Phi_k + P_u,     g = sum_k Phi_k' (P_k phi_k + q_k).
// Outputs: H (full symmetric, HBM), g, the two position rows of Phi_k for the foot-box rows (k >= 4), and phi_k.
// Phi_k itself (12 x nu) lives in shared memory only, double buffered; H is accumulated in shared memory as a
// packed lower triangle.
#include "bgg_kernels.cuh"

namespace bgg {

__device__ __forceinline__ void cross3d(const double a[3], const double b[3], double o[3]) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}

static size_t condense_smem_for(int nu_cap) {
    const size_t nu = nu_cap;
    return 8 * (nu * (nu + 1) / 2 + 2 * kNx * nu + nu + 4 * kNx) + sizeof(NodeLin) + 256;
}
size_t condense_smem_bytes(const WsLayout& L) {
    const size_t nu = L.max_nu;
    return 8 * (nu * (nu + 1) / 2 + 2 * kNx * nu + nu + 4 * kNx) + sizeof(NodeLin) + 256;
}

__global__ void __launch_bounds__(256) k_condense(Params P, WsLayout L, char* __restrict__ ws_base, int cap_nu) {
    const int b = blockIdx.x, tid = threadIdx.x, nth = blockDim.x;
    char* ws = ws_base + static_cast<size_t>(b) * L.stride;
    WsHeader* Hd = reinterpret_cast<WsHeader*>(ws + L.hdr);
    if (Hd->error) return;
    const NodeLin* nodes = reinterpret_cast<const NodeLin*>(ws + L.nodes);
    const double* zprev = reinterpret_cast<const double*>(ws + L.zprev);
    double* Hout = reinterpret_cast<double*>(ws + L.H);
    double* gout = reinterpret_cast<double*>(ws + L.g);
    double* phipos = reinterpret_cast<double*>(ws + L.phipos);
    double* xoff = reinterpret_cast<double*>(ws + L.xoff);

    const int N = P.N, nu = Hd->nu, nf = Hd->nf;
    const int npk = nu * (nu + 1) / 2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* Hp = reinterpret_cast<double*>(smem_raw);          // packed lower triangle
    double* Phi = Hp + cap_nu * (cap_nu + 1) / 2;              // [2][12][nu]
    double* gs = Phi + 2 * kNx * cap_nu;                       // [nu]
    double* phi = gs + cap_nu;                                 // [2][12]
    double* pq = phi + 2 * kNx;                                // [2][12]: P_k (diag) and P_k phi_k + q_k
    NodeLin* nl = reinterpret_cast<NodeLin*>(pq + 2 * kNx);
    __shared__ int s_fbase[kNumEE], s_pbase[kNumEE], s_nfv[kNumEE], s_npv[kNumEE];
    __shared__ double s_cc[kNx];
    double cc = 0.0;   // this thread's share of the constant term of the condensed objective

    for (int i = tid; i < npk; i += nth) Hp[i] = 0.0;
    for (int i = tid; i < 2 * kNx * nu; i += nth) Phi[i] = 0.0;
    for (int i = tid; i < nu; i += nth) gs[i] = 0.0;
    if (tid < kNx) phi[tid] = zprev[tid];   // phi_0 = tangent(state), the right-hand side of the -x_0 row
    if (tid < kNumEE) {
        s_fbase[tid] = Hd->fbase[tid];
        s_pbase[tid] = Hd->pbase[tid];
        s_nfv[tid] = Hd->nfv[tid];
        s_npv[tid] = Hd->npv[tid];
    }
    __syncthreads();

    for (int k = 0; k <= N; ++k) {
        double* Pc = Phi + (k & 1) * kNx * nu;          // Phi_k, row-major [12][nu]
        double* Pn = Phi + ((k + 1) & 1) * kNx * nu;    // Phi_{k+1}
        double* fc = phi + (k & 1) * kNx;
        double* fn = phi + ((k + 1) & 1) * kNx;
        if (k < N) {
            const double* src = reinterpret_cast<const double*>(&nodes[k]);
            double* dst = reinterpret_cast<double*>(nl);
            for (int i = tid; i < static_cast<int>(sizeof(NodeLin) / 8); i += nth) dst[i] = src[i];
        }
        if (tid < kNx) {
            const double pk = ((k < N) ? P.Q[tid] : P.Phi[tid]) + 1e-3;
            const double qk = (k < N) ? P.w[tid] : P.Phi_w[tid];
            pq[tid] = pk;
            pq[kNx + tid] = pk * fc[tid] + qk;
            cc += 0.5 * pk * fc[tid] * fc[tid] + qk * fc[tid];
            xoff[k * kNx + tid] = fc[tid];
        }
        __syncthreads();
        // g += Phi_k' (P_k phi_k + q_k) ; foot-box position rows
        for (int i = tid; i < nu; i += nth) {
            double s = 0;
#pragma unroll
            for (int r = 0; r < kNx; ++r) s += Pc[r * nu + i] * pq[kNx + r];
            gs[i] += s;
            if (k >= kEENodeStart) {
                phipos[static_cast<size_t>((k - kEENodeStart) * 2 + 0) * L.max_nu + i] = Pc[0 * nu + i];
                phipos[static_cast<size_t>((k - kEENodeStart) * 2 + 1) * L.max_nu + i] = Pc[1 * nu + i];
            }
        }
        // H += Phi_k' P_k Phi_k  (lower triangle; 16x16 thread tiling over (i, j))
        if (k > 0) {
            const int ty = tid >> 4, tx = tid & 15;
            for (int i = ty; i < nu; i += 16) {
                double pi[kNx];
#pragma unroll
                for (int r = 0; r < kNx; ++r) pi[r] = pq[r] * Pc[r * nu + i];
                const int rowbase = i * (i + 1) / 2;
                for (int j = tx; j <= i; j += 16) {
                    double s = 0;
#pragma unroll
                    for (int r = 0; r < kNx; ++r) s += pi[r] * Pc[r * nu + j];
                    Hp[rowbase + j] += s;
                }
            }
        }
        // Phi_{k+1} = Ad_k Phi_k + Bd_k ; phi_{k+1} = Ad_k phi_k + cd_k
        if (k < N) {
            for (int i = tid; i < nu; i += nth) {
                double col[kNx];
#pragma unroll
                for (int r = 0; r < kNx; ++r) col[r] = Pc[r * nu + i];
#pragma unroll
                for (int r = 0; r < kNx; ++r) {
                    double s = 0;
#pragma unroll
                    for (int q = 0; q < kNx; ++q) s += nl->Ad[r * kNx + q] * col[q];
                    Pn[r * nu + i] = s;
                }
            }
            if (tid < kNx) {
                double s = nl->cd[tid];
                for (int q = 0; q < kNx; ++q) s += nl->Ad[tid * kNx + q] * fc[q];
                fn[tid] = s;
            }
            __syncthreads();
            // + Bd_k : one thread per (foot, coord, weight); distinct threads touch distinct columns
            if (tid < kNumEE * 3 * 4) {
                const int e = tid / 12, c = (tid / 4) % 3, j = tid % 4;
                if (j < nl->fcnt[e]) {
                    const int col = s_fbase[e] + c * s_nfv[e] + nl->foff[e] + j;
                    const double ec[3] = {c == 0 ? 1.0 : 0.0, c == 1 ? 1.0 : 0.0, c == 2 ? 1.0 : 0.0};
                    double rc[3];
                    cross3d(nl->rel[e], ec, rc);
                    Pn[(3 + c) * nu + col] += P.dt * nl->fw[e][j];
                    for (int r = 0; r < 3; ++r) Pn[(9 + r) * nu + col] += P.dt * (rc[r] * nl->fw[e][j]);
                }
            } else if (tid >= 64 && tid < 64 + kNumEE * 2 * 2) {
                const int q = tid - 64, e = q / 4, c = (q / 2) % 2, j = q % 2;
                if (j < nl->pcnt[e]) {
                    const int col = nf + s_pbase[e] + c * s_npv[e] + nl->poff[e] + j;
                    const double ec[3] = {c == 0 ? 1.0 : 0.0, c == 1 ? 1.0 : 0.0, 0.0};
                    double ef[3];
                    cross3d(ec, nl->f[e], ef);
                    for (int r = 0; r < 3; ++r) Pn[(9 + r) * nu + col] += P.dt * (ef[r] * nl->pw[e][j]);
                }
            }
        }
        __syncthreads();
    }
    if (tid < kNx) s_cc[tid] = cc;
    // P_u: force weight on force variables, +1e-3 on everything (AddForceCost / AddDiagonalCost)
    for (int i = tid; i < nu; i += nth) {
        Hp[i * (i + 1) / 2 + i] += ((i < nf) ? P.force_cost : 0.0) + 1e-3;
        gout[i] = gs[i];
    }
    __syncthreads();
    if (tid == 0) {
        double t = 0;
        for (int r = 0; r < kNx; ++r) t += s_cc[r];
        Hd->cost_const = t;
    }
    // full symmetric H to HBM (the IPM reads it column-wise, coalesced)
    for (int p = tid; p < nu * nu; p += nth) {
        const int i = p / nu, j = p % nu;
        Hout[p] = (j <= i) ? Hp[i * (i + 1) / 2 + j] : Hp[j * (j + 1) / 2 + i];
    }
}

void launch_condense(const Params& P, const WsLayout& L, char* ws, int B, int nu_max, cudaStream_t stream) {
    int cap = (nu_max + 7) / 8 * 8;
    if (cap > L.max_nu) cap = L.max_nu;
    const size_t smem = condense_smem_for(cap);
    static size_t configured = 0;
    if (smem > configured) {
        cudaFuncSetAttribute(k_condense, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        configured = smem;
    }
    k_condense<<<B, 256, smem, stream>>>(P, L, ws, cap);
}

__global__ void k_batch_max(WsLayout L, const char* __restrict__ ws, int B, int* __restrict__ out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const WsHeader* h = reinterpret_cast<const WsHeader*>(ws + static_cast<size_t>(b) * L.stride + L.hdr);
    if (h->error) return;
    atomicMax(&out[0], h->nu);
    atomicMax(&out[1], h->n_samples);
}

void launch_batch_max(const WsLayout& L, const char* ws, int B, int* out, cudaStream_t stream) {
    cudaMemsetAsync(out, 0, 2 * sizeof(int), stream);
    k_batch_max<<<(B + 255) / 256, 256, 0, stream>>>(L, ws, B, out);
}

// ---- parity tap: dense Ad [N][12][12], Bd [N][12][nu_stride], cd [N][12] per instance (what the reference holds in
// A_, B_, C_ after mpc_single_rigid_body.cpp:246-248)
__global__ void k_export_dynamics(Params P, WsLayout L, const char* __restrict__ ws_base, double* __restrict__ Ad,
                                  double* __restrict__ Bd, double* __restrict__ cd, int nu_stride) {
    const int b = blockIdx.x, tid = threadIdx.x, nth = blockDim.x;
    const char* ws = ws_base + static_cast<size_t>(b) * L.stride;
    const WsHeader* Hd = reinterpret_cast<const WsHeader*>(ws + L.hdr);
    const NodeLin* nodes = reinterpret_cast<const NodeLin*>(ws + L.nodes);
    const int N = P.N, nf = Hd->nf;
    double* A = Ad + static_cast<size_t>(b) * N * 144;
    double* Bm = Bd + static_cast<size_t>(b) * N * kNx * nu_stride;
    double* c = cd + static_cast<size_t>(b) * N * kNx;
    for (int i = tid; i < N * kNx * nu_stride; i += nth) Bm[i] = 0.0;
    __syncthreads();
    for (int i = tid; i < N * 144; i += nth) A[i] = nodes[i / 144].Ad[i % 144];
    for (int i = tid; i < N * kNx; i += nth) c[i] = nodes[i / kNx].cd[i % kNx];
    for (int i = tid; i < N * kNumEE; i += nth) {
        const int k = i / kNumEE, e = i % kNumEE;
        const NodeLin& nl = nodes[k];
        double* Bk = Bm + static_cast<size_t>(k) * kNx * nu_stride;
        for (int cc = 0; cc < 3; ++cc) {
            const double ec[3] = {cc == 0 ? 1.0 : 0.0, cc == 1 ? 1.0 : 0.0, cc == 2 ? 1.0 : 0.0};
            double rc[3], ef[3];
            cross3d(nl.rel[e], ec, rc);
            cross3d(ec, nl.f[e], ef);
            for (int j = 0; j < nl.fcnt[e]; ++j) {
                const int col = Hd->fbase[e] + cc * Hd->nfv[e] + nl.foff[e] + j;
                Bk[(3 + cc) * nu_stride + col] = P.dt * nl.fw[e][j];
                for (int r = 0; r < 3; ++r) Bk[(9 + r) * nu_stride + col] = P.dt * (rc[r] * nl.fw[e][j]);
            }
            if (cc < 2)
                for (int j = 0; j < nl.pcnt[e]; ++j) {
                    const int col = nf + Hd->pbase[e] + cc * Hd->npv[e] + nl.poff[e] + j;
                    for (int r = 0; r < 3; ++r) Bk[(9 + r) * nu_stride + col] = P.dt * (ef[r] * nl.pw[e][j]);
                }
        }
    }
}

void launch_export_dynamics(const Params& P, const WsLayout& L, char* ws, int B, double* Ad, double* Bd, double* cd,
                            int nu_stride, cudaStream_t stream) {
    k_export_dynamics<<<B, 256, 0, stream>>>(P, L, ws, Ad, Bd, cd, nu_stride);
}

}  // namespace bgg
