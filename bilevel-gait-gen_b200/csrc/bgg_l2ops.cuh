// bilevel-gait-gen_b200 -- the three pieces of the interior-point operators that read HBM / L2-resident matrices
// (the condensed Hessian H and the condensed position rows phi do not fit in shared memory next to K at two CTAs per
// SM).  In the inlined operators they were chains of dependent L2 round trips -- ncu: 17 % of the kernel's stall
// samples on three source lines -- because the interior-point loop around them sits at the 128-register cap and
// cannot hold a batch of loads.  Here each is a __noinline__ function with scalar / 32-bit shared-window arguments:
// its own register allocation, every load of a batch issued before the first use.
#pragma once
#include "bgg_kkt_mma.cuh"   // lds64 / sts64

namespace bgg {

// a global load the compiler may not sink next to its use: the point of these functions is that a whole batch is in
// flight before the first multiply
__device__ __forceinline__ double ldg64(const double* p) {
    double v;
    asm volatile("ld.global.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ double ldany64(const double* p) {   // phi is staged in shared memory when it fits
    double v;
    asm volatile("ld.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}

// out[i] = sum_j H[j][i] v[j]   (H symmetric, row stride nu: coalesced across i); with `subtract` out[i] -= that sum instead
// (the residual of the iterative refinement accumulates in place).  Two threads per column when the CTA has them.
// Every thread of the CTA; ends with a barrier.
static __device__ __noinline__ void l2_apply_H(const double* __restrict__ Hg, int nu, unsigned v_s, unsigned out_s, bool subtract) {
    const int tid = threadIdx.x, nth = blockDim.x;
    const int half = (2 * nu <= nth) ? 2 : 1;
    for (int base = 0; base < half * nu; base += nth) {
        const int it = base + tid;
        const bool act = it < half * nu;
        const int part = (act && it >= nu && half == 2) ? 1 : 0, i = act ? it - part * nu : 0;
        const int j0 = (half == 2) ? (part * nu) / 2 : 0, j1 = (half == 2) ? ((part + 1) * nu) / 2 : nu;
        double s0 = 0.0, s1 = 0.0;
        if (act) {
            const double* col = Hg + i;
            for (int jb = j0; jb < j1; jb += 20) {
                double hv[20];
#pragma unroll
                for (int k = 0; k < 20; ++k) hv[k] = ldg64(col + static_cast<size_t>(min(jb + k, j1 - 1)) * nu);   // clamped: no branch per load
#pragma unroll
                for (int k = 0; k < 20; k += 2) {
                    if (jb + k < j1) s0 = fma(hv[k], lds64(v_s + 8u * (jb + k)), s0);
                    if (jb + k + 1 < j1) s1 = fma(hv[k + 1], lds64(v_s + 8u * (jb + k + 1)), s1);
                }
            }
        }
        const double s = subtract ? -(s0 + s1) : s0 + s1;
        if (half == 2) {   // the two halves of a column sit nu threads apart: combine through shared memory
            if (act && part == 1) sts64(out_s + 8u * i, subtract ? lds64(out_s + 8u * i) + s : s);
            __syncthreads();
            if (act && part == 0) sts64(out_s + 8u * i, lds64(out_s + 8u * i) + s);
        } else if (act) {
            sts64(out_s + 8u * i, subtract ? lds64(out_s + 8u * i) + s : s);
        }
    }
    __syncthreads();
}

// tkc[q] = sum_{i < nf} phi[q][i] v[i]  for q < nkc.  One warp per row, the loads of up to five rows in flight together.
// Every thread of the CTA; no barrier at the end (the caller has one before tkc is read).
static __device__ __noinline__ void l2_phi_rows_dot(const double* __restrict__ phi, int stride, int nkc, int nf, unsigned v_s, unsigned tkc_s) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nwarp = blockDim.x >> 5;
    double vv[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) vv[k] = (lane + 32 * k < nf) ? lds64(v_s + 8u * (lane + 32 * k)) : 0.0;   // nf <= 160
    for (int q0 = wid; q0 < nkc; q0 += 5 * nwarp) {
        double part[5];
#pragma unroll
        for (int r = 0; r < 5; ++r) {
            const int q = q0 + r * nwarp;
            double ph[5];
#pragma unroll
            for (int k = 0; k < 5; ++k) ph[k] = (q < nkc && lane + 32 * k < nf) ? phi[static_cast<size_t>(q) * stride + lane + 32 * k] : 0.0;
            double sacc = 0.0;
#pragma unroll
            for (int k = 0; k < 5; ++k) sacc = fma(ph[k], vv[k], sacc);
            part[r] = sacc;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int r = 0; r < 5; ++r) part[r] += __shfl_xor_sync(0xffffffffu, part[r], o);
        }
#pragma unroll
        for (int r = 0; r < 5; ++r) {
            const int q = q0 + r * nwarp;
            if (lane == 0 && q < nkc) sts64(tkc_s + 8u * q, part[r]);
        }
    }
}

// out[col] += sum_{q < nkc} ckc[q] phi[q][col]  for col < nf.  Thread (col, part) = (it / 2, it % 2) takes half of the
// rows, all of its loads issued before the first use; partners are adjacent lanes.  Every thread of the CTA; ends with
// a barrier.
static __device__ __noinline__ void l2_phi_cols_dot(const double* __restrict__ phi, int stride, int nkc, int nf, unsigned ckc_s, unsigned out_s) {
    const int tid = threadIdx.x, nth = blockDim.x;
    for (int base = 0; base < 2 * nf; base += nth) {   // whole warps iterate together (shuffle below)
        const int it = base + tid;
        const bool act = it < 2 * nf;
        const int col = act ? it >> 1 : 0, part = it & 1;
        const int qh = (nkc + 1) >> 1, q0 = part ? qh : 0, q1 = part ? nkc : qh;
        double s0 = 0.0, s1 = 0.0;
        if (act) {
            const double* pc = phi + col;
            for (int qb = q0; qb < q1; qb += 24) {
                double ph[24];
#pragma unroll
                for (int k = 0; k < 24; ++k) ph[k] = ldany64(pc + static_cast<size_t>(min(qb + k, q1 - 1)) * stride);   // clamped: no branch per load
#pragma unroll
                for (int k = 0; k < 24; k += 2) {
                    if (qb + k < q1) s0 = fma(lds64(ckc_s + 8u * (qb + k)), ph[k], s0);
                    if (qb + k + 1 < q1) s1 = fma(lds64(ckc_s + 8u * (qb + k + 1)), ph[k + 1], s1);
                }
            }
        }
        double s = s0 + s1;
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        if (act && part == 0) sts64(out_s + 8u * col, lds64(out_s + 8u * col) + s);
    }
    __syncthreads();
}

}  // namespace bgg
