// bilevel-gait-gen_b200 -- Cholesky factorisation and solves of the reduced KKT matrix in shared memory, on the FP64
// tensor-core path (mma.sync.m8n8k4.f64, "DMMA") of sm_100a.
//
// Measured on B200 (tools/microbench.cu, profiles/): a dependent DFMA takes 8 cycles, a DMMA 26 cycles for 256 FMAs,
// both pipes peak at 64 FMA/clk/SM; the scalar version of this factorisation issued ten integer / load instructions
// per FMA and sat on dependent chains of one or two lanes.  Here:
//   * storage: the lower triangle as 8 x 8 row-major blocks, block (bi, bj) at ((bi (bi + 1) / 2 + bj) * 64 doubles;
//   * one 16-byte shared-memory load per lane fetches an operand fragment of a block: lane (g, t) = (lane / 4, lane % 4)
//     holds entries [g][2t], [g][2t+1].  With the contraction index permuted (first DMMA k = {0,2,4,6}, second
//     k = {1,3,5,7}) this one layout is the A fragment of P, the B fragment of Q' in P Q', and the C fragment -- so a
//     product that has just been accumulated feeds the next DMMA from registers, no shuffles and no shared-memory trip;
//   * left-looking by block columns with look-ahead: warp 0 runs the chain of 8 x 8 diagonal blocks (the only serial
//     part: one lane, eight pivots per block, factor and inverse in registers) with nothing but its own block row
//     between two of them; warps 1.. multiply their rows of column j by the inverted block (a DMMA, not a
//     substitution), apply column j to column j + 1 and, after a barrier among themselves, columns <= j to column
//     j + 2.  One CTA-wide barrier per block column.
//   * the inverses of the diagonal blocks are then merged in place into inverses of 64 x 64 diagonal super-blocks, so
//     that a triangular solve is two (not 120, not 15) dependent steps of warp-per-block-row mat-vecs.
// The first stage of the recursion K -> L is the textbook one; nothing here follows reference code (the reference
// hands its KKT system to Clarabel's sparse LDL, clarabel_interface.cpp:68-75).
#pragma once
#include <cuda_runtime.h>

namespace bgg {
namespace chol {

#ifdef BGG_IPM_PROF   // tools/profile_phases.py only: clock64 deltas of thread 0 of CTA 0 inside factor()
static __device__ long long g_chol_prof[16];
#define CPROF_DECL long long cprof_t = clock64();
#define CPROF(k) do { if (threadIdx.x == 0 && blockIdx.x == 0) { const long long t_ = clock64(); g_chol_prof[k] += t_ - cprof_t; cprof_t = t_; } } while (0)
#else
#define CPROF_DECL
#define CPROF(k) do { } while (0)
#endif

__device__ __forceinline__ int blk(int bi, int bj) { return (((bi * (bi + 1)) >> 1) + bj) << 6; }
__device__ __forceinline__ int at(int i, int j) { return blk(i >> 3, j >> 3) + ((i & 7) << 3) + (j & 7); }
__host__ __device__ inline size_t doubles(int nb) { return static_cast<size_t>(nb) * (nb + 1) / 2 * 64; }

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
#ifdef BGG_NO_DMMA   // timing experiment only (wrong results): one DFMA instead of the tensor-core instruction
    c0 = fma(a, b, c0);
    return;
#endif
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// entries [g][2t], [g][2t+1] of a row-major 8 x 8 block
__device__ __forceinline__ double2 ldfrag(const double* b, int lane) {
    return *reinterpret_cast<const double2*>(b + ((lane >> 2) << 3) + ((lane & 3) << 1));
}
__device__ __forceinline__ void stfrag(double* b, int lane, double2 v) {
    *reinterpret_cast<double2*>(b + ((lane >> 2) << 3) + ((lane & 3) << 1)) = v;
}
// entries [2t][g], [2t+1][g]: the fragment of the transposed block
__device__ __forceinline__ double2 ldfragT(const double* b, int lane) {
    const int g = lane >> 2, t = lane & 3;
    return make_double2(b[(2 * t) * 8 + g], b[(2 * t + 1) * 8 + g]);
}
// acc += P Q'   (p, q: fragments of P and Q)          acc += P Q   (p fragment of P, qT transposed fragment of Q)
__device__ __forceinline__ void mma_pqT(double2& acc, double2 p, double2 q) {
    dmma(acc.x, acc.y, p.x, q.x);
    dmma(acc.x, acc.y, p.y, q.y);
}

// 1 / sqrt(d), d positive and normal: hardware double-precision seed refined by two Newton steps
__device__ __forceinline__ double fast_rsqrt(double d) {
    double x = static_cast<double>(rsqrtf(static_cast<float>(d)));
    const double h = 0.5 * d;
    x = x * (1.5 - h * x * x);   // 2^-22 -> 2^-43
    x = x * (1.5 - h * x * x);   // -> below double rounding
    return x;
}

// 1 / sqrt(d), d positive and normal: the double-precision hardware seed (MUFU.RSQ64H, no conversions through float)
// and one third-order correction  x (1 + e/2 + 3 e^2 / 8),  e = 1 - d x^2  -- five dependent FP64 operations on the
// pivot chain instead of the eight (plus two conversions) of fast_rsqrt.  Seed and result accuracy: tools/microbench_diag.cu.
__device__ __forceinline__ double rsqrt_fast(double d) {
    double x;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
    const double t = d * x;
    const double e = fma(-t, x, 1.0);
    const double p = fma(0.375, e, 0.5);
    const double xe = x * e;
    x = fma(xe, p, x);
#ifdef BGG_RSQRT_TWO_STEPS
    {
        const double t2 = d * x;
        const double e2 = fma(-t2, x, 1.0);
        x = fma(x * e2, 0.5, x);
    }
#endif
    return x;
}

// One warp; only lane 0 works (the pivots are a serial chain; shuffles or shared-memory hand-offs between lanes cost
// more than the arithmetic they would spread).  D: 8 x 8 row-major diagonal block (lower triangle valid).  On exit D
// holds X = inv(L), L the Cholesky factor of the block, with an explicit zero upper triangle.  Everything stays in
// registers: right-looking factorisation, and as soon as pivot c is known row c of L is replaced by row c of X
//   X[c][k] = -(1 / L_cc) sum_{m = k}^{c-1} L[c][m] X[m][k]
// whose sums do not depend on pivot c and are scheduled into the latency of its reciprocal square root.
// Returns (on lane 0) whether a non-positive pivot was met.
__device__ __forceinline__ bool factor_invert_diag(double* D, int lane) {
    bool bad_any = false;
    if (lane == 0) {
        double a[8][8];
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int q = 0; 2 * q <= r; ++q) {
                const double2 v = *reinterpret_cast<const double2*>(D + r * 8 + 2 * q);
                a[r][2 * q] = v.x;
                if (2 * q + 1 <= r) a[r][2 * q + 1] = v.y;
            }
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const double d = a[c][c];
            const bool bad = !(d > 0.0);
            bad_any |= bad;
            double tmp[8];
#pragma unroll
            for (int k = 0; k < c; ++k) {   // independent of this pivot
                double sm = a[c][k] * a[k][k];
#pragma unroll
                for (int m = k + 1; m < c; ++m) sm += a[c][m] * a[m][k];
                tmp[k] = sm;
            }
            const double inv = bad ? 1.0 : rsqrt_fast(d);
            a[c][c] = inv;
#pragma unroll
            for (int r = c + 1; r < 8; ++r) a[r][c] *= inv;
#pragma unroll
            for (int c2 = c + 1; c2 < 8; ++c2)
#pragma unroll
                for (int r = c2; r < 8; ++r) a[r][c2] -= a[r][c] * a[c2][c];
#pragma unroll
            for (int k = 0; k < c; ++k) a[c][k] = -inv * tmp[k];
        }
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                double2 v;
                v.x = (2 * q <= r) ? a[r][2 * q] : 0.0;
                v.y = (2 * q + 1 <= r) ? a[r][2 * q + 1] : 0.0;
                *reinterpret_cast<double2*>(D + r * 8 + 2 * q) = v;
            }
    }
    __syncwarp();
    return bad_any;
}

// (first version, kept for tools/microbench_diag.cu)  One warp.  D: 8 x 8 row-major diagonal block (lower triangle valid).  On exit D holds inv(L), L the Cholesky factor
// of the block, with an explicit zero upper triangle.  Lane 0 factors in registers; lanes 0..7 then take one column of
// the inverse each.  Returns (on lane 0) whether a non-positive pivot was met.
__device__ __forceinline__ bool factor_invert_diag_v0(double* D, int lane) {
    bool bad_any = false;
    if (lane == 0) {
        double a[8][8];
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int c = 0; c <= r; ++c) a[r][c] = D[r * 8 + c];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const double d = a[c][c];
            const bool bad = !(d > 0.0);
            bad_any |= bad;
            const double inv = bad ? 1.0 : fast_rsqrt(d);
            a[c][c] = inv;   // the diagonal keeps 1 / L_cc
#pragma unroll
            for (int r = c + 1; r < 8; ++r) a[r][c] *= inv;
#pragma unroll
            for (int c2 = c + 1; c2 < 8; ++c2)
#pragma unroll
                for (int r = c2; r < 8; ++r) a[r][c2] -= a[r][c] * a[c2][c];
        }
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int c = 0; c <= r; ++c) D[r * 8 + c] = a[r][c];
    }
    __syncwarp();
    double x[8];
    if (lane < 8) {   // column `lane` of X = inv(L):  X[r][c] = -(1 / L_rr) sum_{k < r} L[r][k] X[k][c],  X[c][c] = 1 / L_cc
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < r; ++k) s += D[r * 8 + k] * x[k];
            const double ir = D[r * 8 + r];
            x[r] = (r == lane) ? ir : ((r < lane) ? 0.0 : -ir * s);
        }
    }
    __syncwarp();
    if (lane < 8) {
#pragma unroll
        for (int r = 0; r < 8; ++r) D[r * 8 + lane] = x[r];
    }
    return bad_any;
}

// Whole CTA.  K: block-packed lower triangle with nb block rows (rows beyond the matrix padded with the identity).
// On exit: blocks outside the 64 x 64 diagonal super-blocks hold L, the super-blocks hold the inverse of L's.
// *flag (shared) is set to 1 when a pivot was not positive.  Ends with a barrier.
static __device__ __noinline__ void factor(double* K, int nb, int* flag) {
    __builtin_assume(__isShared(K));
    __builtin_assume(__isShared(flag));
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nwarp = blockDim.x >> 5;
    if (tid == 0) *flag = 0;
    __syncthreads();
    // Column j, between two CTA-wide barriers:
    //   warp 0      : L(j+1, j) = C(j+1, j) inv(L_jj)', C(j+1, j+1) -= L(j+1, j) L(j+1, j)', then the serial step -- factor and
    //                 invert the diagonal block (j+1, j+1) -- without waiting for anybody;
    //   warps 1 ..  : the rows below, L(i, j) = C(i, j) inv(L_jj)' and column j applied to column j+1 (each recomputes
    //                 L(j+1, j) instead of waiting for warp 0), a barrier among themselves, then the look-ahead: columns
    //                 <= j applied to column j+2.
    // The chain of 8 x 8 diagonal factorisations is the critical path; nothing but one small block row sits between two of them.
    double2 keep = make_double2(0.0, 0.0);   // warp 0: L(j+1, j), stored one column late (the others still read C(j+1, j))
    CPROF_DECL
    if (wid == 0 && factor_invert_diag(K + blk(0, 0), lane)) *flag = 1;
    __syncthreads();
    const int nother = (nwarp - 1) * 32;
    for (int j = 0; j + 1 < nb; ++j) {
        const double2 X = ldfrag(K + blk(j, j), lane);
        double2 Lj1 = make_double2(0.0, 0.0);
        mma_pqT(Lj1, ldfrag(K + blk(j + 1, j), lane), X);
        if (wid == 0) {
            if (j > 0) stfrag(K + blk(j, j - 1), lane, keep);
            keep = Lj1;
            double2 upd = make_double2(0.0, 0.0);
            mma_pqT(upd, Lj1, Lj1);
            double2 c = ldfrag(K + blk(j + 1, j + 1), lane);
            c.x -= upd.x;
            c.y -= upd.y;
            stfrag(K + blk(j + 1, j + 1), lane, c);
            __syncwarp();
            if (factor_invert_diag(K + blk(j + 1, j + 1), lane)) *flag = 1;
            CPROF(0);
        } else {
            for (int i = j + 2 + (wid - 1); i < nb; i += nwarp - 1) {
                double2 Lij = make_double2(0.0, 0.0);
                mma_pqT(Lij, ldfrag(K + blk(i, j), lane), X);
                stfrag(K + blk(i, j), lane, Lij);
                double2 upd = make_double2(0.0, 0.0);
                mma_pqT(upd, Lij, Lj1);
                double2 c = ldfrag(K + blk(i, j + 1), lane);
                c.x -= upd.x;
                c.y -= upd.y;
                stfrag(K + blk(i, j + 1), lane, c);
            }
            asm volatile("bar.sync 1, %0;" ::"r"(nother) : "memory");   // column j complete among warps 1 ..
            const int jc = j + 2;
            for (int i = jc + (wid - 1); i < nb; i += nwarp - 1) {   // look-ahead: columns 0 .. j applied to column j + 2
                double2 a0 = make_double2(0.0, 0.0), a1 = make_double2(0.0, 0.0);
                const double* Li = K + blk(i, 0);
                const double* Lj = K + blk(jc, 0);
                int k = 0;
                for (; k + 1 <= j; k += 2) {
                    const double2 p0 = ldfrag(Li + (k << 6), lane), q0 = ldfrag(Lj + (k << 6), lane);
                    const double2 p1 = ldfrag(Li + ((k + 1) << 6), lane), q1 = ldfrag(Lj + ((k + 1) << 6), lane);
                    mma_pqT(a0, p0, q0);
                    mma_pqT(a1, p1, q1);
                }
                if (k <= j) mma_pqT(a0, ldfrag(Li + (k << 6), lane), ldfrag(Lj + (k << 6), lane));
                double2 c = ldfrag(Li + (jc << 6), lane);
                c.x -= a0.x + a1.x;
                c.y -= a0.y + a1.y;
                stfrag(K + blk(i, jc), lane, c);
            }
        }
        __syncthreads();
        CPROF(1);
    }
    if (wid == 0 && nb > 1) stfrag(K + blk(nb - 1, nb - 2), lane, keep);
    __syncthreads();
    CPROF(4);
    // ---- inverses of the 16 x 16, 32 x 32, 64 x 64 diagonal super-blocks, in place over L's blocks inside them.
    // Level h (half size in blocks): group q covers blocks [2hq, 2hq + 2h); with X11, X22 the inverses of its two halves
    //   X21 = -X22 (L21 X11)
    // in two sweeps that need no temporary: (A) T = L21 X11, one warp per block ROW of L21, columns ascending -- T(r, c)
    // reads L(r, k >= c) only, so overwriting L(r, c) with it is safe; (B) X21 = -X22 T, one warp per block COLUMN, rows
    // descending -- X21(r, c) reads T(k <= r, c) only.  One barrier after each sweep.
    for (int h = 1; h <= 4; h <<= 1) {
        const int ngroup = (nb + 2 * h - 1) / (2 * h);
        for (int it = wid; it < ngroup * h; it += nwarp) {   // sweep A: `it` = (group, row inside the lower half)
            const int q = it / h, b0 = 2 * h * q, r = b0 + h + it % h;
            if (r >= nb) continue;
            for (int c = b0; c < b0 + h; ++c) {
                double2 T = make_double2(0.0, 0.0);
                for (int k = c; k < b0 + h; ++k) mma_pqT(T, ldfrag(K + blk(r, k), lane), ldfragT(K + blk(k, c), lane));
                __syncwarp();
                stfrag(K + blk(r, c), lane, T);
            }
        }
        __syncthreads();
        for (int it = wid; it < ngroup * h; it += nwarp) {   // sweep B: `it` = (group, column inside the left half)
            const int q = it / h, b0 = 2 * h * q, c = b0 + it % h;
            const int rtop = (b0 + 2 * h < nb) ? b0 + 2 * h : nb;
            for (int r = rtop - 1; r >= b0 + h; --r) {
                double2 Xn = make_double2(0.0, 0.0);
                for (int k = b0 + h; k <= r; ++k) mma_pqT(Xn, ldfrag(K + blk(r, k), lane), ldfragT(K + blk(k, c), lane));
                __syncwarp();
                stfrag(K + blk(r, c), lane, make_double2(-Xn.x, -Xn.y));
                __syncwarp();
            }
        }
        __syncthreads();
    }
    CPROF(5);
}

// Whole CTA, blockDim.x >= 256.  Solves (L L') x = v in place; v has 8 nb entries (padding zero, 16-byte aligned), ys
// is 64 doubles of scratch.  With the 64 x 64 diagonal super-blocks inverted a triangular solve is ceil(nb / 8)
// dependent steps of two phases: (a) y_s = X_ss v_s, one warp per block row, the up to eight block fragments and
// vector pieces loaded before the first FMA; (b) the rows below take L(i, s) y_s.  Starts and ends with a barrier.
static __device__ __noinline__ void solve(const double* K, int nb, double* v, double* ys) {
    __builtin_assume(__isShared(K));
    __builtin_assume(__isShared(v));
    __builtin_assume(__isShared(ys));
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nwarp = blockDim.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int nsup = (nb + 7) >> 3;
    // sum over the blocks q in [qlo, qhi) of a block row:  P(q) x(q)  reduced to one value per row g (all lanes of the
    // row group get it); Kr: first block of the range's row, x: vector piece of block q at x + 8 q
    auto row_dot = [&](const double* Kr, const double* x, int qlo, int qhi) -> double {
        double a0 = 0.0, a1 = 0.0;
#pragma unroll 1
        for (int q0 = qlo; q0 < qhi; q0 += 4) {
            double2 p[4], xx[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (q0 + u < qhi) {
                    p[u] = ldfrag(Kr + ((q0 + u) << 6), lane);
                    xx[u] = *reinterpret_cast<const double2*>(x + 8 * (q0 + u) + 2 * t);
                }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (q0 + u < qhi) {
                    a0 = fma(p[u].x, xx[u].x, a0);
                    a1 = fma(p[u].y, xx[u].y, a1);
                }
        }
        a0 += a1;
        a0 += __shfl_xor_sync(0xffffffffu, a0, 1);
        a0 += __shfl_xor_sync(0xffffffffu, a0, 2);
        return a0;
    };
    // sum over the block rows q in [qlo, qhi) of a block column:  P(q)' x(q)  reduced to the columns (2t, 2t+1) (valid
    // on the lanes g == 0); Kc: block (row qlo.., this column), consecutive rows `stride(q)` apart -- passed as a lambda
    auto col_dot = [&](int rlo, int rhi, int c, const double* x, int xoff) -> double2 {
        double ax = 0.0, ay = 0.0;
#pragma unroll 1
        for (int r0 = rlo; r0 < rhi; r0 += 4) {
            double2 p[4];
            double xv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (r0 + u < rhi) {
                    p[u] = ldfrag(K + blk(r0 + u, c), lane);
                    xv[u] = x[8 * (r0 + u - xoff) + g];
                }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (r0 + u < rhi) {
                    ax = fma(p[u].x, xv[u], ax);
                    ay = fma(p[u].y, xv[u], ay);
                }
        }
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
            ax += __shfl_xor_sync(0xffffffffu, ax, o);
            ay += __shfl_xor_sync(0xffffffffu, ay, o);
        }
        return make_double2(ax, ay);
    };
    __syncthreads();
    for (int s = 0; s < nsup; ++s) {   // L y = v
        const int b0 = 8 * s, b1 = (b0 + 8 < nb) ? b0 + 8 : nb, cnt = b1 - b0;
        if (wid < cnt) {
            const double a = row_dot(K + blk(b0 + wid, 0), v, b0, b0 + wid + 1);
            if (t == 0) ys[8 * wid + g] = a;
        }
        __syncthreads();
        for (int i = b1 + wid; i < nb; i += nwarp) {
            const double a = row_dot(K + blk(i, 0), ys - 8 * b0, b0, b1);
            if (t == 0) v[8 * i + g] -= a;
        }
        if (tid < 8 * cnt) v[8 * b0 + tid] = ys[tid];
        __syncthreads();
    }
    for (int s = nsup - 1; s >= 0; --s) {   // L' x = y
        const int b0 = 8 * s, b1 = (b0 + 8 < nb) ? b0 + 8 : nb, cnt = b1 - b0;
        if (wid < cnt) {
            const double2 a = col_dot(b0 + wid, b1, b0 + wid, v, 0);
            if (g == 0) *reinterpret_cast<double2*>(ys + 8 * wid + 2 * t) = a;
        }
        __syncthreads();
        for (int kb = wid; kb < b0; kb += nwarp) {
            const double2 a = col_dot(b0, b1, kb, ys, b0);
            if (g == 0) {
                double2* dst = reinterpret_cast<double2*>(v + 8 * kb + 2 * t);
                double2 cur = *dst;
                cur.x -= a.x;
                cur.y -= a.y;
                *dst = cur;
            }
        }
        if (tid < 8 * cnt) v[8 * b0 + tid] = ys[tid];
        __syncthreads();
    }
}

}  // namespace chol
}  // namespace bgg
