// Latency and accuracy of the serial step of chol::factor -- factor an 8 x 8 SPD block and invert its factor -- in
// the variants tried (results: profiles/r01d_microbench.txt).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I bilevel-gait-gen_b200/csrc -o tools/bin/microbench_diag tools/microbench_diag.cu
#include <cmath>
#include <cstdio>
#include <cuda_runtime.h>

#include "bgg_chol.cuh"

using namespace bgg::chol;

__device__ __forceinline__ double rsqrt_seed64(double d) {
    double x;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
    return x;
}
__global__ void k_seed_acc(double* out) {
    // relative error of the seed and of the refined value over a sweep of mantissas / exponents
    double worst_seed = 0, worst_ref = 0;
    for (int i = threadIdx.x; i < 1 << 16; i += blockDim.x) {
        const double d = (1.0 + i / 65536.0 * 3.0) * exp2(static_cast<double>((i % 61) - 30));
        const double ex = 1.0 / sqrt(d);
        const double s = rsqrt_seed64(d);
        worst_seed = fmax(worst_seed, fabs(s - ex) / ex);
        const double r = rsqrt_fast(d);
        worst_ref = fmax(worst_ref, fabs(r - ex) / ex);
    }
    for (int o = 16; o > 0; o >>= 1) {
        worst_seed = fmax(worst_seed, __shfl_xor_sync(0xffffffffu, worst_seed, o));
        worst_ref = fmax(worst_ref, __shfl_xor_sync(0xffffffffu, worst_ref, o));
    }
    if (threadIdx.x == 0) {
        out[0] = worst_seed;
        out[1] = worst_ref;
    }
}

template <int V>
__global__ void k_diag(const double* A, double* X, long long* cyc, int n) {
    __shared__ __align__(16) double D[64];
    const int lane = threadIdx.x;
    double a0 = A[lane], a1 = A[32 + lane];
    long long total = 0;
    for (int it = 0; it < n; ++it) {
        D[lane] = a0;
        D[32 + lane] = a1;
        __syncwarp();
        const long long t0 = clock64();
        if (V == 0) factor_invert_diag_v0(D, lane);
        else factor_invert_diag(D, lane);
        __syncwarp();
        total += clock64() - t0;
    }
    X[lane] = D[lane];
    X[32 + lane] = D[32 + lane];
    if (lane == 0) cyc[0] = total;
}

int main() {
    double hA[64], hX[64];
    // SPD block with the spread of the KKT diagonal blocks: entries 1e-3 .. 1e8
    double G[8][8];
    unsigned s = 12345;
    for (int i = 0; i < 8; ++i)
        for (int j = 0; j < 8; ++j) {
            s = s * 1664525u + 1013904223u;
            G[i][j] = ((s >> 8) % 2001) / 1000.0 - 1.0;
        }
    const double sc[8] = {1e4, 3e3, 1e2, 10, 1, 0.3, 0.1, 0.03};
    for (int i = 0; i < 8; ++i)
        for (int j = 0; j < 8; ++j) {
            double v = (i == j) ? 1e-3 : 0.0;
            for (int k = 0; k < 8; ++k) v += G[i][k] * sc[k] * G[j][k] * sc[k];
            hA[i * 8 + j] = v;
        }
    double *dA, *dX, *dO;
    long long* dc;
    cudaMalloc(&dA, 512); cudaMalloc(&dX, 512); cudaMalloc(&dO, 64); cudaMalloc(&dc, 8);
    cudaMemcpy(dA, hA, 512, cudaMemcpyHostToDevice);
    k_seed_acc<<<1, 32>>>(dO);
    double ho[2];
    cudaMemcpy(ho, dO, 16, cudaMemcpyDeviceToHost);
    printf("rsqrt.approx.ftz.f64 seed: worst relative error %.3e ; refined (rsqrt_fast): %.3e\n", ho[0], ho[1]);
    const int n = 200;
    for (int v = 0; v < 2; ++v) {
        if (v == 0) { k_diag<0><<<1, 32>>>(dA, dX, dc, n); k_diag<0><<<1, 32>>>(dA, dX, dc, n); }
        else { k_diag<1><<<1, 32>>>(dA, dX, dc, n); k_diag<1><<<1, 32>>>(dA, dX, dc, n); }
        long long c;
        cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
        cudaMemcpy(hX, dX, 512, cudaMemcpyDeviceToHost);
        // residual max |X A X' - I| and the explicit zeros above the diagonal
        double worst = 0, upper = 0;
        for (int i = 0; i < 8; ++i)
            for (int j = 0; j < 8; ++j) {
                double v2 = 0;
                for (int k = 0; k < 8; ++k)
                    for (int l = 0; l < 8; ++l) v2 += hX[i * 8 + k] * hA[k * 8 + l] * hX[j * 8 + l];
                worst = fmax(worst, fabs(v2 - (i == j ? 1.0 : 0.0)));
                if (j > i) upper = fmax(upper, fabs(hX[i * 8 + j]));
            }
        printf("variant %d: %.0f cycles per block ; max |X A X' - I| = %.3e ; max |upper| = %.1e\n", v, static_cast<double>(c) / n, worst, upper);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
