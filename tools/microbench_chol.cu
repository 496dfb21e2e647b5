// Stand-alone timing and accuracy check of chol::factor / chol::solve (csrc/bgg_chol.cuh) on a 120 x 120 SPD matrix
// with the spread of the interior-point KKT matrix; 256 threads per CTA, 1 CTA alone or 2 CTAs on every SM.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I bilevel-gait-gen_b200/csrc -o tools/bin/microbench_chol tools/microbench_chol.cu
#include <cmath>
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>

#include "bgg_chol.cuh"

using namespace bgg::chol;

__global__ void __launch_bounds__(256, 2) k_chol(const double* Ab, const double* rhs, double* xout, long long* cyc, int nb, int reps) {
    extern __shared__ __align__(16) double sm[];
    double* K = sm;
    double* v = K + doubles(nb);
    double* ys = v + 8 * nb;
    __shared__ int flag;
    const int tid = threadIdx.x;
    long long tf = 0, ts = 0;
    for (int it = 0; it < reps; ++it) {
        for (int i = tid; i < static_cast<int>(doubles(nb)); i += blockDim.x) K[i] = Ab[i];
        for (int i = tid; i < 8 * nb; i += blockDim.x) v[i] = rhs[i];
        __syncthreads();
        const long long t0 = clock64();
        factor(K, nb, &flag);
        const long long t1 = clock64();
        solve(K, nb, v, ys);
        const long long t2 = clock64();
        tf += t1 - t0;
        ts += t2 - t1;
    }
    if (blockIdx.x == 0) {
        for (int i = tid; i < 8 * nb; i += blockDim.x) xout[i] = v[i];
        if (tid == 0) {
            cyc[0] = tf / reps;
            cyc[1] = ts / reps;
            cyc[2] = flag;
        }
    }
}

int main() {
    const int n = 120, nb = n / 8;
    std::vector<double> A(n * n), G(n * 40), b(n), Ab(doubles(nb), 0.0);
    unsigned s = 777;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((s >> 8) % 20001) / 10000.0 - 1.0; };
    for (auto& g : G) g = rnd();
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            double v = (i == j) ? 1e-3 * (1 + (i % 7)) : 0.0;
            for (int k = 0; k < 40; ++k) v += G[i * 40 + k] * G[j * 40 + k] * std::pow(10.0, (k % 9) - 2.0);
            A[i * n + j] = v;
        }
    for (auto& x : b) x = rnd();
    for (int i = 0; i < n; ++i)
        for (int j = 0; j <= i; ++j) {
            const int bi = i >> 3, bj = j >> 3;
            Ab[(static_cast<size_t>(bi) * (bi + 1) / 2 + bj) * 64 + (i & 7) * 8 + (j & 7)] = A[i * n + j];
        }
    double *dA, *db, *dx;
    long long* dc;
    cudaMalloc(&dA, Ab.size() * 8); cudaMalloc(&db, n * 8); cudaMalloc(&dx, n * 8); cudaMalloc(&dc, 64);
    cudaMemcpy(dA, Ab.data(), Ab.size() * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(db, b.data(), n * 8, cudaMemcpyHostToDevice);
    const size_t smem = (doubles(nb) + 8 * nb + 64) * 8 + 50 * 1024;   // + ballast: the footprint of k_ipm (2 CTAs / SM)
    cudaFuncSetAttribute(k_chol, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    for (int grid : {1, 296}) {
        k_chol<<<grid, 256, smem>>>(dA, db, dx, dc, nb, 20);
        k_chol<<<grid, 256, smem>>>(dA, db, dx, dc, nb, 20);
        long long c[3];
        std::vector<double> x(n);
        cudaMemcpy(c, dc, 24, cudaMemcpyDeviceToHost);
        cudaMemcpy(x.data(), dx, n * 8, cudaMemcpyDeviceToHost);
        double rn = 0, bn = 0;
        for (int i = 0; i < n; ++i) {
            double r = -b[i];
            for (int j = 0; j < n; ++j) r += A[i * n + j] * x[j];
            rn = fmax(rn, fabs(r));
            bn = fmax(bn, fabs(b[i]));
        }
        printf("grid %3d: factor %lld cycles, solve %lld cycles, flag %lld, |Ax-b|/|b| = %.3e  (%s)\n", grid, c[0], c[1], c[2], rn / bn,
               cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
