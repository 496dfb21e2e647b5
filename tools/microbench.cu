// FP64 micro-benchmarks that size the k_ipm design on B200 (results: profiles/r01d_microbench.txt):
//   dependent DFMA latency, DFMA throughput per SM, mma.sync.m8n8k4.f64 (DMMA) latency / throughput per SM,
//   shared-memory load latency (dependent LDS.64), __syncthreads cost at 256 threads, MUFU.RSQ + cvt chain.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/microbench tools/microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dfma_lat(double* out, long long* cyc, int n) {
    double a = out[0], b = out[1], c = out[2];
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) a = fma(a, b, c);
    }
    long long t1 = clock64();
    out[3] = a;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_dfma_tput(double* out, long long* cyc, int n) {
    double a[8];
    const double b = out[1], c = out[2];
    for (int j = 0; j < 8; ++j) a[j] = out[0] + j + threadIdx.x;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = fma(a[j], b, c);
    }
    __syncthreads();
    long long t1 = clock64();
    double s = 0;
    for (int j = 0; j < 8; ++j) s += a[j];
    out[4 + threadIdx.x % 4] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void k_dmma_lat(double* out, long long* cyc, int n) {
    double c0 = out[0], c1 = out[1];
    const double a = out[2], b = out[3];
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) dmma(c0, c1, a, b);
    }
    long long t1 = clock64();
    out[4] = c0 + c1;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_dmma_tput(double* out, long long* cyc, int n) {
    double c[8][2];
    const double a = out[2], b = out[3];
    for (int j = 0; j < 8; ++j) c[j][0] = c[j][1] = out[0] + j;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) dmma(c[j][0], c[j][1], a, b);
    }
    __syncthreads();
    long long t1 = clock64();
    double s = 0;
    for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1];
    out[4 + threadIdx.x % 4] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_lds_lat(double* out, long long* cyc, int n) {
    __shared__ double sh[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sh[i] = static_cast<double>((i * 7 + 1) & 1023);
    __syncthreads();
    int idx = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) idx = static_cast<int>(sh[idx & 1023]);
    }
    long long t1 = clock64();
    out[4] = idx;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_bar(double* out, long long* cyc, int n) {
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) __syncthreads();
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
    out[5] = 0;
}
__global__ void k_rsqrt_chain(double* out, long long* cyc, int n) {
    double d = out[0] + 3.0;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            double x = static_cast<double>(rsqrtf(static_cast<float>(d)));
            const double h = 0.5 * d;
            x = x * (1.5 - h * x * x);
            x = x * (1.5 - h * x * x);
            d = d * x + 2.0;
        }
    }
    long long t1 = clock64();
    out[4] = d;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_ddiv_chain(double* out, long long* cyc, int n) {
    double d = out[0] + 3.0;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) d = 1.0 / d + 2.0;
    }
    long long t1 = clock64();
    out[4] = d;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_shfl_chain(double* out, long long* cyc, int n) {
    double d = out[0] + threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) d += __shfl_xor_sync(0xffffffffu, d, 1 << (j & 3));
    }
    long long t1 = clock64();
    out[4] = d;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
    double* out;
    long long* cyc;
    cudaMalloc(&out, 64 * sizeof(double));
    cudaMalloc(&cyc, 8 * sizeof(long long));
    double h[8] = {1.0000001, 0.9999999, 1e-9, 1e-9, 0, 0, 0, 0};
    cudaMemcpy(out, h, sizeof(h), cudaMemcpyHostToDevice);
    long long c;
    const int n = 2000;
    auto rd = [&]() { cudaDeviceSynchronize(); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost); return static_cast<double>(c); };
    k_dfma_lat<<<1, 32>>>(out, cyc, n); k_dfma_lat<<<1, 32>>>(out, cyc, n);
    printf("DFMA dependent latency           %.2f cycles\n", rd() / (n * 16.0));
    for (int th : {128, 256, 512, 1024}) {
        k_dfma_tput<<<1, th>>>(out, cyc, n); k_dfma_tput<<<1, th>>>(out, cyc, n);
        printf("DFMA throughput, 1 CTA x %4d thr  %.1f FMA/clk/SM\n", th, n * 8.0 * th / rd());
    }
    k_dmma_lat<<<1, 32>>>(out, cyc, n); k_dmma_lat<<<1, 32>>>(out, cyc, n);
    printf("DMMA m8n8k4 dependent latency    %.2f cycles\n", rd() / (n * 16.0));
    for (int th : {128, 256, 512, 1024}) {
        k_dmma_tput<<<1, th>>>(out, cyc, n); k_dmma_tput<<<1, th>>>(out, cyc, n);
        printf("DMMA throughput, 1 CTA x %4d thr  %.1f FMA/clk/SM\n", th, n * 8.0 * (th / 32) * 256.0 / rd());
    }
    k_lds_lat<<<1, 32>>>(out, cyc, n); k_lds_lat<<<1, 32>>>(out, cyc, n);
    printf("LDS.64 + cvt dependent chain     %.2f cycles\n", rd() / (n * 16.0));
    for (int th : {32, 256}) {
        k_bar<<<1, th>>>(out, cyc, n); k_bar<<<1, th>>>(out, cyc, n);
        printf("__syncthreads, %3d threads        %.2f cycles\n", th, rd() / (n * 16.0));
    }
    k_rsqrt_chain<<<1, 32>>>(out, cyc, n); k_rsqrt_chain<<<1, 32>>>(out, cyc, n);
    printf("rsqrtf seed + 2 Newton + fma     %.2f cycles per pivot\n", rd() / (n * 8.0));
    k_ddiv_chain<<<1, 32>>>(out, cyc, n); k_ddiv_chain<<<1, 32>>>(out, cyc, n);
    printf("1.0 / d + c chain                %.2f cycles\n", rd() / (n * 8.0));
    k_shfl_chain<<<1, 32>>>(out, cyc, n); k_shfl_chain<<<1, 32>>>(out, cyc, n);
    printf("shfl.f64 + add chain             %.2f cycles\n", rd() / (n * 8.0));
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
