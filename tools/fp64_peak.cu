// Micro-benchmark: FP64 DFMA / DADD / DMUL throughput and 64-bit shared-memory gather throughput.
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void k(double* out, int iters, double a, double b) {
    double x[8];
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) x[i] = fma(x[i], a, b);
            if (OP == 1) x[i] = x[i] + a;
            if (OP == 2) x[i] = x[i] * a;
        }
    }
    double s = 0;
    for (int i = 0; i < 8; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void gather(double* out, int iters, int stride) {
    __shared__ double sm[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i;
    __syncthreads();
    unsigned idx = threadIdx.x * 2654435761u;
    double s = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            idx = idx * 1664525u + 1013904223u;
            s += sm[(idx >> 8) & 4095];
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    double* d;
    cudaMalloc(&d, 148 * 8 * 1024 * sizeof(double));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 20000, blocks = 148 * 4, threads = 512;
    const char* names[3] = {"DFMA", "DADD", "DMUL"};
    for (int op = 0; op < 3; ++op) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (op == 0) k<0><<<blocks, threads>>>(d, iters, 1.0000001, 1e-9);
            if (op == 1) k<1><<<blocks, threads>>>(d, iters, 1.0000001, 1e-9);
            if (op == 2) k<2><<<blocks, threads>>>(d, iters, 1.0000001, 1e-9);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
        }
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        double ops = (double)blocks * threads * iters * 8;
        printf("%s: %.2f Tinstr/s  (%.2f TFLOP/s as FMA=2)  %.3f ms\n", names[op], ops / ms / 1e9,
               ops * (op == 0 ? 2 : 1) / ms / 1e9, ms);
    }
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        gather<<<148 * 4, 512>>>(d, 4000, 1);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
    }
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double g = (double)148 * 4 * 512 * 4000 * 8;
    printf("random 64-bit smem gathers: %.2f Tgather/s = %.1f per cycle per SM (at 1.9 GHz)\n", g / ms / 1e9,
           g / ms / 1e6 / 148 / 1.9e3 * 1e-3 * 1e3);
    return 0;
}
