// Launch-throughput microbenchmark: T host threads, one stream each, N launches of a tiny kernel with
// a 300-byte by-value argument (like SigmaArgs).  Prints microseconds per launch (host side).
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <thread>
#include <vector>
#include <cuda_runtime.h>
struct Big { char b[320]; };
__global__ void tiny(Big a, int* out) { if (a.b[0] == 77) out[0] = 1; }
int main(int argc, char** argv) {
    const int N = 20000;
    int* d; cudaMalloc(&d, 4);
    for (int T : {1, 2, 4, 8}) {
        std::vector<cudaStream_t> st(T);
        for (auto& s : st) cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
        auto t0 = std::chrono::steady_clock::now();
        std::vector<std::thread> th;
        for (int t = 0; t < T; ++t)
            th.emplace_back([&, t] { Big a{}; for (int i = 0; i < N; ++i) tiny<<<1, 32, 0, st[t]>>>(a, d); cudaStreamSynchronize(st[t]); });
        for (auto& x : th) x.join();
        double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
        printf("threads %d: %.2f us per launch per thread, %.2f us per launch overall (%.0f k launches/s)\n", T, us / N, us / N / T, 1e3 * N * T / us);
    }
    return 0;
}
