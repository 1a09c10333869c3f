// dev probe: D2H bandwidth of 1D vs strided 2D copies (pinned host memory) and host NT fill
#include <cstdio>
#include <cstring>
#include <chrono>
#include <thread>
#include <vector>
#include <cuda_runtime.h>
#include <emmintrin.h>
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main()
{
    const size_t nb = 4096, nnz = 19970, head = 13824, tail = nnz - head;
    double *h, *d;
    cudaMallocHost(&h, nb * nnz * 8);
    cudaMalloc(&d, nb * nnz * 8);
    cudaMemset(d, 0, nb * nnz * 8);
    memset(h, 0, nb * nnz * 8);
    cudaStream_t st; cudaStreamCreate(&st);
    for (int rep = 0; rep < 2; ++rep) {
        double t0 = now();
        cudaMemcpyAsync(h, d, nb * nnz * 8, cudaMemcpyDeviceToHost, st); cudaStreamSynchronize(st);
        double t1 = now();
        printf("1D full   %.2f ms  %.1f GB/s\n", (t1 - t0) * 1e3, nb * nnz * 8 / (t1 - t0) / 1e9);
        t0 = now();
        cudaMemcpyAsync(h, d, nb * head * 8, cudaMemcpyDeviceToHost, st); cudaStreamSynchronize(st);
        t1 = now();
        printf("1D head-sized %.2f ms  %.1f GB/s\n", (t1 - t0) * 1e3, nb * head * 8 / (t1 - t0) / 1e9);
        t0 = now();
        cudaMemcpy2DAsync(h, nnz * 8, d, nnz * 8, head * 8, nb, cudaMemcpyDeviceToHost, st); cudaStreamSynchronize(st);
        t1 = now();
        printf("2D head   %.2f ms  %.1f GB/s\n", (t1 - t0) * 1e3, nb * head * 8 / (t1 - t0) / 1e9);
        t0 = now();
        cudaMemcpy2DAsync(h, nnz * 8, d, head * 8, head * 8, nb, cudaMemcpyDeviceToHost, st); cudaStreamSynchronize(st);
        t1 = now();
        printf("2D head (compact src) %.2f ms  %.1f GB/s\n", (t1 - t0) * 1e3, nb * head * 8 / (t1 - t0) / 1e9);
        for (int nt : {4, 8, 16}) {
            std::vector<double> src(tail, 1.0);
            t0 = now();
            std::vector<std::thread> th;
            for (int t = 0; t < nt; ++t) th.emplace_back([=, &src]() {
                for (size_t b = t; b < nb; b += nt) {
                    double* dst = h + b * nnz + head;
                    size_t i = 0;
                    if ((uintptr_t)dst & 15u) { _mm_stream_si64((long long*)dst, *(const long long*)src.data()); i = 1; }
                    for (; i + 2 <= tail; i += 2) _mm_stream_pd(dst + i, _mm_loadu_pd(src.data() + i));
                    if (i < tail) _mm_stream_si64((long long*)(dst + i), *(const long long*)(src.data() + i));
                }
                _mm_sfence();
            });
            for (auto& t : th) t.join();
            t1 = now();
            printf("host NT fill %d threads %.2f ms  %.1f GB/s\n", nt, (t1 - t0) * 1e3, nb * tail * 8 / (t1 - t0) / 1e9);
            t0 = now();
            th.clear();
            for (int t = 0; t < nt; ++t) th.emplace_back([=, &src]() { for (size_t b = t; b < nb; b += nt) memcpy(h + b * nnz + head, src.data(), tail * 8); });
            for (auto& t : th) t.join();
            t1 = now();
            printf("host memcpy fill %d threads %.2f ms  %.1f GB/s\n", nt, (t1 - t0) * 1e3, nb * tail * 8 / (t1 - t0) / 1e9);
        }
    }
    return 0;
}
