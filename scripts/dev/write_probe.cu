// write_probe.cu -- what HBM write bandwidth does the value-scatter pattern of k_cons_jac allow?
//   A: contiguous streaming stores (one 8-byte store per thread per iteration, grid-stride)
//   B: the kernel's pattern: thread = (instance b, node k); NBLK*NROW stores at vb + blk*N, vb = b*nnz + k
//      (256-byte runs per warp, 16-byte misaligned per instance as nnz_jac*8 % 128 = 16)
//   C: as B with the instance pitch padded to a multiple of 128 bytes (what an aligned layout would give)
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/write_probe scripts/dev/write_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_contig(double* out, size_t n, double v)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) __stcs(out + i, v);
}

template <int NBLOCKS>
__global__ void __launch_bounds__(128) k_pattern(double* out, int nb, int N, size_t pitch, double v)
{
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)nb * N) return;
    const int b = (int)(gid / N), k = (int)(gid - (long long)b * N);
    double* vb = out + (size_t)b * pitch + k;
#pragma unroll 8
    for (int blk = 0; blk < NBLOCKS; ++blk) __stcs(vb + (unsigned)blk * (unsigned)N, v + blk);
}

// D: as B, with the node index rotated per instance so that warp 0's 256-byte runs start on a 128-byte line
template <int NBLOCKS>
__global__ void __launch_bounds__(128) k_rotated(double* out, int nb, int N, size_t pitch, double v)
{
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)nb * N) return;
    const int b = (int)(gid / N);
    int k = (int)(gid - (long long)b * N);
    const size_t base = (size_t)b * pitch;
    const unsigned delta = (unsigned)((16 - ((((size_t)out >> 3) + base) & 15)) & 15);
    k += delta;
    if (k >= N) k -= N;
    double* vb = out + base + k;
#pragma unroll 8
    for (int blk = 0; blk < NBLOCKS; ++blk) __stcs(vb + (unsigned)blk * (unsigned)N, v + blk);
}

// E/F: the kernel's real store ORDER: colour-major (for cc: for row i: block i*NB + cc), unrotated / rotated
template <int NROW, int NB, bool ROT>
__global__ void __launch_bounds__(128) k_colour_major(double* out, int nb, int N, size_t pitch, double v)
{
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)nb * N) return;
    const int b = (int)(gid / N);
    int k = (int)(gid - (long long)b * N);
    const size_t base = (size_t)b * pitch;
    if (ROT) {
        k += (int)((16 - ((((size_t)out >> 3) + base) & 15)) & 15);
        if (k >= N) k -= N;
    }
    double* vb = out + base + k;
#pragma unroll 1
    for (int cc = 0; cc < NB; ++cc)
#pragma unroll
        for (int i = 0; i < NROW; ++i) __stcs(vb + (unsigned)(i * NB + cc) * (unsigned)N, v + cc);
}

// G: colour chunks of CH colours, row-major inside a chunk (what staging CH colours would allow)
template <int NROW, int NB, int CH, bool ROT>
__global__ void __launch_bounds__(128) k_chunked(double* out, int nb, int N, size_t pitch, double v)
{
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)nb * N) return;
    const int b = (int)(gid / N);
    int k = (int)(gid - (long long)b * N);
    const size_t base = (size_t)b * pitch;
    if (ROT) {
        k += (int)((16 - ((((size_t)out >> 3) + base) & 15)) & 15);
        if (k >= N) k -= N;
    }
    double* vb = out + base + k;
#pragma unroll 1
    for (int c0 = 0; c0 < NB; c0 += CH)
#pragma unroll
        for (int i = 0; i < NROW; ++i)
#pragma unroll
            for (int cc = c0; cc < c0 + CH && cc < NB; ++cc) __stcs(vb + (unsigned)(i * NB + cc) * (unsigned)N, v + cc);
}

template <class F>
float time_ms(F f, int reps)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) f();
    cudaEventRecord(a);
    for (int i = 0; i < reps; ++i) f();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    return ms / reps;
}

int main()
{
    const int nb = 4096, N = 64, NBLK = 216 + 12; // Jacobian blocks + the 12 g rows
    const size_t nnz = 19970, padded = (nnz + 15) / 16 * 16;
    double* out;
    cudaMalloc(&out, (size_t)nb * padded * 8 * 2);
    const size_t bytes = (size_t)nb * N * NBLK * 8;
    float a = time_ms([&] { k_contig<<<148 * 16, 256>>>(out, bytes / 8, 1.0); }, 20);
    float b = time_ms([&] { k_pattern<NBLK><<<(nb * N + 127) / 128, 128>>>(out, nb, N, nnz, 1.0); }, 20);
    float c = time_ms([&] { k_pattern<NBLK><<<(nb * N + 127) / 128, 128>>>(out, nb, N, padded, 1.0); }, 20);
    float d = time_ms([&] { k_rotated<NBLK><<<(nb * N + 127) / 128, 128>>>(out, nb, N, nnz, 1.0); }, 20);
    printf("D pattern, rotated  %.4f ms  %.0f GB/s\n", d, bytes / d / 1e6);
    float e = time_ms([&] { k_colour_major<12, 19, false><<<(nb * N + 127) / 128, 128>>>(out, nb, N, nnz, 1.0); }, 20);
    float f = time_ms([&] { k_colour_major<12, 19, true><<<(nb * N + 127) / 128, 128>>>(out, nb, N, nnz, 1.0); }, 20);
    printf("E colour-major      %.4f ms  %.0f GB/s\n", e, bytes / e / 1e6);
    printf("F colour-major, rot %.4f ms  %.0f GB/s\n", f, bytes / f / 1e6);
    float g2 = time_ms([&] { k_chunked<12, 19, 2, true><<<(nb * N + 127) / 128, 128>>>(out, nb, N, nnz, 1.0); }, 20);
    float g4 = time_ms([&] { k_chunked<12, 19, 4, true><<<(nb * N + 127) / 128, 128>>>(out, nb, N, nnz, 1.0); }, 20);
    float g8 = time_ms([&] { k_chunked<12, 19, 8, true><<<(nb * N + 127) / 128, 128>>>(out, nb, N, nnz, 1.0); }, 20);
    float g4n = time_ms([&] { k_chunked<12, 19, 4, false><<<(nb * N + 127) / 128, 128>>>(out, nb, N, nnz, 1.0); }, 20);
    printf("G chunk 2, rot      %.4f ms  %.0f GB/s\n", g2, bytes / g2 / 1e6);
    printf("G chunk 4, rot      %.4f ms  %.0f GB/s\n", g4, bytes / g4 / 1e6);
    printf("G chunk 8, rot      %.4f ms  %.0f GB/s\n", g8, bytes / g8 / 1e6);
    printf("G chunk 4, no rot   %.4f ms  %.0f GB/s\n", g4n, bytes / g4n / 1e6);
    printf("bytes per launch %.1f MB\n", bytes / 1e6);
    printf("A contiguous        %.4f ms  %.0f GB/s\n", a, bytes / a / 1e6);
    printf("B kernel pattern    %.4f ms  %.0f GB/s\n", b, bytes / b / 1e6);
    printf("C pattern, aligned  %.4f ms  %.0f GB/s\n", c, bytes / c / 1e6);
    return 0;
}
