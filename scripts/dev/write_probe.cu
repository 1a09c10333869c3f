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

// store flavours for the colour-major pattern: 0 = st.global.cs (what the kernels use), 1 = plain, 2 = .cg, 3 = .wt
template <int MODE> __device__ __forceinline__ void st_mode(double* p, double v)
{
    if (MODE == 0) __stcs(p, v);
    else if (MODE == 1) *p = v;
    else if (MODE == 2) __stcg(p, v);
    else __stwt(p, v);
}
template <int NROW, int NB, int MODE>
__global__ void __launch_bounds__(128) k_colour_major_mode(double* out, int nb, int N, size_t pitch, double v)
{
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)nb * N) return;
    const int b = (int)(gid / N);
    const int k = (int)(gid - (long long)b * N);
    double* vb = out + (size_t)b * pitch + k;
#pragma unroll 1
    for (int cc = 0; cc < NB; ++cc)
#pragma unroll
        for (int i = 0; i < NROW; ++i) st_mode<MODE>(vb + (unsigned)(i * NB + cc) * (unsigned)N, v + cc);
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

// H: staged write-out.  CTA = 128 threads = 2 instances; per pass of CH colours the CTA holds 2 x NROW runs of
// CH*N contiguous doubles (row i, colours c0..c0+CH-1) in shared memory and writes each run contiguously:
// BULK = false: warps store 256-byte pieces back to back; BULK = true: one cp.async.bulk per run.
template <int NROW, int NB, int CH, bool BULK>
__global__ void __launch_bounds__(128) k_staged(double* out, int nb, int N, size_t pitch, double v)
{
    extern __shared__ __align__(128) double sm[];
    const int b0 = blockIdx.x * 2;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int c0 = 0; c0 < NB; c0 += CH) {
        const int ch = c0 + CH <= NB ? CH : NB - c0;
        // "compute": every thread deposits its CH x NROW values (conflict-free: consecutive nodes)
        const int inst = threadIdx.x / N, k = threadIdx.x % N;
        for (int i = 0; i < NROW; ++i)
            for (int c = 0; c < ch; ++c) sm[((inst * NROW + i) * CH + c) * N + k] = v + c;
        if (BULK) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        const int run_len = ch * N;
        if (BULK) {
            if (threadIdx.x < 2 * NROW) {
                const int r = threadIdx.x, ins = r / NROW, i = r % NROW;
                if (b0 + ins < nb) {
                    double* dst = out + (size_t)(b0 + ins) * pitch + (size_t)(i * NB + c0) * N;
                    const unsigned src = (unsigned)__cvta_generic_to_shared(sm + (size_t)r * CH * N);
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(run_len * 8) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                }
            }
        } else {
            for (int r = warp; r < 2 * NROW; r += 4) {
                const int ins = r / NROW, i = r % NROW;
                if (b0 + ins >= nb) continue;
                double* dst = out + (size_t)(b0 + ins) * pitch + (size_t)(i * NB + c0) * N;
                const double* src = sm + (size_t)r * CH * N;
                for (int j = lane; j < run_len; j += 32) __stcs(dst + j, src[j]);
            }
        }
        __syncthreads();
    }
}

// I: two adjacent nodes per thread, 16-byte stores: a warp writes a whole 512-byte block per instruction
template <int NROW, int NB>
__global__ void __launch_bounds__(128) k_pair(double* out, int nb, int N, size_t pitch, double v)
{
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int half = N / 2;
    if (gid >= (long long)nb * half) return;
    const int b = (int)(gid / half);
    const int k = 2 * (int)(gid - (long long)b * half);
    double* vb = out + (size_t)b * pitch + k;
#pragma unroll 1
    for (int cc = 0; cc < NB; ++cc)
#pragma unroll
        for (int i = 0; i < NROW; ++i) __stcs(reinterpret_cast<double2*>(vb + (unsigned)(i * NB + cc) * (unsigned)N), make_double2(v + cc, v));
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
    {
        const int CH = 5;
        const size_t shm = (size_t)2 * 12 * CH * N * 8;
        cudaFuncSetAttribute(k_staged<12, 19, CH, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm);
        cudaFuncSetAttribute(k_staged<12, 19, CH, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm);
        float h1 = time_ms([&] { k_staged<12, 19, CH, false><<<nb / 2, 128, shm>>>(out, nb, N, nnz, 1.0); }, 20);
        float h2 = time_ms([&] { k_staged<12, 19, CH, true><<<nb / 2, 128, shm>>>(out, nb, N, nnz, 1.0); }, 20);
        printf("H staged 5 colours, warp stores  %.4f ms  %.0f GB/s  (err %s)\n", h1, bytes / h1 / 1e6, cudaGetErrorString(cudaGetLastError()));
        printf("H staged 5 colours, bulk copies  %.4f ms  %.0f GB/s  (err %s)\n", h2, bytes / h2 / 1e6, cudaGetErrorString(cudaGetLastError()));
    }
    {
        const int CH = 10;
        const size_t shm = (size_t)2 * 12 * CH * N * 8;
        cudaFuncSetAttribute(k_staged<12, 19, CH, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm);
        cudaFuncSetAttribute(k_staged<12, 19, CH, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm);
        float h1 = time_ms([&] { k_staged<12, 19, CH, false><<<nb / 2, 128, shm>>>(out, nb, N, nnz, 1.0); }, 20);
        float h2 = time_ms([&] { k_staged<12, 19, CH, true><<<nb / 2, 128, shm>>>(out, nb, N, nnz, 1.0); }, 20);
        printf("H staged 10 colours, warp stores %.4f ms  %.0f GB/s  (err %s)\n", h1, bytes / h1 / 1e6, cudaGetErrorString(cudaGetLastError()));
        printf("H staged 10 colours, bulk copies %.4f ms  %.0f GB/s  (err %s)\n", h2, bytes / h2 / 1e6, cudaGetErrorString(cudaGetLastError()));
    }
    float pi = time_ms([&] { k_pair<12, 19><<<(nb * N / 2 + 127) / 128, 128>>>(out, nb, N, nnz, 1.0); }, 20);
    printf("I pairs, 16 B stores, colour-major %.4f ms  %.0f GB/s\n", pi, bytes / pi / 1e6);
    {
        float m0 = time_ms([&] { k_colour_major_mode<12, 19, 0><<<(nb * N + 127) / 128, 128>>>(out, nb, N, nnz, 1.0); }, 20);
        float m1 = time_ms([&] { k_colour_major_mode<12, 19, 1><<<(nb * N + 127) / 128, 128>>>(out, nb, N, nnz, 1.0); }, 20);
        float m2 = time_ms([&] { k_colour_major_mode<12, 19, 2><<<(nb * N + 127) / 128, 128>>>(out, nb, N, nnz, 1.0); }, 20);
        float m3 = time_ms([&] { k_colour_major_mode<12, 19, 3><<<(nb * N + 127) / 128, 128>>>(out, nb, N, nnz, 1.0); }, 20);
        printf("colour-major store flavours: .cs %.0f GB/s, plain %.0f GB/s, .cg %.0f GB/s, .wt %.0f GB/s\n", bytes / m0 / 1e6, bytes / m1 / 1e6, bytes / m2 / 1e6, bytes / m3 / 1e6);
        float b64 = time_ms([&] { k_colour_major_mode<12, 19, 0><<<(nb * N + 63) / 64, 64>>>(out, nb, N, nnz, 1.0); }, 20);
        printf("colour-major .cs with 64-thread CTAs: %.0f GB/s\n", bytes / b64 / 1e6);
    }
    printf("bytes per launch %.1f MB\n", bytes / 1e6);
    printf("A contiguous        %.4f ms  %.0f GB/s\n", a, bytes / a / 1e6);
    printf("B kernel pattern    %.4f ms  %.0f GB/s\n", b, bytes / b / 1e6);
    printf("C pattern, aligned  %.4f ms  %.0f GB/s\n", c, bytes / c / 1e6);
    return 0;
}
