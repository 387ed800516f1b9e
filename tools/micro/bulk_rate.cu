// Microbenchmark: rate of small cp.async.bulk (global -> shared, 1D) copies per SM,
// against the same bytes fetched with per-thread 8-byte LDGSTS and with plain loads.
// Shape of the engine's reduced-cost gather: CTAs of 128 threads, 7 per SM; every
// thread fetches one run of RUN doubles (a column of P) from a scattered position.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template<int MODE, int RUN>
__global__ void __launch_bounds__(128, 7)
k(const double* __restrict__ P, const int* __restrict__ start, int iters, double* out, size_t per_cta)
{
    extern __shared__ __align__(16) unsigned char smem[];
    double* st = reinterpret_cast<double*>(smem);            // [128 * RUN]
    __shared__ __align__(8) unsigned long long bar;
    const int tid = threadIdx.x;
    const double* base = P + (size_t)blockIdx.x * per_cta;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();
    double acc = 0;
    unsigned phase = 0;
    for (int it = 0; it < iters; ++it) {
        const int s = start[(it * 128 + tid) & 0xffff];
        const double* src = base + s;
        if (MODE == 0) {
            if (tid == 0)
                asm volatile("mbarrier.arrive.expect_tx.shared.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(128 * RUN * 8));
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(st + tid * RUN)), "l"(src), "r"(RUN * 8), "r"(smem_u32(&bar)) : "memory");
            unsigned done = 0;
            while (!done)
                asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                             : "=r"(done) : "r"(smem_u32(&bar)), "r"(phase) : "memory");
            phase ^= 1;
        } else if (MODE == 1) {
#pragma unroll
            for (int j = 0; j < RUN; ++j)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(st + tid * RUN + j)), "l"(src + j) : "memory");
            asm volatile("cp.async.wait_all;" ::: "memory");
            __syncthreads();
        } else {
            double v[RUN];
#pragma unroll
            for (int j = 0; j < RUN; ++j)
                v[j] = src[j];
#pragma unroll
            for (int j = 0; j < RUN; ++j)
                st[tid * RUN + j] = v[j];
            __syncthreads();
        }
        acc += st[((tid + 1) & 127) * RUN];
        __syncthreads();
    }
    out[blockIdx.x * 128 + tid] = acc;
}

template<int MODE, int RUN>
void run(const char* name, const double* P, const int* start, double* out, size_t per_cta, int ctas)
{
    const int iters = 2000;
    cudaFuncSetAttribute(k<MODE, RUN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 30 * 1024);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    k<MODE, RUN><<<ctas, 128, 30 * 1024>>>(P, start, 10, out, per_cta);
    cudaEventRecord(a);
    k<MODE, RUN><<<ctas, 128, 30 * 1024>>>(P, start, iters, out, per_cta);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    const double copies = (double)ctas * 128 * iters;
    printf("%-28s run=%2d doubles  %8.2f ms  %7.1f M runs/s  %6.1f cycles per run per SM  %7.1f GB/s  (%s)\n", name, RUN,
           ms, copies / ms / 1e3, ms * 1e-3 * 1.965e9 / (copies / 148.0), copies * RUN * 8 / ms / 1e6,
           cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    const int ctas = 148 * 7;
    const size_t per_cta = 65536 + 64; // doubles: 0.5 MB of P per lane, like the SPP workload
    double* P;
    cudaMalloc(&P, per_cta * ctas * sizeof(double));
    cudaMemset(P, 0, per_cta * ctas * sizeof(double));
    int* h = new int[65536];
    unsigned s = 12345;
    for (int i = 0; i < 65536; ++i) {
        s = s * 1664525u + 1013904223u;
        h[i] = (int)((s >> 8) % 65536) & ~1; // 16-byte aligned runs
    }
    int* start;
    cudaMalloc(&start, 65536 * sizeof(int));
    cudaMemcpy(start, h, 65536 * sizeof(int), cudaMemcpyHostToDevice);
    double* out;
    cudaMalloc(&out, ctas * 128 * sizeof(double));
    run<0, 6>("cp.async.bulk per thread", P, start, out, per_cta, ctas);
    run<1, 6>("LDGSTS 8 B per entry", P, start, out, per_cta, ctas);
    run<2, 6>("LDG + STS", P, start, out, per_cta, ctas);
    run<0, 8>("cp.async.bulk per thread", P, start, out, per_cta, ctas);
    run<1, 8>("LDGSTS 8 B per entry", P, start, out, per_cta, ctas);
    run<2, 8>("LDG + STS", P, start, out, per_cta, ctas);
    run<0, 16>("cp.async.bulk per thread", P, start, out, per_cta, ctas);
    run<2, 16>("LDG + STS", P, start, out, per_cta, ctas);
    return 0;
}
