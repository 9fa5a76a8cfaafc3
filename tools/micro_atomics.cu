// Raw ceilings of scattered 4/8-byte accesses to an L2-resident table on one B200 (no arithmetic):
// what the iterate kernel's scatter step could reach at best.  Build + run:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/micro_atomics tools/micro_atomics.cu && /tmp/micro_atomics
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(unsigned long long *tab, unsigned int mask, int iters, unsigned long long *sink)
{
    unsigned int s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    unsigned long long acc = 0;
#pragma unroll 4
    for (int i = 0; i < iters; ++i) {
        s = s * 1664525u + 1013904223u;
        const unsigned int idx = (s >> 7) & mask;
        if (MODE == 0) asm volatile("red.global.add.u32 [%0], 1;" ::"l"((unsigned int *)(tab + idx)) : "memory");
        if (MODE == 1) asm volatile("red.global.add.u64 [%0], 1;" ::"l"(tab + idx) : "memory");
        if (MODE == 2) acc += atomicAdd((unsigned int *)(tab + idx), 1u);
        if (MODE == 3) acc += atomicAdd(tab + idx, 1ull);
        if (MODE == 4) acc += __ldcg((const unsigned int *)(tab + idx));
        if (MODE == 5) acc += __ldcg(tab + idx);
        if (MODE == 6) { asm volatile("red.global.add.u64 [%0], 1;" ::"l"(tab + idx) : "memory"); acc += __ldcg((const unsigned int *)(tab + idx) + 1); }
        if (MODE == 7) acc += atomicMax(tab + idx, (unsigned long long)s << 20);
        if (MODE == 8) asm volatile("red.global.max.u64 [%0], %1;" ::"l"(tab + idx), "l"((unsigned long long)s << 20) : "memory");
        // separate arrays: count (u32) in the first half of the table, hint (u32) in the second half
        if (MODE == 9) { asm volatile("red.global.add.u32 [%0], 1;" ::"l"((unsigned int *)tab + idx) : "memory"); acc += __ldcg((const unsigned int *)tab + (size_t)mask + 1 + idx); }
        if (MODE == 10) { asm volatile("red.global.add.u32 [%0], 1;" ::"l"((unsigned int *)tab + idx) : "memory"); unsigned int h; asm volatile("ld.global.ca.u32 %0, [%1];" : "=r"(h) : "l"((const unsigned int *)tab + (size_t)mask + 1 + idx)); acc += h; }
        if (MODE == 11) { asm volatile("red.global.add.u32 [%0], 1;" ::"l"((unsigned int *)tab + idx) : "memory"); asm volatile("red.global.max.u32 [%0], %1;" ::"l"((unsigned int *)tab + (size_t)mask + 1 + idx), "r"(s) : "memory"); }
    }
    if (acc == 0x1234567887654321ull) *sink = acc;
}

// shared-memory counterpart: scattered atomics on a 32 K-entry table private to the block
template <int MODE>
__global__ void ks(int iters, unsigned long long *sink)
{
    extern __shared__ unsigned int tab[];
    const unsigned int n = 32768;
    for (unsigned int i = threadIdx.x; i < n; i += blockDim.x) tab[i] = 0;
    __syncthreads();
    unsigned int s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    unsigned long long acc = 0;
#pragma unroll 4
    for (int i = 0; i < iters; ++i) {
        s = s * 1664525u + 1013904223u;
        const unsigned int idx = (s >> 7) & (n - 1);
        if (MODE == 0) atomicAdd(tab + idx, 1u);
        if (MODE == 1) { atomicAdd(tab + idx, 1u); atomicMax(tab + ((idx + 16384u) & (n - 1)), s); }
        if (MODE == 2) acc += atomicAdd(tab + idx, 1u);
    }
    __syncthreads();
    if (acc == 0x1234567887654321ull || tab[threadIdx.x] == 0xdeadbeefu) *sink = acc;
}
template <int MODE>
static void run_smem(const char *name, unsigned long long *sink)
{
    const int iters = 4096, block = 1024, grid = 148;
    cudaFuncSetAttribute(ks<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    ks<MODE><<<grid, block, 131072>>>(64, sink);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    ks<MODE><<<grid, block, 131072>>>(iters, sink);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    const double ops = (double)grid * block * iters;
    printf("%-38s table  128 KB smem/SM, threads/SM 1024 : %8.2f G iterations/s  (%.3f lane-iterations/cycle/SM @1.965 GHz)\n", name,
           ops / ms / 1e6, ops / ms / 1e6 / 148 / 1.965);
}

template <int MODE>
static void run(const char *name, unsigned long long *tab, unsigned int mask, unsigned long long *sink, int threads_per_sm)
{
    const int iters = 4096, block = 256, grid = 148 * threads_per_sm / block;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<grid, block>>>(tab, mask, 64, sink);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<grid, block>>>(tab, mask, iters, sink);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    const double ops = (double)grid * block * iters;
    printf("%-38s table %4u MB  threads/SM %4d : %8.2f G ops/s  (%.3f lane-ops/cycle/SM @1.965 GHz)\n", name,
           (unsigned)(((size_t)mask + 1) * 8 >> 20), threads_per_sm, ops / ms / 1e6, ops / ms / 1e6 / 148 / 1.965);
}

int main()
{
    unsigned long long *tab, *sink;
    const size_t n = 1u << 25;   // up to 256 MB
    cudaMalloc(&tab, n * 8); cudaMemset(tab, 0, n * 8); cudaMalloc(&sink, 8);
    for (unsigned int mask : {(1u << 22) - 1u, (1u << 24) - 1u}) {      // 32 MB (2048^2 x 8 B), 128 MB
        for (int tps : {1024}) {
            run<0>("RED.ADD.32", tab, mask, sink, tps);
            run<1>("RED.ADD.64", tab, mask, sink, tps);
            run<8>("RED.MAX.64", tab, mask, sink, tps);
            run<2>("ATOM.ADD.32 (return)", tab, mask, sink, tps);
            run<3>("ATOM.ADD.64 (return)", tab, mask, sink, tps);
            run<7>("ATOM.MAX.64 (return)", tab, mask, sink, tps);
            run<4>("LDG.32 (ld.cg)", tab, mask, sink, tps);
            run<5>("LDG.64 (ld.cg)", tab, mask, sink, tps);
            run<6>("RED.ADD.64 + LDG.32 same word", tab, mask, sink, tps);
            run<9>("RED.ADD.32 + LDG.32 (ld.cg) 2 arrays", tab, mask, sink, tps);
            run<10>("RED.ADD.32 + LDG.32 (ld.ca) 2 arrays", tab, mask, sink, tps);
            run<11>("RED.ADD.32 + RED.MAX.32 2 arrays", tab, mask, sink, tps);
        }
    }
    run_smem<0>("ATOMS.ADD.32 (no return)", sink);
    run_smem<2>("ATOMS.ADD.32 (return)", sink);
    run_smem<1>("ATOMS.ADD.32 + ATOMS.MAX.32", sink);
    return 0;
}
