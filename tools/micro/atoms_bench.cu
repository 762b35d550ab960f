// Micro-benchmark: shared-memory atomic throughput on one SM (lanes per clock), spread addresses in an 8192-slot table.
// Tells whether k_count_skm (one 64-bit CAS per k-mer instance + one 32-bit add per repeat) is bound by the ATOMS rate or
// by instruction issue.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/atoms_bench tools/micro/atoms_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int OP>
__global__ void __launch_bounds__(512, 2) k(int iters, unsigned long long* sink) {
    extern __shared__ __align__(16) unsigned long long keys[];   // [8192]
    unsigned int* cnt = reinterpret_cast<unsigned int*>(keys + 8192);   // [8192]
    for (int i = threadIdx.x; i < 8192; i += 512) { keys[i] = 0; cnt[i] = 0; }
    __syncthreads();
    unsigned int x = threadIdx.x * 2654435761u + blockIdx.x * 40503u + 12345u;
    unsigned long long acc = 0;
    for (int i = 0; i < iters; ++i) {
        x = x * 1664525u + 1013904223u;
        const unsigned int s = (x >> 13) & 8191u;
        if (OP == 0) acc += cnt[s];
        else if (OP == 1) cnt[s] = x;
        else if (OP == 2) atomicAdd(&cnt[s], 1u);
        else if (OP == 3) acc += atomicAdd(&cnt[s], 1u);
        else if (OP == 4) atomicOr(&cnt[s], 1u << (x & 31));
        else if (OP == 5) acc += atomicCAS(&cnt[s], 0u, x | 1u);
        else if (OP == 6) acc += atomicCAS(&keys[s], 0ull, ((unsigned long long)x << 32) | x | 1ull);
        else if (OP == 7) acc += atomicExch(&keys[s], (unsigned long long)x);
        else if (OP == 8) acc += keys[s];
        else if (OP == 9) { acc += atomicCAS(&keys[s], 0ull, ((unsigned long long)x << 32) | x | 1ull); atomicAdd(&cnt[s], 1u); }
    }
    if (acc == 0x123456789ull) *sink = acc;
}

template <int OP> void run(const char* name, int iters, unsigned long long* sink) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaFuncSetAttribute(k<OP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 12);
    k<OP><<<296, 512, 8192 * 12>>>(iters, sink);
    cudaEventRecord(a);
    k<OP><<<296, 512, 8192 * 12>>>(iters, sink);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double ops = 296.0 * 512 * iters;
    printf("%-28s %8.3f ms  %7.1f G lane-ops/s  %.2f lanes/clk/SM (1.965 GHz, 148 SMs)\n", name, ms, ops / ms / 1e6, ops / (ms * 1e-3) / 148 / 1.965e9);
}

int main() {
    unsigned long long* sink; cudaMalloc(&sink, 8);
    const int it = 20000;
    run<0>("LDS.32", it, sink); run<8>("LDS.64", it, sink); run<1>("STS.32", it, sink);
    run<2>("ATOMS.ADD (no return)", it, sink); run<3>("ATOMS.ADD (return)", it, sink); run<4>("ATOMS.OR", it, sink);
    run<5>("ATOMS.CAS.32", it, sink); run<6>("ATOMS.CAS.64", it, sink); run<7>("ATOMS.EXCH.64", it, sink);
    run<9>("CAS.64 + ADD.32", it, sink);
    return 0;
}
