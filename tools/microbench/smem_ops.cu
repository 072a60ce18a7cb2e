// Microbenchmark (B200): per-SM throughput of the shared-memory / warp primitives a block
// sort can be built from.  Prints lane-operations per clock per SM for each primitive
// with 256 threads per CTA and 1, 2, 4 CTAs per SM.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o smem_ops smem_ops.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ITER = 2048;
constexpr int BINS = 4096;

__device__ __forceinline__ uint32_t lcg(uint32_t& s) { s = s * 1664525u + 1013904223u; return s >> 8; }

template <int OP>
__global__ void __launch_bounds__(256) bench(long long* cycles, uint32_t* sink, int bins_mask) {
    __shared__ uint32_t sm[BINS];
    __shared__ float smf[BINS];
    for (int i = threadIdx.x; i < BINS; i += 256) { sm[i] = 0; smf[i] = 0.f; }
    __syncthreads();
    uint32_t s = threadIdx.x * 2654435761u + blockIdx.x * 97u + 1u;
    uint32_t acc = 0;
    const long long t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < ITER; ++i) {
        const uint32_t a = lcg(s) & bins_mask;
        if (OP == 0) atomicAdd(&sm[a], 1u);                       // RED/ATOMS without use of the result
        if (OP == 1) acc += atomicAdd(&sm[a], 1u);                // ATOMS with result
        if (OP == 2) atomicAdd(&smf[a], 1.0f);                    // float add
        if (OP == 3) acc += __match_any_sync(0xffffffffu, a & 0xff);
        if (OP == 4) sm[a] = i;                                   // plain STS random
        if (OP == 5) acc += sm[a];                                // plain LDS random
        if (OP == 6) acc += __reduce_or_sync(0xffffffffu, a);
        if (OP == 7) acc += __ballot_sync(0xffffffffu, a & 1);
        if (OP == 8) acc += __shfl_xor_sync(0xffffffffu, a, 1);
        if (OP == 9) acc += a;                                    // loop overhead only
    }
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    sink[blockIdx.x * 256 + threadIdx.x] = acc + sm[threadIdx.x] + (uint32_t)smf[threadIdx.x];
}

template <int OP>
void run(const char* name, int bins_mask) {
    long long* d_cycles; uint32_t* d_sink;
    for (int occ : {1, 2, 4}) {
        const int grid = 148 * occ;
        cudaMalloc(&d_cycles, grid * sizeof(long long));
        cudaMalloc(&d_sink, grid * 256 * sizeof(uint32_t));
        bench<OP><<<grid, 256>>>(d_cycles, d_sink, bins_mask);
        bench<OP><<<grid, 256>>>(d_cycles, d_sink, bins_mask);
        cudaDeviceSynchronize();
        long long h[148 * 4];
        cudaMemcpy(h, d_cycles, grid * sizeof(long long), cudaMemcpyDeviceToHost);
        double mean = 0;
        for (int i = 0; i < grid; ++i) mean += h[i];
        mean /= grid;
        // lane-ops per clock per SM = occ * 256 * ITER / cycles
        printf("%-28s bins=%5d ctas/SM=%d  %8.0f cyc  -> %6.2f lane-ops/clk/SM\n", name, bins_mask + 1, occ, mean,
               occ * 256.0 * ITER / mean);
        cudaFree(d_cycles); cudaFree(d_sink);
    }
}

int main() {
    run<9>("loop overhead (lcg)", 4095);
    run<0>("atomicAdd u32 no result", 4095);
    run<0>("atomicAdd u32 no result", 255);
    run<1>("atomicAdd u32 result", 4095);
    run<1>("atomicAdd u32 result", 255);
    run<2>("atomicAdd f32", 4095);
    run<2>("atomicAdd f32", 255);
    run<3>("match_any 8-bit", 4095);
    run<4>("STS random", 4095);
    run<5>("LDS random", 4095);
    run<6>("reduce_or", 4095);
    run<7>("ballot", 4095);
    run<8>("shfl_xor", 4095);
    cudaError_t e = cudaGetLastError();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
