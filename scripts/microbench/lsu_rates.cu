// Scattered-access rates of one B200 SM's load/store path, measured to size the fused VEGAS pass:
// global reductions (RED) in several lane->sector patterns, 128-bit gathers from global and shared memory,
// shared-memory atomics.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/lsu_rates lsu_rates.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t next(uint32_t& s) { s = s * 1664525u + 1013904223u; return s >> 8; }

enum { RED_F64 = 0, RED_F64_U64, RED_F64_PAIRED, RED_V2F32, RED_F32_U32, RED_F32, LDG128, LDG128_NC, LDS128, ATOMS_U32, ATOMS_F32,
       ATOMS_F64, ATOMS_U64, RED_F64_QUAD, LDG64, RED_U32, KINDS };
static const char* NAMES[] = {"red.f64 (1 lane/sample)", "red.f64 + red.u64 adjacent words (2 instr/sample)",
                              "red.f64 lane pairs {w,c} (1 instr / 16 samples)", "red.v2.f32 {w,c}", "red.f32 + red.u32 adjacent",
                              "red.f32", "ldg.128 gather (ld.global)", "ldg.128 gather (ld.global.nc)", "lds.128 gather (smem)",
                              "atoms.u32 spread", "atoms.f32 spread", "atoms.f64 spread", "atoms.u64 spread",
                              "red.f64 lane quads same sector (1 instr / 8 samples x 4 words)", "ldg.64 gather (nc)", "red.u32"};

template <int KIND>
__global__ void __launch_bounds__(256) k(void* tab, uint32_t mask, int iters, double* sink) {
    extern __shared__ __align__(16) unsigned char sm[];
    uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    double acc = 0.0;
    const int lane = threadIdx.x & 31;
    if (KIND == LDS128 || (KIND >= ATOMS_U32 && KIND <= ATOMS_U64)) {
        for (int i = threadIdx.x; i < (int)(mask + 1) * 4; i += blockDim.x) ((uint32_t*)sm)[i] = 0;
        __syncthreads();
    }
#pragma unroll 4
    for (int it = 0; it < iters; ++it) {
        uint32_t idx = next(s) & mask;  // bin
        if (KIND == RED_F64) {
            atomicAdd((double*)tab + 2 * (size_t)idx, 1.5);
        } else if (KIND == RED_F64_U64) {
            atomicAdd((double*)tab + 2 * (size_t)idx, 1.5);
            atomicAdd((unsigned long long*)tab + 2 * (size_t)idx + 1, 1ull);
        } else if (KIND == RED_F64_PAIRED) {
            // lanes 2i, 2i+1 serve sample i of a 16-sample group: two rounds cover the warp's 32 samples
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const uint32_t b = __shfl_sync(0xffffffffu, idx, (lane >> 1) + 16 * h);
                atomicAdd((double*)tab + 2 * (size_t)b + (lane & 1), (lane & 1) ? 1.0 : 1.5);
            }
        } else if (KIND == RED_F64_QUAD) {
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                const uint32_t b = __shfl_sync(0xffffffffu, idx, (lane >> 2) + 8 * h);
                atomicAdd((double*)tab + 4 * (size_t)b + (lane & 3), 1.5);
            }
        } else if (KIND == RED_V2F32) {
            float* p = (float*)tab + 2 * (size_t)idx;
            asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(1.5f), "f"(1.0f) : "memory");
        } else if (KIND == RED_F32_U32) {
            atomicAdd((float*)tab + 2 * (size_t)idx, 1.5f);
            atomicAdd((unsigned int*)tab + 2 * (size_t)idx + 1, 1u);
        } else if (KIND == RED_F32) {
            atomicAdd((float*)tab + 2 * (size_t)idx, 1.5f);
        } else if (KIND == RED_U32) {
            atomicAdd((unsigned int*)tab + 2 * (size_t)idx, 1u);
        } else if (KIND == LDG128) {
            double2 v;
            asm volatile("ld.global.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"((const double2*)tab + idx));
            acc += v.x + v.y;
        } else if (KIND == LDG128_NC) {
            const double2 v = __ldg((const double2*)tab + idx);
            acc += v.x + v.y;
        } else if (KIND == LDG64) {
            const float2 v = __ldg((const float2*)tab + idx);
            acc += v.x + v.y;
        } else if (KIND == LDS128) {
            const double2 v = ((const double2*)sm)[idx];
            acc += v.x + v.y;
            s += (uint32_t)__double2loint(acc) & 1u;
        } else if (KIND == ATOMS_U32) {
            atomicAdd((unsigned int*)sm + idx, 1u);
        } else if (KIND == ATOMS_F32) {
            atomicAdd((float*)sm + idx, 1.5f);
        } else if (KIND == ATOMS_F64) {
            atomicAdd((double*)sm + idx, 1.5);
        } else if (KIND == ATOMS_U64) {
            atomicAdd((unsigned long long*)sm + idx, 1ull);
        }
    }
    if (acc == 123.456) sink[0] = acc;
}

template <int KIND>
static void run(const char* tag, void* tab, uint32_t bins, int ctas_per_sm, size_t smem, double samples_per_iter_thread, double* sink,
                int sms, double clk_ghz) {
    const int iters = 2000;
    const int grid = sms * ctas_per_sm;
    if (smem > 48 * 1024) CK(cudaFuncSetAttribute(k<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    k<KIND><<<grid, 256, smem>>>(tab, bins - 1, 200, sink);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    k<KIND><<<grid, 256, smem>>>(tab, bins - 1, iters, sink);
    CK(cudaEventRecord(b));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, a, b));
    const double samples = (double)grid * 256 * iters * samples_per_iter_thread;
    const double rate = samples / (ms * 1e-3);
    printf("%-66s %-10s bins=%-9u ctas/sm=%d  %8.3f ms  %10.3e samples/s  %6.2f cyc/sample/SM\n", NAMES[KIND], tag, bins, ctas_per_sm, ms, rate,
           sms * clk_ghz * 1e9 / rate);
    fflush(stdout);
}

int main() {
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    int clk_khz = 0;
    CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    const int sms = p.multiProcessorCount;
    const double ghz = clk_khz * 1e-6;
    printf("%s, %d SMs, %.3f GHz (cyc figures assume this clock)\n", p.name, sms, ghz);
    const size_t big_bins = (size_t)1 << 26;  // 64 Mi bins x 32 B = 2 GiB
    void* tab;
    CK(cudaMalloc(&tab, big_bins * 32));
    CK(cudaMemset(tab, 0, big_bins * 32));
    double* sink;
    CK(cudaMalloc(&sink, 8));
    for (uint32_t bins : {4096u, 32768u, 1u << 20, 1u << 26}) {
        for (int cps : {4, 8}) {
            const char* tag = bins <= (1u << 20) ? "L2" : "DRAM";
            run<RED_F64>(tag, tab, bins, cps, 0, 1, sink, sms, ghz);
            run<RED_F64_U64>(tag, tab, bins, cps, 0, 1, sink, sms, ghz);
            run<RED_F64_PAIRED>(tag, tab, bins, cps, 0, 1, sink, sms, ghz);
            run<RED_F64_QUAD>(tag, tab, bins >> 1, cps, 0, 1, sink, sms, ghz);
            run<RED_V2F32>(tag, tab, bins, cps, 0, 1, sink, sms, ghz);
            run<RED_F32_U32>(tag, tab, bins, cps, 0, 1, sink, sms, ghz);
            run<RED_F32>(tag, tab, bins, cps, 0, 1, sink, sms, ghz);
            run<RED_U32>(tag, tab, bins, cps, 0, 1, sink, sms, ghz);
            run<LDG128>(tag, tab, bins, cps, 0, 1, sink, sms, ghz);
            run<LDG128_NC>(tag, tab, bins, cps, 0, 1, sink, sms, ghz);
            run<LDG64>(tag, tab, bins, cps, 0, 1, sink, sms, ghz);
        }
    }
    for (int cps : {2, 4}) {
        run<LDS128>("smem", tab, 2048, cps, 2048 * 16, 1, sink, sms, ghz);
        run<ATOMS_U32>("smem", tab, 2048, cps, 2048 * 16, 1, sink, sms, ghz);
        run<ATOMS_F32>("smem", tab, 2048, cps, 2048 * 16, 1, sink, sms, ghz);
        run<ATOMS_F64>("smem", tab, 2048, cps, 2048 * 16, 1, sink, sms, ghz);
        run<ATOMS_U64>("smem", tab, 2048, cps, 2048 * 16, 1, sink, sms, ghz);
    }
    return 0;
}
