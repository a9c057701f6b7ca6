// Library plumbing: error text, device query, workspace size, peak-rate microbenchmarks.
#include <stdarg.h>

#include "common.cuh"

namespace tq {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static unsigned long long g_launches = 0;  // host-side statistic; the library is driven from one thread per process

unsigned long long count_launch() { return ++g_launches; }

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return (int)e;
    }
    return TQ_OK;
}

int num_sms() {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev != cached_dev) {
        cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev);
        cached_dev = dev;
        if (cached <= 0) cached = 148;
    }
    return cached;
}

// ---- microbenchmarks: measured denominators for the fused-path roofline (bench.py) ----
template <typename T>
__global__ void __launch_bounds__(256) fma_chain_kernel(int64_t iters, double* sink) {
    T a0 = (T)threadIdx.x * (T)1e-3, a1 = a0 + (T)1, a2 = a0 + (T)2, a3 = a0 + (T)3;
    T a4 = a0 + (T)4, a5 = a0 + (T)5, a6 = a0 + (T)6, a7 = a0 + (T)7;
    const T m = (T)0.999, c = (T)1e-4;
    for (int64_t i = 0; i < iters; ++i) {
        a0 = a0 * m + c; a1 = a1 * m + c; a2 = a2 * m + c; a3 = a3 * m + c;
        a4 = a4 * m + c; a5 = a5 * m + c; a6 = a6 * m + c; a7 = a7 * m + c;
    }
    T s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == (T)123456789) sink[0] = (double)s;
}

__global__ void __launch_bounds__(256) philox_chain_kernel(int64_t iters, double* sink) {
    uint32_t acc = 0;
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    for (int64_t i = 0; i < iters; ++i) {
        uint4 r = Philox::run(t, (uint32_t)i, 0u, 0u, 0x1234u, 0x5678u);
        acc ^= r.x ^ r.y ^ r.z ^ r.w;
    }
    if (acc == 0x9e3779b9u) sink[0] = (double)acc;
}

// Sector-paired RED.F64 (lanes 2i / 2i+1 -> the two words of one random bin) on a [bins, 2] fp64 table: the reduction
// pattern of the fused VEGAS pass alone, i.e. the L2 reduction rate that bounds it (bench.py, roofline of vegas*_cap).
__global__ void __launch_bounds__(256) red_pairs_kernel(double* table, uint32_t bins, int64_t iters) {
    uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    const int lane = threadIdx.x & 31;
    for (int64_t it = 0; it < iters; ++it) {
        s = s * 1664525u + 1013904223u;
        const uint32_t idx = (uint32_t)(((uint64_t)(s >> 8) * bins) >> 24);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const uint32_t b = __shfl_sync(0xffffffffu, idx, (lane >> 1) + 16 * h);
            atomicAdd(table + 2 * (size_t)b + (lane & 1), (lane & 1) ? 1.0 : 1.5);
        }
    }
}

}  // namespace tq

extern "C" {

const char* tq_last_error(void) { return tq::g_err; }

int tq_version(void) { return 100; }

uint64_t tq_kernel_launches(void) { return tq::g_launches; }

size_t tq_workspace_bytes(void) { return (size_t)32 << 20; }  // scan tile sums of 2^31 cubes need 16 MiB

int tq_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        tq::set_error("cudaGetDevice: %s", cudaGetErrorString(e));
        return (int)e;
    }
    cudaDeviceGetAttribute(sm_count, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(cc_major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(cc_minor, cudaDevAttrComputeCapabilityMinor, dev);
    return TQ_OK;
}

int tq_l2_fetch_granularity(int32_t bytes, int32_t* previous_host) {
    size_t prev = 0;
    cudaError_t e = cudaDeviceGetLimit(&prev, cudaLimitMaxL2FetchGranularity);
    if (e != cudaSuccess) { tq::set_error("cudaDeviceGetLimit: %s", cudaGetErrorString(e)); return (int)e; }
    if (previous_host) *previous_host = (int32_t)prev;
    if (bytes > 0) {
        e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)bytes);
        if (e != cudaSuccess) { tq::set_error("cudaDeviceSetLimit: %s", cudaGetErrorString(e)); return (int)e; }
    }
    return TQ_OK;
}

int tq_red_microbench(double* table, int64_t bins, int64_t iters, double* ops_out_host, void* stream) {
    TQ_REQUIRE(table && bins >= 1 && bins < (1LL << 31) && iters >= 1 && ops_out_host, "tq_red_microbench: bad arguments");
    const int grid = tq::num_sms() * 8;
    tq::red_pairs_kernel<<<TQ_GRID(grid), 256, 0, tq::as_stream(stream)>>>(table, (uint32_t)bins, iters);
    *ops_out_host = (double)grid * 256.0 * (double)iters;  // one reduction sector ({sum, count} of one bin) per thread and iteration
    return tq::check_launch("tq_red_microbench");
}

int tq_peak_microbench(int32_t kind, int64_t iters, double* sink, double* ops_out_host, void* stream) {
    const int block = 256;
    const int grid = tq::num_sms() * 8;
    const double threads = (double)grid * block;
    if (kind == 0) {
        tq::fma_chain_kernel<float><<<TQ_GRID(grid), block, 0, tq::as_stream(stream)>>>(iters, sink);
        *ops_out_host = threads * (double)iters * 8.0;  // FMA instructions (2 flop each)
    } else if (kind == 1) {
        tq::fma_chain_kernel<double><<<TQ_GRID(grid), block, 0, tq::as_stream(stream)>>>(iters, sink);
        *ops_out_host = threads * (double)iters * 8.0;
    } else if (kind == 2) {
        tq::philox_chain_kernel<<<TQ_GRID(grid), block, 0, tq::as_stream(stream)>>>(iters, sink);
        *ops_out_host = threads * (double)iters;  // Philox4x32-10 blocks
    } else {
        tq::set_error("tq_peak_microbench: unknown kind %d", kind);
        return TQ_ERR_INVALID_ARGUMENT;
    }
    return tq::check_launch("tq_peak_microbench");
}

}  // extern "C"
