// Shared device/host helpers for the tqb200 kernels (sm_100a).
// Nothing here is a port: the reference (esa/torchquad) is pure Python over ATen.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <math.h>

#include "../../include/tqb200.h"

namespace tq {

// ---------------------------------------------------------------- errors
void set_error(const char* fmt, ...);
int check_launch(const char* what);

#define TQ_REQUIRE(cond, ...)                \
    do {                                     \
        if (!(cond)) {                       \
            tq::set_error(__VA_ARGS__);      \
            return TQ_ERR_INVALID_ARGUMENT;  \
        }                                    \
    } while (0)

int num_sms();

// Every kernel launch of the library writes its grid as TQ_GRID(...): the wrapper counts the launch
// (tq_kernel_launches, reported by bench.py as gpu_launches) and evaluates to the grid.
unsigned long long count_launch();
#define TQ_GRID(...) (tq::count_launch(), (__VA_ARGS__))

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// Persistent-style grid: enough CTAs to fill every SM `per_sm` times, never more than needed.
static inline int grid_for(int64_t work_items, int block, int per_sm) {
    int64_t need = (work_items + block - 1) / block;
    int64_t cap = (int64_t)num_sms() * per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

// ---------------------------------------------------------------- exact (non-contracted) arithmetic
// The reference evaluates `a*b` and `+ c` as two ATen kernels; nvcc would fuse them into an FMA and
// change the last bit, which flips floor() results (bin ids) and breaks bit parity.
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ float div_rn(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ double div_rn(double a, double b) { return __ddiv_rn(a, b); }

// a / b for a divisor b that is constant over the launch, given inv = RN(1/b): q0 = RN(a*inv), r = a - q0*b (exact, FMA),
// q = RN(q0 + r*inv) is the CORRECTLY ROUNDED quotient (Markstein 1990: inv correctly rounded, q0 faithful), i.e. bit-identical
// to __fdiv_rn/__ddiv_rn for the operands of the stratified sampling (a = digit + u in [0, N_strat), b = N_strat <= 1000;
// checked exhaustively in u for fp32 and on 4e8 random operands for fp64, tests/test_host_logic.py).  Three pipelined
// operations instead of the ~30-instruction fp64 division sequence.
__device__ __forceinline__ float div_by_const(float a, float b, float inv) {
    const float q0 = __fmul_rn(a, inv);
    return __fmaf_rn(__fmaf_rn(-q0, b, a), inv, q0);
}
__device__ __forceinline__ double div_by_const(double a, double b, double inv) {
    const double q0 = __dmul_rn(a, inv);
    return __fma_rn(__fma_rn(-q0, b, a), inv, q0);
}

// Exact 32-bit division by a runtime constant (Granlund-Montgomery round-up form): q = x / d for all x.
struct FastDiv {
    uint32_t d, m, s;
    __host__ __device__ void set(uint32_t d_) {
        d = d_;
        if (d_ <= 1) { m = 0; s = 0; return; }
        uint32_t l = 0;
        while ((1ull << l) < d_) ++l;  // ceil(log2 d)
        m = (uint32_t)((((1ull << l) - d_) << 32) / d_ + 1);
        s = l;
    }
    __device__ __forceinline__ uint32_t div(uint32_t x) const {
        if (d <= 1) return x;
        const uint32_t t = __umulhi(m, x);
        return (t + ((x - t) >> 1)) >> (s - 1);
    }
    // same for callers that know d >= 2 (no test, no branch)
    __device__ __forceinline__ uint32_t div_ge2(uint32_t x) const {
        const uint32_t t = __umulhi(m, x);
        return (t + ((x - t) >> 1)) >> (s - 1);
    }
};


// ---------------------------------------------------------------- Philox4x32-10
// Salmon et al. SC'11.  key = 64-bit seed, counter = 4 x u32.  Known-answer vectors are checked in
// tests/test_philox.py against oracle/ref_oracle.py:philox4x32_10.
struct Philox {
    static constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    static constexpr uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;

    __device__ __forceinline__ static uint4 run(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                uint32_t k0, uint32_t k1) {
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
            uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
            uint32_t n0 = hi1 ^ c1 ^ k0;
            uint32_t n2 = hi0 ^ c3 ^ k1;
            c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
            k0 += W0; k1 += W1;
        }
        return make_uint4(c0, c1, c2, c3);
    }
};

// Uniform lanes per Philox block and the u32 -> [0,1) conversion for each working type.
template <typename T> struct U01;
template <> struct U01<float> {
    static constexpr int LANES = 4;
    __device__ __forceinline__ static void convert(const uint4& r, float* u) {
        const float s = 5.9604644775390625e-08f;  // 2^-24
        u[0] = __uint2float_rn(r.x >> 8) * s;
        u[1] = __uint2float_rn(r.y >> 8) * s;
        u[2] = __uint2float_rn(r.z >> 8) * s;
        u[3] = __uint2float_rn(r.w >> 8) * s;
    }
};
template <> struct U01<double> {
    static constexpr int LANES = 2;
    __device__ __forceinline__ static void convert(const uint4& r, double* u) {
        const double s = 1.1102230246251565e-16;  // 2^-53
        unsigned long long a = ((unsigned long long)r.y << 32) | r.x;
        unsigned long long b = ((unsigned long long)r.w << 32) | r.z;
        u[0] = __ull2double_rn(a >> 11) * s;
        u[1] = __ull2double_rn(b >> 11) * s;
    }
};

// Stream layout (DESIGN.md "Philox layout"): counter = (i0, i1, block, call).
//   row-keyed    (RNG.uniform, Monte Carlo, VEGAS warm-up): i0 = row & 0xffffffff, i1 = row >> 32
//   cube-keyed   (VEGAS stratified sampling):               i0 = cube, i1 = sample index within the cube
template <typename T>
__device__ __forceinline__ void philox_block(uint64_t seed, uint32_t call, uint32_t i0, uint32_t i1,
                                             uint32_t blk, T* u) {
    uint4 r = Philox::run(i0, i1, blk, call, (uint32_t)seed, (uint32_t)(seed >> 32));
    U01<T>::convert(r, u);
}

// ---------------------------------------------------------------- reductions
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

// Deterministic block sum of NV doubles per thread; result valid in thread 0.  `sh` needs 32*NV doubles.
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double* sh) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) sh[warp * NV + i] = v[i];
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            double x = lane < nwarps ? sh[lane * NV + i] : 0.0;
            v[i] = warp_sum(x);
        }
    }
    __syncthreads();
}

// Two-stage deterministic grid reduction: every CTA stores its partials, the last CTA to arrive
// (threadfence + ticket) sums them in CTA order and writes out[0..NV).  `partials` holds gridDim.x*NV
// doubles, `ticket` is a zero-initialised u32 that the last CTA resets so the workspace is reusable.
template <int NV>
__device__ __forceinline__ void grid_sum_finish(double (&v)[NV], double* sh, double* partials,
                                                unsigned int* ticket, double* out) {
    block_sum<NV>(v, sh);
    __shared__ bool is_last;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) partials[(size_t)blockIdx.x * NV + i] = v[i];
        __threadfence();
        unsigned int t = atomicAdd(ticket, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        double acc[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) acc[i] = 0.0;
        for (unsigned int b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
#pragma unroll
            for (int i = 0; i < NV; ++i) acc[i] += __ldcg(&partials[(size_t)b * NV + i]);
        }
        block_sum<NV>(acc, sh);
        if (threadIdx.x == 0) {
#pragma unroll
            for (int i = 0; i < NV; ++i) out[i] = acc[i];
            *ticket = 0u;
        }
    }
}

// The same reduction of two values with separate destinations (out1 may be NULL).
__device__ __forceinline__ void grid_sum_finish_split(double (&v)[2], double* sh, double* partials, unsigned int* ticket,
                                                      double* out0, double* out1) {
    block_sum<2>(v, sh);
    __shared__ bool is_last;
    if (threadIdx.x == 0) {
        partials[(size_t)blockIdx.x * 2] = v[0];
        partials[(size_t)blockIdx.x * 2 + 1] = v[1];
        __threadfence();
        is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        double acc[2] = {0.0, 0.0};
        for (unsigned int b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
            acc[0] += __ldcg(&partials[(size_t)b * 2]);
            acc[1] += __ldcg(&partials[(size_t)b * 2 + 1]);
        }
        block_sum<2>(acc, sh);
        if (threadIdx.x == 0) {
            *out0 = acc[0];
            if (out1) *out1 = acc[1];
            *ticket = 0u;
        }
    }
}

// ---------------------------------------------------------------- scans
template <typename V>
__device__ __forceinline__ V warp_incl_scan(V v) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        V t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// Exclusive prefix of one value per thread over the CTA (blockDim.x multiple of 32, <= 1024);
// `total` receives the CTA sum in every thread.  `sh` needs 33 entries.
template <typename V>
__device__ __forceinline__ V block_excl_scan(V v, V* sh, V& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    V incl = warp_incl_scan(v);
    V excl = __shfl_up_sync(0xffffffffu, incl, 1);  // no subtraction: exact for floating point
    if (lane == 0) excl = (V)0;
    if (lane == 31) sh[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        V w = lane < nwarps ? sh[lane] : (V)0;
        V wi = warp_incl_scan(w);
        V we = __shfl_up_sync(0xffffffffu, wi, 1);
        sh[lane] = lane == 0 ? (V)0 : we;  // exclusive prefix of warp sums
        if (lane == 31) sh[32] = wi;
    }
    __syncthreads();
    V res = sh[warp] + excl;
    total = sh[32];
    __syncthreads();
    return res;
}

// ---------------------------------------------------------------- workspace
// The caller (Python host) owns one scratch buffer per device and passes it to every call that needs
// temporaries; the library never allocates device memory.
struct Workspace {
    char* base;
    size_t bytes;
    size_t used;
    Workspace(void* p, size_t n) : base((char*)p), bytes(n), used(0) {}
    template <typename U> U* take(size_t count) {
        size_t off = (used + 255) & ~(size_t)255;
        size_t end = off + count * sizeof(U);
        if (base == nullptr || end > bytes) return nullptr;
        used = end;
        return reinterpret_cast<U*>(base + off);
    }
};

// First TQ_WS_HEADER bytes of the workspace: zero-initialised tickets/flags (the host zeroes the buffer once
// at allocation; kernels restore zeros before they exit).
static constexpr size_t WS_HEADER = 256;

template <typename T> __host__ __device__ inline int dtype_of();
template <> __host__ __device__ inline int dtype_of<float>() { return TQ_F32; }
template <> __host__ __device__ inline int dtype_of<double>() { return TQ_F64; }

#define TQ_DISPATCH_DTYPE(dtype, ...)                                    \
    do {                                                                 \
        if ((dtype) == TQ_F32) { using T = float; __VA_ARGS__; }         \
        else if ((dtype) == TQ_F64) { using T = double; __VA_ARGS__; }   \
        else { tq::set_error("unsupported dtype %d", (int)(dtype)); return TQ_ERR_INVALID_ARGUMENT; } \
    } while (0)

}  // namespace tq
