// Newton-Cotes tensor-product grids: point generation from 1-D nodes and the weighted contraction.
// Replaces IntegrationGrid (integration_grid.py:79-99: meshgrid + ravel + stack), the einsum/reshape of
// GridIntegrator.calculate_result (grid_integrator.py:70-82) and the per-axis stencil passes of
// trapezoid.py:28-37 / simpson.py:30-46 / boole.py:30-48, which are one weighted sum
//   sum_p f(p) * prod_d w[d, i_d(p)],   i_d(p) = (p / n^(dim-1-d)) % n   (dim 0 slowest).
#include "common.cuh"

namespace tq {

// Multi-index of point p in base n; dim 0 is the most significant digit (integration_grid.py:98-99).
struct GridIndex {
    uint32_t n;
    int dim;
};

// One thread per (row, d) element so that the row-major store is coalesced; the digit is extracted with a
// division by the precomputed stride n^(dim-1-d).
template <typename T>
__global__ void __launch_bounds__(256)
grid_points_kernel(const T* __restrict__ nodes, uint32_t n, int dim, int64_t p_begin, int64_t p_end,
                   T* __restrict__ out) {
    __shared__ uint64_t s_stride[TQ_MAX_DIM];
    if (threadIdx.x == 0) {
        uint64_t s = 1;
        for (int d = dim - 1; d >= 0; --d) { s_stride[d] = s; s *= n; }
    }
    __syncthreads();
    const int64_t total = (p_end - p_begin) * dim;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = e / dim;
        const int d = (int)(e - r * dim);
        const uint64_t p = (uint64_t)(p_begin + r);
        const uint32_t i = (uint32_t)((p / s_stride[d]) % n);
        out[e] = __ldg(&nodes[(int64_t)d * n + i]);
    }
}

// grad_nodes[d, j] += grad_points[p, d] for i_d(p) = j: shared-memory privatised [dim, n] table per CTA.
template <typename T>
__global__ void __launch_bounds__(256)
grid_points_backward_kernel(const T* __restrict__ g, uint32_t n, int dim, int64_t p_begin, int64_t p_end,
                            double* __restrict__ grad_nodes, bool use_smem) {
    extern __shared__ double s_acc[];  // [dim*n]
    __shared__ uint64_t s_stride[TQ_MAX_DIM];
    if (use_smem)
        for (int i = threadIdx.x; i < dim * (int)n; i += blockDim.x) s_acc[i] = 0.0;
    if (threadIdx.x == 0) {
        uint64_t s = 1;
        for (int d = dim - 1; d >= 0; --d) { s_stride[d] = s; s *= n; }
    }
    __syncthreads();
    const int64_t total = (p_end - p_begin) * dim;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = e / dim;
        const int d = (int)(e - r * dim);
        const uint64_t p = (uint64_t)(p_begin + r);
        const uint32_t i = (uint32_t)((p / s_stride[d]) % n);
        if (use_smem) atomicAdd(&s_acc[d * n + i], (double)g[e]);
        else atomicAdd(&grad_nodes[(int64_t)d * n + i], (double)g[e]);
    }
    if (!use_smem) return;
    __syncthreads();
    for (int i = threadIdx.x; i < dim * (int)n; i += blockDim.x)
        if (s_acc[i] != 0.0) atomicAdd(&grad_nodes[i], s_acc[i]);
}

// The [dim, n] table of 1-D weights is staged in shared memory when it fits (always, except 1-D grids
// with very many nodes), otherwise read through the read-only path.
constexpr size_t NC_TABLE_SMEM = 96 * 1024;

template <typename T>
__device__ __forceinline__ const T* stage_weights(const T* __restrict__ w, T* smem, int dim, uint32_t n, bool use_smem) {
    if (!use_smem) return w;
    for (int i = threadIdx.x; i < dim * (int)n; i += blockDim.x) smem[i] = w[i];
    __syncthreads();
    return smem;
}

// prod_d w[d, i_d(p)], digits peeled from the least significant (last) dimension; no digit array.
template <typename T>
__device__ __forceinline__ T point_weight(const T* __restrict__ sw, const GridIndex& gi, uint64_t p) {
    T w = (T)1;
    if (p <= 0xffffffffull) {
        uint32_t q = (uint32_t)p;
        for (int d = gi.dim - 1; d >= 0; --d) {
            const uint32_t t = q / gi.n;
            w *= sw[d * gi.n + (q - t * gi.n)];
            q = t;
        }
    } else {
        for (int d = gi.dim - 1; d >= 0; --d) {
            const uint64_t t = p / gi.n;
            w *= sw[d * gi.n + (uint32_t)(p - t * gi.n)];
            p = t;
        }
    }
    return w;
}

template <typename T>
__global__ void __launch_bounds__(256)
point_weights_kernel(const T* __restrict__ w, uint32_t n, int dim, int64_t p_begin, int64_t p_end, T* __restrict__ out,
                     bool use_smem) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const T* sw = stage_weights<T>(w, reinterpret_cast<T*>(smem_raw), dim, n, use_smem);
    const GridIndex gi{n, dim};
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < p_end - p_begin;
         r += (int64_t)gridDim.x * blockDim.x)
        out[r] = point_weight<T>(sw, gi, (uint64_t)(p_begin + r));
}

// cols == 1 contraction: fp64 accumulation of f[p]*W[p], deterministic two-stage reduction.
template <typename T>
__global__ void __launch_bounds__(256)
contract1_kernel(const T* __restrict__ f, const T* __restrict__ w, uint32_t n, int dim, int64_t p_begin,
                 int64_t p_end, double* partials, unsigned int* ticket, double* out, bool use_smem) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double sh[32];
    const T* sw = stage_weights<T>(w, reinterpret_cast<T*>(smem_raw), dim, n, use_smem);
    const GridIndex gi{n, dim};
    double acc[1] = {0.0};
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < p_end - p_begin;
         r += (int64_t)gridDim.x * blockDim.x) {
        const T wt = point_weight<T>(sw, gi, (uint64_t)(p_begin + r));
        acc[0] += (double)__ldcs(&f[r]) * (double)wt;
    }
    grid_sum_finish<1>(acc, sh, partials, ticket, out);
}

// cols > 1: each thread owns one column (flat stride a multiple of cols), shared accumulators per CTA.
template <typename T>
__global__ void __launch_bounds__(256)
contractk_kernel(const T* __restrict__ f, const T* __restrict__ w, uint32_t n, int dim, int64_t p_begin,
                 int64_t p_end, int64_t cols, int64_t S, double* partials, unsigned int* ticket, double* out,
                 bool use_smem) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ bool is_last;
    double* sacc = reinterpret_cast<double*>(smem_raw);  // [cols]
    for (int64_t c = threadIdx.x; c < cols; c += blockDim.x) sacc[c] = 0.0;
    const T* sw = stage_weights<T>(w, reinterpret_cast<T*>(sacc + cols), dim, n, use_smem);  // [dim*n]
    const GridIndex gi{n, dim};
    const int64_t total = (p_end - p_begin) * cols;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid < S) {
        double acc = 0.0;
        for (int64_t e = tid; e < total; e += S) {
            const int64_t r = e / cols;
            acc += (double)f[e] * (double)point_weight<T>(sw, gi, (uint64_t)(p_begin + r));
        }
        atomicAdd(&sacc[tid % cols], acc);
    }
    __syncthreads();
    for (int64_t c = threadIdx.x; c < cols; c += blockDim.x) partials[(size_t)blockIdx.x * cols + c] = sacc[c];
    if (threadIdx.x == 0) {
        __threadfence();
        is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        for (int64_t c = threadIdx.x; c < cols; c += blockDim.x) {
            double a = 0.0;
            for (unsigned int b = 0; b < gridDim.x; ++b) a += __ldcg(&partials[(size_t)b * cols + c]);
            out[c] = a;
        }
        if (threadIdx.x == 0) *ticket = 0u;
    }
}

}  // namespace tq

using namespace tq;

static int check_grid_args(const char* who, int32_t n, int32_t dim, int64_t p_begin, int64_t p_end) {
    TQ_REQUIRE(n >= 1 && dim >= 1 && dim <= TQ_MAX_DIM, "%s: need n >= 1 and 1 <= dim <= %d", who, TQ_MAX_DIM);
    TQ_REQUIRE(p_begin >= 0 && p_end >= p_begin, "%s: bad point range", who);
    return TQ_OK;
}

extern "C" {

int tq_nc_grid_points(const void* nodes, int32_t n, int32_t dim, int64_t p_begin, int64_t p_end,
                      void* points, int32_t dtype, void* stream) {
    int rc = check_grid_args("tq_nc_grid_points", n, dim, p_begin, p_end);
    if (rc) return rc;
    if (p_end == p_begin) return TQ_OK;
    const int grid = grid_for((p_end - p_begin) * dim, 256, 8);
    TQ_DISPATCH_DTYPE(dtype, {
        grid_points_kernel<T><<<grid, 256, 0, as_stream(stream)>>>((const T*)nodes, (uint32_t)n, dim, p_begin, p_end, (T*)points);
    });
    return check_launch("grid_points_kernel");
}

int tq_nc_grid_points_backward(const void* grad_points, int32_t n, int32_t dim, int64_t p_begin,
                               int64_t p_end, double* grad_nodes_f64, int32_t dtype, void* stream) {
    int rc = check_grid_args("tq_nc_grid_points_backward", n, dim, p_begin, p_end);
    if (rc) return rc;
    cudaStream_t st = as_stream(stream);
    cudaMemsetAsync(grad_nodes_f64, 0, (size_t)dim * n * sizeof(double), st);
    if (p_end == p_begin) return TQ_OK;
    const int grid = grid_for((p_end - p_begin) * dim, 256, 2);
    const bool use_smem = (size_t)dim * n * sizeof(double) <= NC_TABLE_SMEM;
    const size_t smem = use_smem ? (size_t)dim * n * sizeof(double) : 0;
    TQ_DISPATCH_DTYPE(dtype, {
        cudaFuncSetAttribute(grid_points_backward_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)NC_TABLE_SMEM);
        grid_points_backward_kernel<T><<<grid, 256, smem, st>>>((const T*)grad_points, (uint32_t)n, dim, p_begin, p_end, grad_nodes_f64, use_smem);
    });
    return check_launch("grid_points_backward_kernel");
}

int tq_nc_point_weights(const void* w, int32_t n, int32_t dim, int64_t p_begin, int64_t p_end,
                        void* out, int32_t dtype, void* stream) {
    int rc = check_grid_args("tq_nc_point_weights", n, dim, p_begin, p_end);
    if (rc) return rc;
    if (p_end == p_begin) return TQ_OK;
    const int grid = grid_for(p_end - p_begin, 256, 8);
    TQ_DISPATCH_DTYPE(dtype, {
        const bool use_smem = (size_t)dim * n * sizeof(T) <= NC_TABLE_SMEM;
        const size_t smem = use_smem ? (size_t)dim * n * sizeof(T) : 0;
        cudaFuncSetAttribute(point_weights_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)NC_TABLE_SMEM);
        point_weights_kernel<T><<<grid, 256, smem, as_stream(stream)>>>((const T*)w, (uint32_t)n, dim, p_begin, p_end, (T*)out, use_smem);
    });
    return check_launch("point_weights_kernel");
}

int tq_nc_contract(const void* f, const void* w, int32_t n, int32_t dim, int64_t p_begin, int64_t p_end,
                   int64_t cols, int32_t dtype, double* out_f64, void* ws, size_t ws_bytes, void* stream) {
    int rc = check_grid_args("tq_nc_contract", n, dim, p_begin, p_end);
    if (rc) return rc;
    TQ_REQUIRE(cols >= 1 && cols <= 2048, "tq_nc_contract: 1 <= cols <= 2048 (got %lld)", (long long)cols);
    Workspace wk(ws, ws_bytes);
    unsigned int* ticket = wk.take<unsigned int>(WS_HEADER / sizeof(unsigned int));
    cudaStream_t st = as_stream(stream);
    const int64_t rows = p_end - p_begin;
    if (cols == 1) {
        const int grid = grid_for(rows, 256, 4);
        double* partials = wk.take<double>((size_t)grid);
        if (!ticket || !partials) { set_error("tq_nc_contract: workspace too small"); return TQ_ERR_WORKSPACE; }
        TQ_DISPATCH_DTYPE(dtype, {
            const bool use_smem = (size_t)dim * n * sizeof(T) <= NC_TABLE_SMEM;
            const size_t smem = use_smem ? (size_t)dim * n * sizeof(T) : 0;
            cudaFuncSetAttribute(contract1_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)NC_TABLE_SMEM);
            contract1_kernel<T><<<grid, 256, smem, st>>>((const T*)f, (const T*)w, (uint32_t)n, dim, p_begin, p_end, partials, ticket, out_f64, use_smem);
        });
        return check_launch("contract1_kernel");
    }
    int grid = grid_for(rows * cols, 256, 2);
    int64_t threads = (int64_t)grid * 256;
    if (threads < cols) { grid = (int)((cols + 255) / 256); threads = (int64_t)grid * 256; }
    const int64_t S = (threads / cols) * cols;
    double* partials = wk.take<double>((size_t)grid * cols);
    if (!ticket || !partials) { set_error("tq_nc_contract: workspace too small"); return TQ_ERR_WORKSPACE; }
    TQ_DISPATCH_DTYPE(dtype, {
        const bool use_smem = (size_t)dim * n * sizeof(T) <= NC_TABLE_SMEM;
        const size_t smem = (size_t)cols * sizeof(double) + (use_smem ? (size_t)dim * n * sizeof(T) : 0);
        cudaFuncSetAttribute(contractk_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(NC_TABLE_SMEM + 2048 * sizeof(double)));
        contractk_kernel<T><<<grid, 256, smem, st>>>((const T*)f, (const T*)w, (uint32_t)n, dim, p_begin, p_end, cols, S, partials, ticket, out_f64, use_smem);
    });
    return check_launch("contractk_kernel");
}

}  // extern "C"
