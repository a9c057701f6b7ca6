// Row -> cube lookup for CTA tiles of consecutive rows of a cube-sorted VEGAS iteration.
// Shared by the stratified sampling kernel (vegas_strat.cu) and the fused VEGAS kernel (fused.cu).
#pragma once
#include "common.cuh"

namespace tq {

constexpr int ST_ROWS = 256;                // rows per CTA tile
constexpr int ST_SLICE = ST_ROWS / 2 + 4;   // cubes a tile can overlap (nh >= 2) + slack for an early start

// Double-buffered shared state of the walk: one barrier per tile.
struct StratTile {
    long long off[2][ST_SLICE];        // offsets of the cubes c_lo .. c_lo + ST_SLICE - 1
    unsigned char cube[2][ST_ROWS];    // slice index of every row of the tile
    long long first;
};

// largest c in [0, n_cubes) with offsets[c] <= row
__device__ __forceinline__ long long cube_of_row(const long long* __restrict__ offsets, int64_t n_cubes, long long row) {
    long long lo = 0, hi = n_cubes - 1;
    while (lo < hi) {
        const long long mid = (lo + hi + 1) >> 1;
        if (__ldg(&offsets[mid]) <= row) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// Called by all threads of the CTA (blockDim.x >= ST_SLICE).  One thread per overlapped cube reads its two
// offsets and writes its slice index into cube[buf][row - rb] for the rows it owns in [rb, re); ends with the
// tile's only barrier.  c_lo must satisfy offsets[c_lo] <= rb with the cube of row rb at most one above it.
__device__ __forceinline__ void strat_tile_fill(StratTile& st, int buf, const long long* __restrict__ offsets,
                                                int64_t n_cubes, long long c_lo, int64_t rb, int64_t re) {
    if (threadIdx.x < ST_SLICE) {
        const long long c = c_lo + threadIdx.x;
        const long long lo = __ldg(&offsets[c < n_cubes ? c : n_cubes]);
        const long long hi = __ldg(&offsets[c + 1 < n_cubes ? c + 1 : n_cubes]);
        st.off[buf][threadIdx.x] = lo;
        const long long a = lo > rb ? lo : rb, b = hi < re ? hi : re;
        for (long long r = a; r < b; ++r) st.cube[buf][r - rb] = (unsigned char)threadIdx.x;
    }
    __syncthreads();
}

}  // namespace tq
