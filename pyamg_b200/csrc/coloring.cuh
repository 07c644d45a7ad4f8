// coloring.cuh -- the reference's 'MIS' vertex colouring on the device.
//
// pyamg.graph.vertex_coloring(G, 'MIS') = amg_core vertex_coloring_mis (graph.h:218-235): colour k is the
// lexicographically first maximal independent set (maximal_independent_set_serial, graph.h:128-199, walks the
// vertices in index order) of what colours 0..k-1 left over.  Vertex i misses set k exactly when a SMALLER-index
// neighbour is in it, hence
//        colour(i) = the smallest colour held by no neighbour j < i            (natural-order first fit),
// a pure function of the final colours of i's smaller-index neighbours.  That recurrence is evaluated here the way
// the engine evaluates the reference's lexicographic Gauss-Seidel: as a wavefront.  Every round each still
// uncoloured vertex looks at its smaller-index neighbours; once all of them hold a colour it takes its own.  A colour
// is written once and never changes, so reading a neighbour that gets coloured in the same round is harmless
// (the vertex merely finishes a round earlier): the result is THE colouring of the sequential reference, vertex by
// vertex, whatever the thread interleaving.  Rounds needed = longest chain i1 < i2 < ... of adjacent vertices
// (~3 n^(1/3) on a 7-point lattice), each a streaming pass over the uncoloured rows.
#pragma once
#include "csr_kernels.cuh"

namespace amgb {

constexpr int kMaxColors = 256;

__device__ __forceinline__ int first_zero_bit(unsigned long long m)       // m != all ones
{
    const unsigned lo = (unsigned)~m, hi = (unsigned)(~m >> 32);
    return lo ? __ffs((int)lo) - 1 : 32 + __ffs((int)hi) - 1;
}

// color[i] = -1: uncoloured.  *remaining counts the vertices this round left uncoloured; *overflow is set when a
// vertex needs more than kMaxColors colours.
__global__ void __launch_bounds__(256) mis_color_round_kernel(int n, const int *__restrict__ Ap,
                                                              const int *__restrict__ Aj, int *color,
                                                              unsigned long long *remaining, int *overflow)
{
    unsigned long long left = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        if (__ldcg(color + i) >= 0) continue;
        unsigned long long used[kMaxColors / 64] = {};
        bool ready = true;
        const int e = __ldg(Ap + i + 1);
        for (int jj = __ldg(Ap + i); jj < e; jj++) {
            const int j = __ldg(Aj + jj);
            if (j >= i) continue;                       // larger-index neighbours (and the diagonal) do not matter
            const int c = __ldcg(color + j);            // colours are written by other SMs: read them from L2
            if (c < 0) { ready = false; break; }
            used[c >> 6] |= 1ull << (c & 63);
        }
        if (!ready) { left++; continue; }
        int c = -1;
#pragma unroll
        for (int w = 0; w < kMaxColors / 64; w++)
            if (c < 0 && ~used[w] != 0ull) c = w * 64 + first_zero_bit(used[w]);
        if (c < 0) { *overflow = 1; c = kMaxColors - 1; }
        color[i] = c;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) left += __shfl_xor_sync(0xffffffffu, left, o);
    if ((threadIdx.x & 31) == 0 && left) atomicAdd(remaining, left);
}

}  // namespace amgb
