// tail_kernel.cuh -- the coarse part of the multigrid cycle as ONE persistent thread-block-cluster
// kernel ("cycle interpreter").
//
// Below a few million stored entries a level is no longer bandwidth- but latency-bound: a Gauss-
// Seidel sweep of a 40 k-row operator with 60 colours is 60 dependent launches of a few microseconds
// each, and the RS hierarchy of BASELINE configs[2] spends ~900 of its ~1000 launches per V-cycle
// there.  Here the host records the launch sequence of MultilevelSolver.__solve (multilevel.py:584-662)
// for all levels >= tail_level ONCE into a step list in HBM, and a single cluster of up to 16 CTAs
// (sm_100a thread-block clusters: co-scheduled on one GPC, hardware barrier.cluster ~0.2 us) walks the
// list: every step is one former kernel launch (a GS wave, a residual, a restriction, ...), steps are
// separated by a cluster barrier with release/acquire semantics.  All of these operators live in L2.
//
// Memory model: vectors are rewritten between steps by other CTAs of the cluster, so every vector
// load goes to L2 (ld.global.cg); operator arrays and row lists are immutable and use the normal
// read-only path.
#pragma once
#include "csr_kernels.cuh"

namespace amgb {

enum TailOp { T_SPMV = 0, T_RESID = 1, T_PADD = 2, T_JACOBI = 3, T_GS = 4, T_FILL = 5, T_DENSE = 6, T_COPY = 7 };

struct TailStep {
    int op;
    int G;                 // lanes per row (power of two <= 32)
    int row0, nrows;       // rows row0 .. row0+nrows-1, or rows[0..nrows) when rows != nullptr
    const int *rows;
    const int *Ap;
    const int *Aj;
    const double *Ax;      // T_DENSE: the dense matrix (nrows x ncols, row-major)
    const double *x;       // gathered vector / copy source
    const double *b;
    double *y;             // output (see csr_kernels.cuh epilogues); T_FILL/T_COPY target
    double omega;
    int ncols;
    int solo;              // 1: small step, executed by CTA 0 alone (__syncthreads, L1-cached loads)
};

constexpr int kTailThreads = 1024;

__device__ __forceinline__ void cluster_barrier()
{
#ifdef AMGB_EMU
    ::emu::cluster_barrier();
#else
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
#endif
}
__device__ __forceinline__ unsigned cluster_rank()
{
#ifdef AMGB_EMU
    return ::emu::cluster_ctarank();
#else
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
#endif
}
__device__ __forceinline__ unsigned cluster_size()
{
#ifdef AMGB_EMU
    return ::emu::cluster_nctarank();
#else
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
#endif
}

// SOLO = true: the step runs inside one CTA; vectors may be read through L1 (same-SM coherent).
template <bool SOLO>
__device__ __forceinline__ double ld_vec(const double *p)
{
    return SOLO ? *p : __ldcg(p);
}

template <int G, bool SOLO>
__device__ __forceinline__ void tail_rows(const TailStep &st, int tid, int nthreads)
{
    const int lane = tid & (G - 1);
    const int group = tid / G, ngroups = nthreads / G;
    const bool need_diag = (st.op == T_JACOBI || st.op == T_GS);
    for (int base = 0; base < st.nrows; base += ngroups) {     // uniform trip count: shuffles stay converged
        const int k = base + group;
        const bool active = k < st.nrows;
        int row = 0, start = 0, end = 0;
        if (active) {
            row = st.rows ? st.rows[k] : st.row0 + k;
            start = st.Ap[row];
            end = st.Ap[row + 1];
        }
        double sum = 0.0, diag = 0.0;
        int jd = -1;
        // latency-bound: issue a chunk of U independent (col,val) loads, then U independent gathers
        constexpr int U = 8;
        for (int j0 = start + lane; j0 < end; j0 += G * U) {
            int c[U];
            double v[U], xv[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int jj = j0 + u * G;
                const bool ok = jj < end;
                c[u] = ok ? __ldg(st.Aj + jj) : -1;
                v[u] = ok ? __ldg(st.Ax + jj) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
                const bool isd = need_diag && c[u] == row;
                xv[u] = (c[u] >= 0 && !isd) ? ld_vec<SOLO>(st.x + c[u]) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
                if (need_diag && c[u] == row && c[u] >= 0) { diag = v[u]; jd = j0 + u * G; }
                else sum += v[u] * xv[u];
            }
        }
        sum = group_sum<G>(sum);
        if (need_diag) {
#pragma unroll
            for (int o = G / 2; o > 0; o >>= 1) {
                const int jo = __shfl_xor_sync(0xffffffffu, jd, o, G);
                const double dv = __shfl_xor_sync(0xffffffffu, diag, o, G);
                if (jo > jd) { jd = jo; diag = dv; }
            }
        }
        if (active && lane == 0) {
            switch (st.op) {
            case T_SPMV: st.y[row] = sum; break;
            case T_RESID: st.y[row] = ld_vec<SOLO>(st.b + row) - sum; break;
            case T_PADD: st.y[row] = ld_vec<SOLO>(st.y + row) + sum; break;
            case T_JACOBI: {
                const double xi = ld_vec<SOLO>(st.x + row), bi = ld_vec<SOLO>(st.b + row);
                st.y[row] = (diag != 0.0) ? (1.0 - st.omega) * xi + st.omega * ((bi - sum) / diag) : xi;
                break;
            }
            default:   // T_GS
                if (diag != 0.0) {
                    const double g = (ld_vec<SOLO>(st.b + row) - sum) / diag;
                    st.y[row] = (st.omega == 1.0) ? g : st.omega * g + (1.0 - st.omega) * ld_vec<SOLO>(st.y + row);
                }
            }
        }
    }
}

template <bool SOLO>
__device__ __forceinline__ void tail_step(const TailStep &st, int tid, int nthreads)
{
    if (st.op <= T_GS) {
        switch (st.G) {
        case 1: tail_rows<1, SOLO>(st, tid, nthreads); break;
        case 2: tail_rows<2, SOLO>(st, tid, nthreads); break;
        case 4: tail_rows<4, SOLO>(st, tid, nthreads); break;
        case 8: tail_rows<8, SOLO>(st, tid, nthreads); break;
        case 16: tail_rows<16, SOLO>(st, tid, nthreads); break;
        default: tail_rows<32, SOLO>(st, tid, nthreads); break;
        }
    } else if (st.op == T_FILL) {
        for (int i = tid; i < st.nrows; i += nthreads) st.y[i] = 0.0;
    } else if (st.op == T_COPY) {
        for (int i = tid; i < st.nrows; i += nthreads) st.y[i] = ld_vec<SOLO>(st.x + i);
    } else {   // T_DENSE: y = M x, one warp per row (coarsest-level pseudo-inverse)
        const int w = tid >> 5, l = tid & 31, nw = nthreads >> 5;
        for (int base = 0; base < st.nrows; base += nw) {
            const int r = base + w;
            double acc = 0.0;
            if (r < st.nrows)
                for (int j = l; j < st.ncols; j += 32)
                    acc += __ldg(st.Ax + (size_t)r * st.ncols + j) * ld_vec<SOLO>(st.x + j);
            acc = group_sum<32>(acc);
            if (r < st.nrows && l == 0) st.y[r] = acc;
        }
    }
}

// Steps flagged `solo` are too small to be worth a cluster-wide barrier: CTA 0 runs them alone, separated
// by __syncthreads, reading vectors through its own L1; the other CTAs wait at the next cluster barrier.
__global__ void __launch_bounds__(kTailThreads) tail_kernel(const TailStep *__restrict__ steps, int nsteps)
{
    const int rank = (int)cluster_rank();
    const int nthreads = (int)cluster_size() * kTailThreads;
    const int tid = rank * kTailThreads + threadIdx.x;
    bool in_solo = false;
    TailStep next = steps[0];
    for (int s = 0; s < nsteps; s++) {
        const TailStep st = next;
        if (s + 1 < nsteps) next = steps[s + 1];     // the step list is immutable: fetch ahead of the barrier
        if (st.solo) {
            in_solo = true;
            if (rank == 0) {
                tail_step<true>(st, threadIdx.x, kTailThreads);
                __syncthreads();
            }
        } else {
            if (in_solo) { cluster_barrier(); in_solo = false; }   // publish CTA 0's solo results
            tail_step<false>(st, tid, nthreads);
            cluster_barrier();
        }
    }
}

}  // namespace amgb
