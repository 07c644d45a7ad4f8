// grid_kernel.cuh -- the coarse part of the multigrid cycle as ONE persistent, cooperatively launched kernel
// that spans every SM ("cycle interpreter", grid-wide edition of tail_kernel.cuh).
//
// Why: levels 3..9 of the BASELINE configs[2] hierarchy (176 750 ... 5 rows, 27 / 60 / 43 / 29 colours) cost
// ~640 dependent Gauss-Seidel wave launches per V-cycle, each 0.1 - 5 MB of L2-resident operator: ~8 us per
// launch inside a CUDA graph (launch + drain + three dependent L2 round trips), ~5 ms of a 9.9 ms cycle at
// 0.00 - 0.08 of the HBM roofline ("what's weak" 3).  A launch boundary is the most expensive
// barrier the machine offers; here the host records the launch sequence of MultilevelSolver.__solve
// (multilevel.py:584-662) for all levels >= tail_level ONCE into a step list (the TailStep records of
// tail_kernel.cuh), and one CTA per SM walks it:
//   * a step = one former launch (a GS wave, a residual, a restriction, a prolongation, the dense coarse solve);
//     its rows are dealt warp by warp ROUND-ROBIN OVER THE SMs, so even a 600-row wave has 148 load/store units
//     and 148 L2 request streams working on its latency chain;
//   * steps are separated by a grid barrier: one release-add per CTA on a monotone 64-bit counter in L2 and an
//     acquire spin on it (~1 us, against ~8 us for a launch boundary);
//   * steps too small to be worth even that (<= tail_solo_bytes) run on CTA 0 alone, separated by __syncthreads;
//   * vectors are rewritten between steps by other SMs, so every vector load of a grid-wide step goes to L2
//     (ld.global.cg); operator arrays are immutable and use the read-only path.
// The thread-block-cluster variant (tail_kernel) has a cheaper barrier (barrier.cluster, ~0.2 us) but only 16 SMs:
// it lost against per-wave launches on the 5-11 M-entry levels, which is what this kernel is for.
//
// Deadlock safety: launched with cudaLaunchAttributeCooperative (the driver refuses the launch unless all CTAs are
// co-resident); the spin traps after ~2^27 polls instead of hanging the device.
#pragma once
#include "tail_kernel.cuh"

namespace amgb {

// one 128-byte line of its own: `count` is the barrier counter (monotone across launches), `base` its value at
// the start of the running launch (rewritten by CTA 0 at the very end of every launch)
struct GridSync {
    unsigned long long count;
    unsigned long long base;
    unsigned long long pad[14];
};

constexpr int kGridThreads = 1024;

__device__ __forceinline__ unsigned long long grid_ld_acquire(const unsigned long long *p)
{
#ifdef AMGB_EMU
    return *p;
#else
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
#endif
}

// every thread of every CTA calls it; `target` = counter value once all CTAs have arrived
__device__ __forceinline__ void grid_barrier(GridSync *gs, unsigned long long target)
{
    __syncthreads();                          // this CTA's writes of the step are complete ...
    if (threadIdx.x == 0) {
#ifdef AMGB_EMU
        ::emu::grid_arrive_wait(&gs->count, target);
#else
        __threadfence();                      // ... and ordered before the arrival (release, cumulative over bar.sync)
        atomicAdd(&gs->count, 1ull);
        unsigned long long spins = 0;
        while (grid_ld_acquire(&gs->count) < target) {
            if (++spins > 64) __nanosleep(100);               // long waits (solo phases of CTA 0): stop hammering L2
            if (spins > (1ull << 27)) __trap();               // a lost CTA must not hang the device
        }
        __threadfence();                      // acquire side for the whole CTA (invalidates this SM's L1)
#endif
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kGridThreads, 1)
coarse_grid_kernel(const TailStep *__restrict__ steps, int nsteps, GridSync *gs)
{
    const int nctas = (int)gridDim.x;
    const int nthreads = nctas * kGridThreads;
    // virtual thread id: warp w of CTA c is warp (w * nctas + c) of the grid, so consecutive 32-lane chunks of a
    // step's rows go to different SMs
    const int vtid = (((int)threadIdx.x >> 5) * nctas + (int)blockIdx.x) * 32 + ((int)threadIdx.x & 31);
    unsigned long long target = __ldcg(&gs->base);      // nobody rewrites it before the final barrier below
    bool in_solo = false;
    TailStep next = steps[0];
    for (int s = 0; s < nsteps; s++) {
        const TailStep st = next;
        if (s + 1 < nsteps) next = steps[s + 1];     // the step list is immutable: fetched ahead of the barrier
        if (st.solo) {
            in_solo = true;
            if (blockIdx.x == 0) {
                tail_step<true>(st, (int)threadIdx.x, kGridThreads);
                __syncthreads();
            }
        } else {
            if (in_solo) {                            // publish CTA 0's solo results
                target += (unsigned long long)nctas;
                grid_barrier(gs, target);
                in_solo = false;
            }
            tail_step<false>(st, vtid, nthreads);
            target += (unsigned long long)nctas;
            grid_barrier(gs, target);
        }
    }
    // every CTA has read `base` long before it arrives here: CTA 0 may now publish the next launch's base
    target += (unsigned long long)nctas;
    grid_barrier(gs, target);
    if (blockIdx.x == 0 && threadIdx.x == 0) gs->base = target;
}

}  // namespace amgb
