// resident_kernel.cuh -- EXPERIMENTAL (opt-in with AMGB_RESIDENT=1; written at the end of round 1 without
// GPU time left to validate it -- the default path never launches it).
//
// A whole Gauss-Seidel smoother application (all waves of all sweeps) on a small level as ONE thread-block-
// cluster kernel with the iterate RESIDENT in distributed shared memory:
//   * the level's x is split in power-of-two slices over the cluster's CTAs (CTA c owns x[c*m, (c+1)*m));
//   * a wave's rows are spread over all threads of the cluster; gathers x[col] and the update of x[row] go
//     through DSMEM (mapa + ld/st.shared::cluster, ~200 cycles) instead of an L2 round trip;
//   * waves are separated by barrier.cluster (release/acquire), not by kernel launches;
//   * the operator rows are immutable and still come from L2/HBM, in chunks of independent loads.
// Motivation: levels 3-9 of the 256^3 hierarchy spend ~5 ms per cycle in ~700 dependent waves
// of 6-9 us each (launch + three dependent global round trips); here a wave costs one barrier plus one
// operator round trip.  Semantics identical to OP_GS of csr_kernels.cuh (relaxation.h:48-76 / :116-145 / :736-768).
#pragma once
#include "csr_kernels.cuh"
#include "tail_kernel.cuh"   // cluster_barrier / cluster_rank / cluster_size

namespace amgb {

struct ResidentArgs {
    int n;                   // rows of the level
    int log2m;               // slice size m = 1 << log2m (doubles per CTA)
    const int *Ap;
    const int *Aj;
    const double *Ax;
    double *x;               // in/out, level numbering (wave-major)
    const double *b;
    double omega;
    const long long *wave_ptr;   // device: rows of wave w = [wave_ptr[w], wave_ptr[w+1])
    const int *seq;              // device: wave indices in execution order (forward / backward / symmetric x iterations)
    int seq_len;
    int G;                       // lanes per row
};

constexpr int kResidentThreads = 1024;
constexpr int kResidentMaxLog2m = 14;      // 16384 doubles = 128 KB of shared memory per CTA

__device__ __forceinline__ unsigned dsmem_addr(const double *local, unsigned cta)
{
#ifdef AMGB_EMU
    return ::emu::dsmem_addr(local, cta);
#else
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"((unsigned)__cvta_generic_to_shared(local)), "r"(cta));
    return r;
#endif
}
__device__ __forceinline__ double dsmem_ld(unsigned addr)
{
#ifdef AMGB_EMU
    return *::emu::dsmem_ptr(addr);
#else
    double v;
    asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
    return v;
#endif
}
__device__ __forceinline__ void dsmem_st(unsigned addr, double v)
{
#ifdef AMGB_EMU
    *::emu::dsmem_ptr(addr) = v;
#else
    asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
#endif
}

template <int G>
__device__ __forceinline__ void resident_wave(const ResidentArgs &a, double *xs, int r0, int nr, int tid, int nthreads)
{
    const int lane = tid & (G - 1);
    const int group = tid / G, ngroups = nthreads / G;
    const int mask = (1 << a.log2m) - 1;
    for (int base = 0; base < nr; base += ngroups) {          // uniform trip count across the cluster
        const int k = base + group;
        const bool active = k < nr;
        int row = 0, start = 0, end = 0;
        if (active) {
            row = r0 + k;
            start = __ldg(a.Ap + row);
            end = __ldg(a.Ap + row + 1);
        }
        double sum = 0.0, diag = 0.0;
        int jd = -1;
        constexpr int U = 8;
        for (int j0 = start + lane; j0 < end; j0 += G * U) {
            int c[U];
            double v[U], xv[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int jj = j0 + u * G;
                const bool ok = jj < end;
                c[u] = ok ? __ldg(a.Aj + jj) : -1;
                v[u] = ok ? __ldg(a.Ax + jj) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
                const bool skip = c[u] < 0 || c[u] == row;
                xv[u] = skip ? 0.0 : dsmem_ld(dsmem_addr(xs + (c[u] & mask), (unsigned)(c[u] >> a.log2m)));
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
                if (c[u] == row && c[u] >= 0) { diag = v[u]; jd = j0 + u * G; }   // last stored duplicate wins
                else sum += v[u] * xv[u];
            }
        }
        sum = group_sum<G>(sum);
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) {
            const int jo = __shfl_xor_sync(0xffffffffu, jd, o, G);
            const double dv = __shfl_xor_sync(0xffffffffu, diag, o, G);
            if (jo > jd) { jd = jo; diag = dv; }
        }
        if (active && lane == 0 && diag != 0.0) {
            const unsigned addr = dsmem_addr(xs + (row & mask), (unsigned)(row >> a.log2m));
            const double g = (__ldg(a.b + row) - sum) / diag;
            dsmem_st(addr, (a.omega == 1.0) ? g : a.omega * g + (1.0 - a.omega) * dsmem_ld(addr));
        }
    }
}

__global__ void __launch_bounds__(kResidentThreads) resident_gs_kernel(const ResidentArgs a)
{
    extern __shared__ __align__(16) unsigned char res_smem[];
    double *xs = reinterpret_cast<double *>(res_smem);
    const int rank = (int)cluster_rank();
    const int nthreads = (int)cluster_size() * kResidentThreads;
    const int tid = rank * kResidentThreads + threadIdx.x;
    const int m = 1 << a.log2m;
    for (int i = threadIdx.x; i < m; i += kResidentThreads) {          // this CTA's slice of x -> shared memory
        const long long g = (long long)rank * m + i;
        xs[i] = g < a.n ? a.x[g] : 0.0;
    }
    cluster_barrier();
    for (int s = 0; s < a.seq_len; s++) {
        const int w = __ldg(a.seq + s);
        const int r0 = (int)__ldg(a.wave_ptr + w);
        const int nr = (int)__ldg(a.wave_ptr + w + 1) - r0;
        switch (a.G) {
        case 1: resident_wave<1>(a, xs, r0, nr, tid, nthreads); break;
        case 2: resident_wave<2>(a, xs, r0, nr, tid, nthreads); break;
        case 4: resident_wave<4>(a, xs, r0, nr, tid, nthreads); break;
        case 8: resident_wave<8>(a, xs, r0, nr, tid, nthreads); break;
        case 16: resident_wave<16>(a, xs, r0, nr, tid, nthreads); break;
        default: resident_wave<32>(a, xs, r0, nr, tid, nthreads); break;
        }
        cluster_barrier();                                              // publish the wave's updates
    }
    for (int i = threadIdx.x; i < m; i += kResidentThreads) {
        const long long g = (long long)rank * m + i;
        if (g < a.n) a.x[g] = xs[i];
    }
}

}  // namespace amgb
