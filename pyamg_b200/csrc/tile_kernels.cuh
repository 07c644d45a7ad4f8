// tile_kernels.cuh -- the sm_100a streaming kernel of the V-cycle: CSR row segments staged in shared
// memory by the TMA (1-D bulk copies, cp.async.bulk -> SASS UBLKCP) with mbarrier completion.
//
// Why: the hot path is HBM-bound CSR traffic (12 B per stored entry, 0.17 flop/B).  A lanes-per-row
// kernel (csr_kernels.cuh) keeps only a few hundred bytes per warp in flight and pays a dependent
// row-pointer -> entries -> gather latency chain per row group.  Here every warp owns a ring of
// shared-memory stages; one elected lane asks the TMA for the NEXT tile's contiguous segments
// (values, column indices, row pointers, and the b / x_old slices the epilogue needs -- each a single
// 16-B-aligned bulk copy) while the warp reduces the CURRENT tile out of shared memory.  The copy
// engine, not the LSU, streams the operator; the warps only issue the x gathers and the fused
// epilogue store.  Bytes in flight per SM = warps x stage size, independent of register pressure.
//
// Tiles are built on the host at upload (engine.cu build_tiles): whole rows, at most T stored entries
// and RMAX rows, never crossing a Gauss-Seidel wave boundary; a row longer than T is a tile of its own
// and is streamed straight from global memory.  Work distribution is static (tile t -> warp t mod
// #warps), the grid persistent: 148 SMs x resident CTAs.
//
// L2 policy: the operator stream is read exactly once per launch -> evict-first; the gathered vector
// is re-read by every row that references it (and, for multi-colour Gauss-Seidel, by every wave) ->
// evict-last, so it survives in the 126 MB L2 while the operator streams past it.
//
// Epilogues (template OP, same semantics as csr_kernels.cuh / the reference, file:line there):
//   OP_SPMV, OP_RESID (+|r|^2), OP_PADD, OP_JACOBI (+ optional residual by-product), OP_GS (rows of one
//   wave, contiguous thanks to the wave-major row permutation applied at upload).
#pragma once
#include "csr_kernels.cuh"

namespace amgb {

struct TileDesc { int row0; int nz0; };   // tile t = rows [row0[t], row0[t+1]), entries [nz0[t], nz0[t+1])

struct TileArgs {
    const TileDesc *tiles;   // n_tiles + 1 descriptors (sentinel last)
    int tile_begin, tile_end;
    const int *Ap;
    const int *Aj;
    const double *Ax;
    const double *x;
    const double *b;
    double *y;
    double *r;
    double omega;
    double *partials;        // one per CTA (grid is fixed), or nullptr
    int hints;               // 1: L2 eviction hints (operator stream evict-first, x gathers evict-last)
};

// Tile geometry (compile-time): T stored entries and RMAX rows per tile, ring depth, warps per CTA.
// Shared memory per CTA = WARPS * STAGES * (12 T + 20 RMAX + ~150) bytes; what is left of the
// 228 KB/SM is L1 for the x gathers, so smaller tiles trade TMA efficiency for gather hit rate.
template <int T_, int RMAX_, int STAGES_, int WARPS_>
struct TileCfg {
    static constexpr int T = T_, RMAX = RMAX_, STAGES = STAGES_, WARPS = WARPS_;
};
using TileCfg0 = TileCfg<512, 128, 2, 8>;
using TileCfg1 = TileCfg<256, 64, 2, 8>;
using TileCfg4 = TileCfg<256, 64, 1, 8>;     // single stage, ~48 warps/SM: cross-warp overlap only
using TileCfg6 = TileCfg<224, 64, 1, 8>;     // 32 seven-point rows per tile; 7-8 CTAs/SM = 56-64 warps/SM
using TileCfg7 = TileCfg<352, 64, 1, 8>;     // denser coarse operators (~19-36 entries per row): 16 rows of 19 entries
                                             // fill both passes of a G = 2 warp (224 entries hold 11 such rows: 69 % of
                                             // the lanes, ncu profiles/r02_ncu_kernels.csv: 17.5 threads per instruction)

// the b / x_old slices are staged only for the epilogues that read them: shared memory is what limits
// the number of resident warps, and resident warps are what hides the gather latency
template <class C, int OP>
struct __align__(16) TileStageT {
    static constexpr bool kB = (OP == OP_RESID || OP == OP_JACOBI || OP == OP_GS);
    static constexpr bool kX = (OP == OP_JACOBI);
    double val[C::T + 8];
    double bseg[kB ? C::RMAX + 4 : 2];    // b[row0..row1)
    double xseg[kX ? C::RMAX + 4 : 2];    // x[row0..row1), old iterate
    int col[C::T + 8];
    int ptr[C::RMAX + 8];
};
template <class C, int OP>
struct __align__(16) TileWarpSmemT {
    TileStageT<C, OP> st[C::STAGES];
    unsigned long long bar[C::STAGES];
    unsigned long long pad_;
};
template <class C, int OP>
constexpr size_t tile_smem_bytes() { return sizeof(TileWarpSmemT<C, OP>) * C::WARPS; }

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
#ifdef AMGB_EMU
    ::emu::mbar_init(bar, count);
#else
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
#endif
}
__device__ __forceinline__ void fence_mbar_init()
{
#ifndef AMGB_EMU
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#endif
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
#ifdef AMGB_EMU
    ::emu::mbar_arrive_expect_tx(bar, bytes);
#else
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
#endif
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
#ifdef AMGB_EMU
    ::emu::mbar_wait(bar, parity);
#else
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
#endif
}
__device__ __forceinline__ unsigned long long policy_evict_first()
{
#ifdef AMGB_EMU
    return 0ull;
#else
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
#endif
}
__device__ __forceinline__ unsigned long long policy_evict_last()
{
#ifdef AMGB_EMU
    return 0ull;
#else
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
#endif
}
// 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier (TMA, SASS UBLKCP)
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar)
{
#ifdef AMGB_EMU
    ::emu::bulk_g2s(dst, src, bytes, bar);
#else
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
#endif
}
__device__ __forceinline__ void bulk_g2s_hint(void *dst, const void *src, unsigned bytes, unsigned long long *bar,
                                              unsigned long long policy)
{
#ifdef AMGB_EMU
    (void)policy;
    ::emu::bulk_g2s(dst, src, bytes, bar);
#else
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::
            "r"(smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
#endif
}
__device__ __forceinline__ double ld_x_hint_nc(const double *p, unsigned long long policy)
{
#ifdef AMGB_EMU
    (void)policy;
    return *p;
#else
    double v;
    asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(policy));
    return v;
#endif
}
__device__ __forceinline__ double ld_x_hint(const double *p, unsigned long long policy)
{
#ifdef AMGB_EMU
    (void)policy;
    return *p;
#else
    double v;
    asm volatile("ld.global.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(policy) : "memory");
    return v;
#endif
}

// One lane: ask the TMA for tile t's segments.  Every source starts on a 16-byte boundary: entry ranges
// are widened to multiples of 4 entries, vector ranges to multiples of 2 (the arrays are padded at upload).
// PART: bit 0 = arm the barrier for the whole tile and copy the immutable operator segments, bit 1 = copy the
// vector segments (b / x_old).  The split lets a programmatically-launched kernel (PDL) fetch its first
// operator tile while the previous launch is still writing the vectors.
template <class C, int OP, int PART = 3>
__device__ __forceinline__ void tile_issue(const TileArgs &a, int t, TileStageT<C, OP> &st, unsigned long long *bar,
                                           unsigned long long pol)
{
    const TileDesc d0 = a.tiles[t], d1 = a.tiles[t + 1];
    const int len = d1.nz0 - d0.nz0;
    if (len > C::T) {              // long row: streamed from global memory, nothing to stage
        if (PART & 1) mbar_expect_tx(bar, 0);
        return;
    }
    constexpr bool kB = (OP == OP_RESID || OP == OP_JACOBI || OP == OP_GS);
    constexpr bool kX = (OP == OP_JACOBI);
    const int s4 = d0.nz0 & ~3;
    const int cnt = ((d1.nz0 + 3) & ~3) - s4;
    const int r4 = d0.row0 & ~3;
    const int rcnt = ((d1.row0 + 1 + 3) & ~3) - r4;
    const int v2 = d0.row0 & ~1;
    const int vcnt = ((d1.row0 + 1) & ~1) - v2;
    if (PART & 1) {
        mbar_expect_tx(bar, (unsigned)(cnt * 12 + rcnt * 4 + ((kB ? 1 : 0) + (kX ? 1 : 0)) * vcnt * 8));
        if (cnt > 0) {
            if (a.hints) {
                bulk_g2s_hint(st.val, a.Ax + s4, (unsigned)cnt * 8u, bar, pol);
                bulk_g2s_hint(st.col, a.Aj + s4, (unsigned)cnt * 4u, bar, pol);
            } else {
                bulk_g2s(st.val, a.Ax + s4, (unsigned)cnt * 8u, bar);
                bulk_g2s(st.col, a.Aj + s4, (unsigned)cnt * 4u, bar);
            }
        }
        bulk_g2s(st.ptr, a.Ap + r4, (unsigned)rcnt * 4u, bar);
    }
    if (PART & 2) {
        if (kB && vcnt > 0) bulk_g2s(st.bseg, a.b + v2, (unsigned)vcnt * 8u, bar);
        if (kX && vcnt > 0) bulk_g2s(st.xseg, a.x + v2, (unsigned)vcnt * 8u, bar);
    }
}

// G lanes per row inside a tile (G = 1: thread per row -- 5/7-point stencils; larger G for the
// denser coarse operators).  32/G rows are reduced per pass.
// PDL = true (opt-in AMGB_TILE_PDL=1, launched with the programmatic-stream-serialization attribute): the
// kernel may start while its predecessor still runs; it arms its barriers and fetches the operator segments
// of its first tile(s), then waits for the predecessor before any vector (x, b, y, r) is touched.
template <int G, int OP, class C, bool PDL = false>
__global__ void __launch_bounds__(C::WARPS * 32) csr_tile_kernel(const TileArgs a)
{
    if (PDL) pdl_launch_dependents();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr bool kNeedDiag = (OP == OP_JACOBI || OP == OP_GS);
    constexpr int RPP = 32 / G;                       // rows per pass
    constexpr int STAGES = C::STAGES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sub = lane & (G - 1), grp = lane / G;
    TileWarpSmemT<C, OP> &ws = reinterpret_cast<TileWarpSmemT<C, OP> *>(smem_raw)[warp];

    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; s++) mbar_init(&ws.bar[s], 1);
        fence_mbar_init();
    }
    __syncwarp();
    const unsigned long long pol_first = policy_evict_first();
    const unsigned long long pol_last = policy_evict_last();

    const int nwarps = gridDim.x * C::WARPS;
    // consecutive tiles go to the warps of one CTA: neighbouring rows share x lines in L1
    int t = a.tile_begin + blockIdx.x * C::WARPS + warp;
    unsigned phase = 0;       // bit s = parity to wait for on stage s
    double r2 = 0.0;

    // tiles whose operator segments are requested before the wait: the ring's prologue, or (single stage) tile t
    constexpr int kPre = (STAGES > 1) ? STAGES - 1 : 1;
    if (PDL) {
        if (lane == 0) {
#pragma unroll
            for (int s = 0; s < kPre; s++)
                if (t + s * nwarps < a.tile_end)
                    tile_issue<C, OP, 1>(a, t + s * nwarps, ws.st[s], &ws.bar[s], pol_first);
        }
        pdl_wait();               // every thread: the predecessor's vectors are complete and visible from here on
        if (lane == 0) {
#pragma unroll
            for (int s = 0; s < kPre; s++)
                if (t + s * nwarps < a.tile_end)
                    tile_issue<C, OP, 2>(a, t + s * nwarps, ws.st[s], &ws.bar[s], pol_first);
        }
    } else if (lane == 0) {
#pragma unroll
        for (int s = 0; s < STAGES - 1; s++)
            if (t + s * nwarps < a.tile_end) tile_issue<C, OP>(a, t + s * nwarps, ws.st[s], &ws.bar[s], pol_first);
    }
    int stage = 0;
    bool requested = PDL && STAGES == 1;      // single stage + PDL: the first tile was requested above
    for (; t < a.tile_end; t += nwarps) {
        // keep the ring full: the stage consumed in the previous iteration is free again
        const int tn = t + (STAGES - 1) * nwarps;
        const int sn = (stage + STAGES - 1) % STAGES;
        if (lane == 0 && tn < a.tile_end && !requested) tile_issue<C, OP>(a, tn, ws.st[sn], &ws.bar[sn], pol_first);
        requested = false;

        const TileDesc d0 = a.tiles[t], d1 = a.tiles[t + 1];
        const int row0 = d0.row0, nrows = d1.row0 - d0.row0;
        const int s0 = d0.nz0, len = d1.nz0 - d0.nz0;
        mbar_wait(&ws.bar[stage], (phase >> stage) & 1u);
        phase ^= 1u << stage;
        const TileStageT<C, OP> &st = ws.st[stage];

        if (len <= C::T) {
            const int soff = s0 & ~3;                 // smem index = global entry index - soff
            const int poff = row0 & ~3;
            const int voff = row0 & ~1;
            for (int rbase = 0; rbase < nrows; rbase += RPP) {
                const int lr = rbase + grp;
                const bool active = lr < nrows;
                const int row = row0 + lr;
                int jb = 0, je = 0;
                if (active) {
                    jb = st.ptr[row - poff] - soff;
                    je = st.ptr[row - poff + 1] - soff;
                }
                double sum = 0.0, diag = 0.0;
                int jd = -1;
#pragma unroll 8
                for (int jj = jb + sub; jj < je; jj += G) {
                    const int c = st.col[jj];
                    const double v = st.val[jj];
                    double xv;
                    if (a.hints) xv = (OP == OP_GS) ? ld_x_hint(a.x + c, pol_last) : ld_x_hint_nc(a.x + c, pol_last);
                    else xv = (OP == OP_GS) ? a.x[c] : __ldg(a.x + c);
                    if (kNeedDiag && c == row) {
                        diag = v;
                        jd = jj;
                    } else {
                        sum += v * xv;
                    }
                }
                if (G > 1) {
                    sum = group_sum<G>(sum);
                    if (kNeedDiag) {
#pragma unroll
                        for (int o = G / 2; o > 0; o >>= 1) {
                            const int jo = __shfl_xor_sync(0xffffffffu, jd, o, G);
                            const double dv = __shfl_xor_sync(0xffffffffu, diag, o, G);
                            if (jo > jd) { jd = jo; diag = dv; }
                        }
                    }
                }
                if (active && sub == 0) {
                    if (OP == OP_SPMV) {
                        a.y[row] = sum;
                    } else if (OP == OP_RESID) {
                        const double r = st.bseg[row - voff] - sum;
                        a.y[row] = r;
                        r2 += r * r;
                    } else if (OP == OP_PADD) {
                        a.y[row] += sum;
                    } else if (OP == OP_JACOBI) {
                        const double xi = st.xseg[row - voff], bi = st.bseg[row - voff];
                        double xn = xi;
                        if (diag != 0.0) xn = (1.0 - a.omega) * xi + a.omega * ((bi - sum) / diag);
                        a.y[row] = xn;
                        if (a.r != nullptr) {
                            const double r = bi - sum - diag * xi;
                            a.r[row] = r;
                            r2 += r * r;
                        }
                    } else {
                        if (diag != 0.0) {
                            const double g = (st.bseg[row - voff] - sum) / diag;
                            a.y[row] = (a.omega == 1.0) ? g : a.omega * g + (1.0 - a.omega) * a.y[row];
                        }
                    }
                }
            }
        } else {
            // a single row longer than a tile: the whole warp strides over it in global memory
            const int row = row0;
            double sum = 0.0, diag = 0.0;
            int jd = -1;
            for (int jj = s0 + lane; jj < s0 + len; jj += 32) {
                const int c = ld_stream_i32(a.Aj + jj);
                const double v = ld_stream_f64(a.Ax + jj);
                const double xv = (OP == OP_GS) ? a.x[c] : __ldg(a.x + c);
                if (kNeedDiag && c == row) { diag = v; jd = jj; }
                else sum += v * xv;
            }
            sum = group_sum<32>(sum);
            if (kNeedDiag) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const int jo = __shfl_xor_sync(0xffffffffu, jd, o);
                    const double dv = __shfl_xor_sync(0xffffffffu, diag, o);
                    if (jo > jd) { jd = jo; diag = dv; }
                }
            }
            if (lane == 0) {
                if (OP == OP_SPMV) a.y[row] = sum;
                else if (OP == OP_RESID) { const double r = a.b[row] - sum; a.y[row] = r; r2 += r * r; }
                else if (OP == OP_PADD) a.y[row] += sum;
                else if (OP == OP_JACOBI) {
                    const double xi = a.x[row], bi = a.b[row];
                    a.y[row] = (diag != 0.0) ? (1.0 - a.omega) * xi + a.omega * ((bi - sum) / diag) : xi;
                    if (a.r != nullptr) { const double r = bi - sum - diag * xi; a.r[row] = r; r2 += r * r; }
                } else if (diag != 0.0) {
                    const double g = (a.b[row] - sum) / diag;
                    a.y[row] = (a.omega == 1.0) ? g : a.omega * g + (1.0 - a.omega) * a.y[row];
                }
            }
        }
        __syncwarp();             // every lane is done with this stage before the TMA refills it
        stage = (stage + 1) % STAGES;
    }
    if ((OP == OP_RESID || OP == OP_JACOBI) && a.partials != nullptr) {
        const double tsum = block_sum<C::WARPS * 32>(r2);
        if (threadIdx.x == 0) a.partials[blockIdx.x] = tsum;
    }
}

}  // namespace amgb
