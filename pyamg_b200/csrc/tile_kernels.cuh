// tile_kernels.cuh -- the sm_100a streaming kernel of the V-cycle: CSR row segments staged in shared
// memory by the TMA (1-D bulk copies, cp.async.bulk -> SASS UBLKCP) with mbarrier completion.
//
// Why: the hot path is HBM-bound CSR traffic (12 B per stored entry, 0.17 flop/B).  A lanes-per-row
// kernel (csr_kernels.cuh) keeps only a few hundred bytes per warp in flight and pays a dependent
// row-pointer -> entries -> gather latency chain per row group.  Here every warp owns a ring of
// shared-memory stages; one elected lane asks the TMA for the NEXT tile's three contiguous segments
// (values, column indices, row pointers -- each a single 16-B-aligned bulk copy) while the warp reduces
// the CURRENT tile out of shared memory.  The copy engine, not the LSU, streams the operator; the
// warps only issue the x gathers and the fused epilogue.  Bytes in flight per SM = warps x stage size
// (~100 KB), independent of register pressure.
//
// Tiles are built on the host at upload (engine.cu build_tiles): whole rows, at most T stored entries
// and RMAX rows, never crossing a Gauss-Seidel wave boundary; a row longer than T is a tile of its own
// and is streamed straight from global memory.  Work distribution is static (tile t -> warp t mod
// #warps), the grid persistent: 148 SMs x resident CTAs.
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
};

constexpr int kTileWarps = 8;            // warps per CTA
constexpr int kTileNnz = 512;            // T: stored entries per tile
constexpr int kTileRows = 128;           // RMAX: rows per tile
constexpr int kTileStages = 2;

struct __align__(16) TileStage {
    double val[kTileNnz + 8];
    int col[kTileNnz + 8];
    int ptr[kTileRows + 8];
};
struct __align__(16) TileWarpSmem {
    TileStage st[kTileStages];
    unsigned long long bar[kTileStages];
};
constexpr size_t kTileSmemBytes = sizeof(TileWarpSmem) * kTileWarps;

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
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
}
// 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier (TMA, SASS UBLKCP)
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// One lane: ask the TMA for tile t's segments.  All three sources start on 16-byte boundaries: the
// entry range is widened to a multiple of 4 entries on both sides (the arrays are padded at upload).
__device__ __forceinline__ void tile_issue(const TileArgs &a, int t, TileStage &st, unsigned long long *bar)
{
    const TileDesc d0 = a.tiles[t], d1 = a.tiles[t + 1];
    const int len = d1.nz0 - d0.nz0;
    if (len > kTileNnz) {          // long row: streamed from global memory, nothing to stage
        mbar_expect_tx(bar, 0);
        return;
    }
    const int s4 = d0.nz0 & ~3;
    const int cnt = ((d1.nz0 + 3) & ~3) - s4;
    const int r4 = d0.row0 & ~3;
    const int rcnt = ((d1.row0 + 1 + 3) & ~3) - r4;
    mbar_expect_tx(bar, (unsigned)(cnt * 12 + rcnt * 4));
    if (cnt > 0) {
        bulk_g2s(st.val, a.Ax + s4, (unsigned)cnt * 8u, bar);
        bulk_g2s(st.col, a.Aj + s4, (unsigned)cnt * 4u, bar);
    }
    bulk_g2s(st.ptr, a.Ap + r4, (unsigned)rcnt * 4u, bar);
}

// G lanes per row inside a tile (G = 1: thread per row -- 5/7-point stencils; larger G for the
// denser coarse operators).  32/G rows are reduced per pass.
template <int G, int OP>
__global__ void __launch_bounds__(kTileWarps * 32, 2) csr_tile_kernel(const TileArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr bool kNeedDiag = (OP == OP_JACOBI || OP == OP_GS);
    constexpr int RPP = 32 / G;                       // rows per pass
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sub = lane & (G - 1), grp = lane / G;
    TileWarpSmem &ws = reinterpret_cast<TileWarpSmem *>(smem_raw)[warp];

    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < kTileStages; s++) mbar_init(&ws.bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    const int nwarps = gridDim.x * kTileWarps;
    // consecutive tiles go to the warps of one CTA: neighbouring rows share x lines in L1
    int t = a.tile_begin + blockIdx.x * kTileWarps + warp;
    unsigned phase = 0;       // bit s = parity to wait for on stage s
    double r2 = 0.0;

    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < kTileStages - 1; s++)
            if (t + s * nwarps < a.tile_end) tile_issue(a, t + s * nwarps, ws.st[s], &ws.bar[s]);
    }
    int stage = 0;
    for (; t < a.tile_end; t += nwarps) {
        // keep the ring full: the stage consumed in the previous iteration is free again
        const int tn = t + (kTileStages - 1) * nwarps;
        const int sn = (stage + kTileStages - 1) % kTileStages;
        if (lane == 0 && tn < a.tile_end) tile_issue(a, tn, ws.st[sn], &ws.bar[sn]);

        const TileDesc d0 = a.tiles[t], d1 = a.tiles[t + 1];
        const int row0 = d0.row0, nrows = d1.row0 - d0.row0;
        const int s0 = d0.nz0, len = d1.nz0 - d0.nz0;
        mbar_wait(&ws.bar[stage], (phase >> stage) & 1u);
        phase ^= 1u << stage;
        const TileStage &st = ws.st[stage];

        if (len <= kTileNnz) {
            const int soff = s0 & ~3;                 // smem index = global entry index - soff
            const int poff = row0 & ~3;
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
#pragma unroll 4
                for (int jj = jb + sub; jj < je; jj += G) {
                    const int c = st.col[jj];
                    const double v = st.val[jj];
                    const double xv = (OP == OP_GS) ? a.x[c] : __ldg(a.x + c);
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
                        const double r = a.b[row] - sum;
                        a.y[row] = r;
                        r2 += r * r;
                    } else if (OP == OP_PADD) {
                        a.y[row] += sum;
                    } else if (OP == OP_JACOBI) {
                        const double xi = a.x[row], bi = a.b[row];
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
                            const double g = (a.b[row] - sum) / diag;
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
        stage = (stage + 1) % kTileStages;
    }
    if ((OP == OP_RESID || OP == OP_JACOBI) && a.partials != nullptr) {
        const double tsum = block_sum<kTileWarps * 32>(r2);
        if (threadIdx.x == 0) a.partials[blockIdx.x] = tsum;
    }
}

}  // namespace amgb
