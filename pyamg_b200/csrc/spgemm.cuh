// spgemm.cuh -- C = A B on the GPU with SciPy's results bit for bit: the Galerkin triple product of the setup
// phase.
//
// Reference call sites: `A = R @ A @ P` (pyamg/classical/classical.py:201, pyamg/aggregation/aggregation.py:425),
// evaluated left to right by SciPy's `csr_matmat` (scipy/sparse/sparsetools/csr.h; SciPy is a third-party
// dependency of the reference, un-pinned `scipy>=1.11`, 1.18.1 in this image -- its published algorithm, restated):
// for every row i, walk the entries (j, v) of A_i in storage order and the entries (k, w) of B_j in storage order,
// accumulate sums[k] += v*w, remember the order in which columns k first appear in a linked list that is
// prepended to; emit the list from its head (= REVERSE first-appearance order), dropping exact zeros.
//
// Here: "expand - sort - compress" per row in shared memory, arranged so that every floating-point operation and
// the output order equal SciPy's:
//   expand   products p_q = v*w (one rounding, no FMA) in SciPy's enumeration order q = 0, 1, ...; key (k << 32 | q)
//   sort     bitonic sort of the 64-bit keys (unique, so the order is the stable order by column)
//   compress one thread per column segment adds its products in ascending q -- exactly sums[k] += ... from 0.0
//   order    surviving (non-zero) columns are sorted by DESCENDING first q and written out
// Two passes (count, then fill) with the same code; rows are binned by their number of products (host) so that a
// 32-thread block handles the short rows and 256 threads the long ones.  Rows with more than 8192 products are
// outside the supported range (AMGB_ENOTIMPL; classical / SA Galerkin rows are 10^2..10^3).
//
// HBM-bound integer/float streaming like the rest of the path; no tensor cores (no dense contraction).
#pragma once
#include "csr_kernels.cuh"

namespace amgb {

struct SpgemmArgs {
    const int *rows;        // the rows this launch handles (one bin), n_rows of them
    int n_rows;
    const int *Ap, *Aj;
    const double *Ax;
    const int *Bp, *Bj;
    const double *Bx;
    int *row_nnz;           // COUNT pass: row_nnz[i] = entries of C_i
    const int *Cp;          // FILL pass: row pointers of C
    int *Cj;
    double *Cx;
    int fill;               // 0 = count, 1 = fill
};

constexpr unsigned long long kSpgemmPad = ~0ull;

template <int CAP>
constexpr size_t spgemm_smem_bytes() { return (size_t)CAP * 24 + ((size_t)CAP + 2) * 4; }

// in-place bitonic sort of N (power of two) 64-bit keys in shared memory by THREADS threads
template <int THREADS>
__device__ __forceinline__ void bitonic_sort_u64(unsigned long long *key, int N)
{
    for (int k = 2; k <= N; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < N; t += THREADS) {
                const int ixj = t ^ j;
                if (ixj > t) {
                    const unsigned long long a = key[t], b = key[ixj];
                    const bool up = (t & k) == 0;
                    if ((a > b) == up) { key[t] = b; key[ixj] = a; }
                }
            }
            __syncthreads();
        }
    }
}

// one block per row; dynamic shared memory: key[CAP] | key2[CAP] | val[CAP] | off[CAP + 2]
template <int CAP, int THREADS>
__global__ void __launch_bounds__(THREADS) spgemm_row_kernel(const SpgemmArgs a)
{
    extern __shared__ __align__(16) unsigned char spg_smem[];
    unsigned long long *key = reinterpret_cast<unsigned long long *>(spg_smem);
    unsigned long long *key2 = key + CAP;
    double *val = reinterpret_cast<double *>(key2 + CAP);
    int *off = reinterpret_cast<int *>(val + CAP);     // exclusive product offsets of the entries of A_i (+ total)
    __shared__ int s_cnt;

    const int i = a.rows[blockIdx.x];
    const int a0 = a.Ap[i];
    const int na = a.Ap[i + 1] - a0;                   // <= CAP (host binning)

    for (int t = threadIdx.x; t < na; t += THREADS) {
        const int j = a.Aj[a0 + t];
        off[t + 1] = a.Bp[j + 1] - a.Bp[j];
    }
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int t = 0; t < na; t++) {                 // in place: off[t + 1] (length of entry t) is read before off[t + 1]
            const int len = off[t + 1];                // is overwritten by the next step's off[t] store
            off[t] = run;
            run += len;
        }
        off[na] = run;
    }
    __syncthreads();
    const int total = off[na];                         // <= CAP (host binning)
    if (total == 0) {                                  // uniform across the block
        if (!a.fill && threadIdx.x == 0) a.row_nnz[i] = 0;
        return;
    }
    int N = 1;
    while (N < total) N <<= 1;

    // expand: products in SciPy's enumeration order
    for (int t = threadIdx.x; t < na; t += THREADS) {
        const int j = a.Aj[a0 + t];
        const double v = a.Ax[a0 + t];
        const int b0 = a.Bp[j], nb = a.Bp[j + 1] - b0;
        const int base = off[t];
        for (int q = 0; q < nb; q++) {
            const unsigned seq = (unsigned)(base + q);
            key[seq] = ((unsigned long long)(unsigned)a.Bj[b0 + q] << 32) | seq;
            val[seq] = __dmul_rn(v, a.Bx[b0 + q]);
        }
    }
    for (int t = total + threadIdx.x; t < N; t += THREADS) key[t] = kSpgemmPad;
    __syncthreads();
    bitonic_sort_u64<THREADS>(key, N);                 // ends with a barrier

    // compress (1): the thread at a segment head adds the segment up in ascending q, from 0.0 -- SciPy's order
    for (int p = threadIdx.x; p < N; p += THREADS) {
        double s = 0.0;
        if (p < total) {
            const unsigned col = (unsigned)(key[p] >> 32);
            if (p == 0 || (unsigned)(key[p - 1] >> 32) != col) {
                for (int q = p; q < total && (unsigned)(key[q] >> 32) == col; q++)
                    s = __dadd_rn(s, val[(unsigned)key[q]]);
            }
        }
        key2[p] = (unsigned long long)__double_as_longlong(s);
    }
    __syncthreads();                                   // every product has been read: val may be reused
    // compress (2): surviving columns get the key (descending first appearance, position); their sum moves to val[p]
    for (int p = threadIdx.x; p < N; p += THREADS) {
        unsigned long long k2 = kSpgemmPad;
        if (p < total) {
            const unsigned col = (unsigned)(key[p] >> 32);
            if (p == 0 || (unsigned)(key[p - 1] >> 32) != col) {
                const double s = __longlong_as_double((long long)key2[p]);
                if (s != 0.0) {                        // csr_matmat drops exact zeros
                    val[p] = s;
                    k2 = ((unsigned long long)(0xFFFFFFFFu - (unsigned)key[p]) << 32) | (unsigned)p;
                }
            }
        }
        key2[p] = k2;
    }
    __syncthreads();
    bitonic_sort_u64<THREADS>(key2, N);
    for (int p = threadIdx.x; p < N; p += THREADS)
        if (key2[p] != kSpgemmPad && (p == N - 1 || key2[p + 1] == kSpgemmPad)) s_cnt = p + 1;
    __syncthreads();
    const int cnt = s_cnt;
    if (!a.fill) {
        if (threadIdx.x == 0) a.row_nnz[i] = cnt;
        return;
    }
    const int c0 = a.Cp[i];
    for (int q = threadIdx.x; q < cnt; q += THREADS) {
        const unsigned p = (unsigned)key2[q];
        a.Cj[c0 + q] = (int)(unsigned)(key[p] >> 32);
        a.Cx[c0 + q] = val[p];
    }
}

}  // namespace amgb
