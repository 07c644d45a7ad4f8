// ref_shim.cpp -- extern "C" window onto the REFERENCE's own native sweep kernels.
//
// TEST INFRASTRUCTURE ONLY (see oracle/amg_oracle.c header).  This file contains no
// algorithm: it #includes the reference header *where it lies* under /root/reference
// (passed with -I by oracle/build.py; never copied into this repo) and instantiates the
// <int, double, double> templates behind C symbols, so that
//   * tests can check the C restatement (amg_oracle.c) against the real thing, and
//   * bench.py's cpu_baseline / --impl reference legs can time the reference's compiled
//     code ("kind": "reference") on the GPU box's host cores.
// Output: oracle/_ref/libamg_ref.so (git-ignored, travels to the GPU box).
//
// Built only where /root/reference exists (this container); the GPU box uses the prebuilt .so.
// standard headers the reference header relies on its includer for (its *_bind.cpp gets them via pybind11)
#include <algorithm>
#include <cmath>
#include <complex>
#include <iostream>
#include <vector>

#include "relaxation.h"   // -> /root/reference/pyamg/amg_core/relaxation.h (+ linalg.h)

extern "C" {

// pyamg/amg_core/relaxation.h:309-346
void ref_jacobi(const int *Ap, int n, const int *Aj, const double *Ax, int nnz, double *x,
                const double *b, double *temp, int row_start, int row_stop, int row_step,
                double omega)
{
    jacobi<int, double, double>(Ap, n + 1, Aj, nnz, Ax, nnz, x, n, b, n, temp, n,
                                row_start, row_stop, row_step, &omega, 1);
}

// pyamg/amg_core/relaxation.h:48-76
void ref_gauss_seidel(const int *Ap, int n, const int *Aj, const double *Ax, int nnz,
                      double *x, const double *b, int row_start, int row_stop, int row_step)
{
    gauss_seidel<int, double, double>(Ap, n + 1, Aj, nnz, Ax, nnz, x, n, b, n,
                                      row_start, row_stop, row_step);
}

// pyamg/amg_core/relaxation.h:736-768
void ref_gauss_seidel_indexed(const int *Ap, int n, const int *Aj, const double *Ax, int nnz,
                              double *x, const double *b, const int *Id, int n_id,
                              int row_start, int row_stop, int row_step)
{
    gauss_seidel_indexed<int, double, double>(Ap, n + 1, Aj, nnz, Ax, nnz, x, n, b, n,
                                              Id, n_id, row_start, row_stop, row_step);
}

// pyamg/amg_core/relaxation.h:472-562
void ref_bsr_jacobi(const int *Ap, int nb, const int *Aj, const double *Ax, int nblk,
                    double *x, const double *b, double *temp, int row_start, int row_stop,
                    int row_step, int bs, double omega)
{
    int n = nb * bs;
    bsr_jacobi<int, double, double>(Ap, nb + 1, Aj, nblk, Ax, nblk * bs * bs, x, n, b, n,
                                    temp, n, row_start, row_stop, row_step, bs, &omega, 1);
}

// pyamg/amg_core/relaxation.h:1021-1090
void ref_block_jacobi(const int *Ap, int nb, const int *Aj, const double *Ax, int nblk,
                      double *x, const double *b, const double *Dinv, double *temp,
                      int row_start, int row_stop, int row_step, double omega, int bs)
{
    int n = nb * bs;
    block_jacobi<int, double, double>(Ap, nb + 1, Aj, nblk, Ax, nblk * bs * bs, x, n, b, n,
                                      Dinv, nb * bs * bs, temp, n, row_start, row_stop,
                                      row_step, &omega, 1, bs);
}

// pyamg/amg_core/relaxation.h:382-427
void ref_jacobi_indexed(const int *Ap, int n, const int *Aj, const double *Ax, int nnz, double *x,
                        const double *b, const int *Id, int n_id, double omega)
{
    jacobi_indexed<int, double, double>(Ap, n + 1, Aj, nnz, Ax, nnz, x, n, b, n, Id, n_id, &omega, 1);
}

// pyamg/amg_core/relaxation.h:1242-1298
void ref_block_gauss_seidel(const int *Ap, int nb, const int *Aj, const double *Ax, int nblk,
                            double *x, const double *b, const double *Dinv, int row_start,
                            int row_stop, int row_step, int bs)
{
    int n = nb * bs;
    block_gauss_seidel<int, double, double>(Ap, nb + 1, Aj, nblk, Ax, nblk * bs * bs, x, n, b, n,
                                            Dinv, nb * bs * bs, row_start, row_stop, row_step, bs);
}

// pyamg/amg_core/relaxation.h:1113-1172
void ref_block_jacobi_indexed(const int *Ap, int nb, const int *Aj, const double *Ax, int nblk, double *x,
                              const double *b, const double *Dinv, const int *Id, int n_id, double omega, int bs)
{
    int n = nb * bs;
    block_jacobi_indexed<int, double, double>(Ap, nb + 1, Aj, nblk, Ax, nblk * bs * bs, x, n, b, n,
                                              Dinv, nb * bs * bs, Id, n_id, &omega, 1, bs);
}

// pyamg/amg_core/relaxation.h:579-606
void ref_jacobi_ne(const int *Ap, int n, const int *Aj, const double *Ax, int nnz, double *x, const double *b,
                   const double *delta, double *temp, double omega)
{
    jacobi_ne<int, double, double>(Ap, n + 1, Aj, nnz, Ax, nnz, x, n, b, n, delta, n, temp, n, 0, n, 1, &omega, 1);
}

// pyamg/amg_core/relaxation.h:633-657
void ref_gauss_seidel_ne(const int *Ap, int n, const int *Aj, const double *Ax, int nnz, double *x, const double *b,
                         int row_start, int row_stop, int row_step, const double *Dinv, double omega)
{
    gauss_seidel_ne<int, double, double>(Ap, n + 1, Aj, nnz, Ax, nnz, x, n, b, n, row_start, row_stop, row_step,
                                         Dinv, n, omega);
}

// pyamg/amg_core/relaxation.h:684-713
void ref_gauss_seidel_nr(const int *Ap, int n, const int *Aj, const double *Ax, int nnz, double *x, double *z,
                         int col_start, int col_stop, int col_step, const double *Dinv, double omega)
{
    gauss_seidel_nr<int, double, double>(Ap, n + 1, Aj, nnz, Ax, nnz, x, n, z, n, col_start, col_stop, col_step,
                                         Dinv, n, omega);
}

// pyamg/amg_core/relaxation.h:818-880
void ref_overlapping_schwarz_csr(const int *Ap, int n, const int *Aj, const double *Ax, int nnz, double *x,
                                 const double *b, const double *Tx, int tx_size, const int *Tp, const int *Sj,
                                 int sj_size, const int *Sp, int nsub, int row_start, int row_stop, int row_step)
{
    overlapping_schwarz_csr<int, double, double>(Ap, n + 1, Aj, nnz, Ax, nnz, x, n, b, n, Tx, tx_size, Tp, nsub + 1,
                                                 Sj, sj_size, Sp, nsub + 1, nsub, n, row_start, row_stop, row_step);
}

}  // extern "C"
