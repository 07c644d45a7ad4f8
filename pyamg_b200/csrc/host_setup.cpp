// host_setup.cpp -- host-side (CPU, C++) hierarchy-setup helpers: libpyamg_b200_host.so.
//
// NOT on the solve-phase hot path.  The reference's setup (strength / splitting / interpolation,
// pyamg/amg_core/ruge_stuben.h + graph.h) stays the authority whenever a hierarchy comes from the
// reference.  These routines exist so that the BASELINE inputs -- ruge_stuben_solver hierarchies,
// colour-sorted Gauss-Seidel row lists -- can be synthesised on the GPU box, where the reference
// is not installed, with the same published algorithms (Ruge & Stueben 1987; Briggs/Henson/
// McCormick 2000 ch. 8) and the same tie-breaking, so the hierarchies coincide with the
// reference's (checked in tests/test_setup.py against splittings / operators the reference
// produced).  SURVEY.md 8(f)-3/4 "next" rows start here.
//
//   amgb_setup_classical_strength  <->  classical_strength_of_connection (abs norm)   ruge_stuben.h:64-110
//                                       + |.|, row scaling, zero drop                 strength.py:236-241
//   amgb_setup_rs_splitting        <->  rs_cf_splitting (first pass)                  ruge_stuben.h:285-466
//   amgb_setup_classical_interp_*  <->  remove_strong_FF_connections + rs_classical_interpolation_pass1/2
//                                                                                     ruge_stuben.h:1083-1383
//   amgb_setup_greedy_coloring     <->  role of vertex_coloring (graph.h:218-235); first-fit greedy
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <limits>
#include <vector>

extern "C" {

// S = pattern of strong connections of A (|a_ij| >= theta * max_{k != i} |a_ik|, diagonal always kept),
// values |a_ij| / max_j |S_ij| per row, exact zeros dropped.  Sp/Sj/Sx/Sidx sized like A; Sidx[k] is
// the position in A of S's k-th entry (so A's values on S's pattern are Ax[Sidx]).  Returns nnz(S).
int64_t amgb_setup_classical_strength(int32_t n, const int32_t *Ap, const int32_t *Aj, const double *Ax,
                                      double theta, int32_t *Sp, int32_t *Sj, double *Sx, int32_t *Sidx)
{
    int64_t nnz = 0;
    Sp[0] = 0;
    for (int32_t i = 0; i < n; i++) {
        double max_off = std::numeric_limits<double>::min();
        for (int32_t jj = Ap[i]; jj < Ap[i + 1]; jj++)
            if (Aj[jj] != i) max_off = std::max(max_off, std::fabs(Ax[jj]));
        const double thr = theta * max_off;
        const int64_t row_begin = nnz;
        double row_max = 0.0;
        for (int32_t jj = Ap[i]; jj < Ap[i + 1]; jj++) {
            const double v = std::fabs(Ax[jj]);
            if (Aj[jj] == i || v >= thr) {
                Sj[nnz] = Aj[jj];
                Sx[nnz] = v;
                Sidx[nnz] = jj;
                row_max = std::max(row_max, v);
                nnz++;
            }
        }
        // scale by the largest entry of the row, then drop exact zeros
        int64_t w = row_begin;
        for (int64_t k = row_begin; k < nnz; k++) {
            const double v = (row_max != 0.0) ? Sx[k] / row_max : Sx[k];
            if (v != 0.0) { Sj[w] = Sj[k]; Sx[w] = v; Sidx[w] = Sidx[k]; w++; }
        }
        nnz = w;
        Sp[i + 1] = (int32_t)nnz;
    }
    return nnz;
}

namespace {
// Nodes kept sorted by an integer key in one array; each key owns a contiguous slice.  Raising or
// lowering a key by one is a swap with the slice boundary -- the classic O(1) bucket update used
// for the Ruge-Stueben measure.  Boundary conventions (raise: to the END of the slice, which then
// becomes the first slot of key+1; lower: to the BEGINNING, which becomes the last slot of key-1)
// determine the tie-breaking and hence the splitting.
struct MeasureBuckets {
    std::vector<int32_t> key, slot_of, node_at, first, size;
    void build(const std::vector<int32_t> &keys, int32_t nkeys)
    {
        const int32_t n = (int32_t)keys.size();
        key = keys;
        first.assign((size_t)nkeys, 0);
        size.assign((size_t)nkeys, 0);
        slot_of.resize((size_t)n);
        node_at.resize((size_t)n);
        for (int32_t i = 0; i < n; i++) size[(size_t)key[i]]++;
        int32_t run = 0;
        for (int32_t k = 0; k < nkeys; k++) { first[k] = run; run += size[k]; size[k] = 0; }
        for (int32_t i = 0; i < n; i++) {   // ascending node id inside a slice
            const int32_t s = first[key[i]] + size[key[i]]++;
            node_at[s] = i;
            slot_of[i] = s;
        }
    }
    void swap_slots(int32_t a, int32_t b)
    {
        const int32_t na = node_at[a], nb = node_at[b];
        node_at[a] = nb; node_at[b] = na;
        slot_of[na] = b; slot_of[nb] = a;
    }
    void raise(int32_t v)
    {
        const int32_t k = key[v];
        const int32_t last = first[k] + size[k] - 1;
        swap_slots(slot_of[v], last);
        size[k]--;
        size[k + 1]++;
        first[k + 1] = last;
        key[v] = k + 1;
    }
    void lower(int32_t v)
    {
        const int32_t k = key[v];
        const int32_t head = first[k];
        swap_slots(slot_of[v], head);
        size[k]--;
        size[k - 1]++;
        first[k]++;
        first[k - 1] = first[k] - size[k - 1];
        key[v] = k - 1;
    }
};
}  // namespace

// First-pass Ruge-Stueben C/F splitting on the strength graph S (no diagonal) and its transpose T.
// splitting[i] = 1 (C) or 0 (F).
void amgb_setup_rs_splitting(int32_t n, const int32_t *Sp, const int32_t *Sj, const int32_t *Tp,
                             const int32_t *Tj, int32_t *splitting)
{
    enum { F_PT = 0, C_PT = 1, UNDECIDED = 2, NEW_F = 3 };
    std::vector<int32_t> measure((size_t)n);
    int32_t top = 0;
    for (int32_t i = 0; i < n; i++) {
        measure[i] = Tp[i + 1] - Tp[i];      // how many points depend strongly on i
        top = std::max(top, measure[i]);
    }
    MeasureBuckets q;
    q.build(measure, std::max(2 * top, n + 1) + 2);
    for (int32_t i = 0; i < n; i++) {
        const bool isolated = q.key[i] == 0 || (q.key[i] == 1 && Tj[Tp[i]] == i);
        splitting[i] = isolated ? F_PT : UNDECIDED;
    }
    for (int32_t slot = n - 1; slot >= 0; slot--) {   // always the largest remaining measure
        const int32_t i = q.node_at[slot];
        q.size[q.key[i]]--;
        if (q.key[i] <= 0) break;
        if (splitting[i] != UNDECIDED) continue;
        splitting[i] = C_PT;
        // points that depend on the new C point become F points ...
        for (int32_t jj = Tp[i]; jj < Tp[i + 1]; jj++)
            if (splitting[Tj[jj]] == UNDECIDED) splitting[Tj[jj]] = NEW_F;
        // ... and every undecided point such an F point depends on gets more attractive
        for (int32_t jj = Tp[i]; jj < Tp[i + 1]; jj++) {
            const int32_t j = Tj[jj];
            if (splitting[j] != NEW_F) continue;
            splitting[j] = F_PT;
            for (int32_t kk = Sp[j]; kk < Sp[j + 1]; kk++) {
                const int32_t k = Sj[kk];
                if (splitting[k] == UNDECIDED && q.key[k] < n - 1) q.raise(k);
            }
        }
        // points the new C point depends on are less needed as C points
        for (int32_t jj = Sp[i]; jj < Sp[i + 1]; jj++) {
            const int32_t j = Sj[jj];
            if (splitting[j] == UNDECIDED && q.key[j] != 0) q.lower(j);
        }
    }
    for (int32_t i = 0; i < n; i++)
        if (splitting[i] != C_PT) splitting[i] = F_PT;
}

static inline int sign_of(double v) { return v < 0 ? -1 : 1; }

// Modified classical interpolation, pass 0+1: on the strength pattern S (same row order as A's
// strong entries, diagonal allowed) zero out -- keep[jj] = 0 -- strong F-F connections of F rows
// that share no strong C point, then count the entries of P.  Pp has n+1 entries.
void amgb_setup_classical_interp_count(int32_t n, const int32_t *Sp, const int32_t *Sj,
                                       const int32_t *splitting, uint8_t *keep, int32_t *Pp)
{
    const int64_t nnzS = Sp[n];
    for (int64_t k = 0; k < nnzS; k++) keep[k] = 1;
    // rows are independent here (each writes only its own keep[] entries): threads when built with OpenMP
#pragma omp parallel for schedule(dynamic, 4096)
    for (int32_t i = 0; i < n; i++) {
        if (splitting[i] != 0) continue;
        for (int32_t jj = Sp[i]; jj < Sp[i + 1]; jj++) {
            const int32_t j = Sj[jj];
            if (splitting[j] != 0) continue;
            bool common = false;
            for (int32_t ii = Sp[i]; ii < Sp[i + 1] && !common; ii++) {
                const int32_t c = Sj[ii];
                if (splitting[c] != 1) continue;
                for (int32_t kk = Sp[j]; kk < Sp[j + 1]; kk++)
                    if (Sj[kk] == c) { common = true; break; }
            }
            if (!common) keep[jj] = 0;
        }
    }
    int32_t nnz = 0;
    Pp[0] = 0;
    for (int32_t i = 0; i < n; i++) {
        if (splitting[i] == 1) {
            nnz++;
        } else {
            for (int32_t jj = Sp[i]; jj < Sp[i + 1]; jj++)
                if (keep[jj] && splitting[Sj[jj]] == 1 && Sj[jj] != i) nnz++;
        }
        Pp[i + 1] = nnz;
    }
}

// pass 2: weights.  For an F point i with strong C set C_i and strong F set F_i (after the F-F
// filter):  w_ij = -( a_ij + sum_{k in F_i} a_ik a~_kj / sum_{l in C_i} a~_kl ) / ( a_ii + sum_{weak} a_im ),
// where a~_kx = a_kx if sign(a_kx) != sign(a_kk), else 0 ("modified" classical interpolation).
// Sv holds A's values on the (filtered) strength pattern: Sv[jj] = a_{i,Sj[jj]}.
void amgb_setup_classical_interp_fill(int32_t n, const int32_t *Ap, const int32_t *Aj, const double *Ax,
                                      const int32_t *Sp, const int32_t *Sj, const double *Sv,
                                      const uint8_t *keep, const int32_t *splitting, const int32_t *Pp,
                                      int32_t *Pj, double *Px)
{
    std::vector<int32_t> cmap((size_t)n);
    for (int32_t i = 0, c = 0; i < n; i++) { cmap[i] = c; c += splitting[i]; }
    // every row fills its own, pre-counted slice of P: independent -> threads when built with OpenMP
#pragma omp parallel for schedule(dynamic, 4096)
    for (int32_t i = 0; i < n; i++) {
        if (splitting[i] == 1) {
            Pj[Pp[i]] = cmap[i];
            Px[Pp[i]] = 1.0;
            continue;
        }
        double denom = 0.0;
        for (int32_t mm = Ap[i]; mm < Ap[i + 1]; mm++) denom += Ax[mm];
        for (int32_t mm = Sp[i]; mm < Sp[i + 1]; mm++)
            if (keep[mm] && Sj[mm] != i) denom -= Sv[mm];
        int32_t out = Pp[i];
        for (int32_t jj = Sp[i]; jj < Sp[i + 1]; jj++) {
            const int32_t j = Sj[jj];
            if (!keep[jj] || splitting[j] != 1) continue;
            double numer = Sv[jj];
            for (int32_t kk = Sp[i]; kk < Sp[i + 1]; kk++) {
                const int32_t k = Sj[kk];
                if (!keep[kk] || splitting[k] != 0 || k == i) continue;
                const double a_ik = Sv[kk];
                double a_kj = 0.0, a_kk = 0.0;
                for (int32_t s = Ap[k]; s < Ap[k + 1]; s++) {
                    if (Aj[s] == j) a_kj = Ax[s];
                    else if (Aj[s] == k) a_kk = Ax[s];
                }
                if (sign_of(a_kj) == sign_of(a_kk)) a_kj = 0.0;
                if (std::fabs(a_kj) > 1e-15 * std::fabs(a_ik)) {
                    double inner = 0.0;
                    for (int32_t ll = Sp[i]; ll < Sp[i + 1]; ll++) {
                        const int32_t l = Sj[ll];
                        if (!keep[ll] || splitting[l] != 1) continue;
                        for (int32_t s = Ap[k]; s < Ap[k + 1]; s++) {
                            if (Aj[s] == l) {
                                if (sign_of(Ax[s]) != sign_of(a_kk)) inner += Ax[s];
                                break;
                            }
                        }
                    }
                    numer += a_ik * a_kj / inner;
                }
            }
            Pj[out] = cmap[j];
            Px[out] = -numer / denom;
            out++;
        }
    }
}

// First-fit greedy vertex colouring in natural order (diagonal ignored). Returns #colours.
int32_t amgb_setup_greedy_coloring(int32_t n, const int32_t *Ap, const int32_t *Aj, int32_t *colors)
{
    std::vector<int32_t> mark;   // mark[c] == i  <=>  colour c is taken by a neighbour of i
    int32_t ncol = 0;
    for (int32_t i = 0; i < n; i++) colors[i] = -1;
    for (int32_t i = 0; i < n; i++) {
        for (int32_t jj = Ap[i]; jj < Ap[i + 1]; jj++) {
            const int32_t j = Aj[jj];
            if (j != i && j >= 0 && j < n && colors[j] >= 0) mark[(size_t)colors[j]] = i;
        }
        int32_t c = 0;
        while (c < ncol && mark[(size_t)c] == i) c++;
        if (c == ncol) { mark.push_back(-1); ncol++; }
        colors[i] = c;
    }
    return ncol;
}

// Greedy colouring in SMALLEST-LAST order (Matula & Beck): repeatedly remove a vertex of minimum remaining
// degree; colour in reverse removal order.  Uses at most (degeneracy + 1) colours -- noticeably fewer than
// natural-order first fit on the dense coarse operators of an RS hierarchy, i.e. fewer dependent waves per
// Gauss-Seidel sweep.  `order` = 0: natural, 1: smallest-last, 2: largest-degree-first.
int32_t amgb_setup_greedy_coloring_ordered(int32_t n, const int32_t *Ap, const int32_t *Aj, int32_t order,
                                           int32_t *colors)
{
    if (order == 0) return amgb_setup_greedy_coloring(n, Ap, Aj, colors);
    std::vector<int32_t> seq((size_t)n);
    std::vector<int32_t> deg((size_t)n);
    int32_t maxdeg = 0;
    for (int32_t i = 0; i < n; i++) {
        int32_t d = 0;
        for (int32_t jj = Ap[i]; jj < Ap[i + 1]; jj++) d += (Aj[jj] != i);
        deg[(size_t)i] = d;
        maxdeg = std::max(maxdeg, d);
    }
    if (order == 2) {
        for (int32_t i = 0; i < n; i++) seq[(size_t)i] = i;
        std::stable_sort(seq.begin(), seq.end(), [&](int32_t a, int32_t b) { return deg[(size_t)a] > deg[(size_t)b]; });
    } else {
        // bucket queue over current degrees
        std::vector<int32_t> head((size_t)maxdeg + 1, -1), next((size_t)n, -1), prev((size_t)n, -1);
        std::vector<char> removed((size_t)n, 0);
        auto push = [&](int32_t v) {
            const int32_t d = deg[(size_t)v];
            prev[(size_t)v] = -1; next[(size_t)v] = head[(size_t)d];
            if (head[(size_t)d] >= 0) prev[(size_t)head[(size_t)d]] = v;
            head[(size_t)d] = v;
        };
        auto unlink = [&](int32_t v) {
            const int32_t d = deg[(size_t)v];
            if (prev[(size_t)v] >= 0) next[(size_t)prev[(size_t)v]] = next[(size_t)v]; else head[(size_t)d] = next[(size_t)v];
            if (next[(size_t)v] >= 0) prev[(size_t)next[(size_t)v]] = prev[(size_t)v];
        };
        for (int32_t i = n - 1; i >= 0; i--) push(i);
        int32_t cur = 0;
        for (int32_t k = n - 1; k >= 0; k--) {
            while (cur > 0 && head[(size_t)cur - 1] >= 0) cur--;
            while (head[(size_t)cur] < 0) cur++;
            const int32_t v = head[(size_t)cur];
            unlink(v);
            removed[(size_t)v] = 1;
            seq[(size_t)k] = v;                      // coloured in reverse removal order
            for (int32_t jj = Ap[v]; jj < Ap[v + 1]; jj++) {
                const int32_t u = Aj[jj];
                if (u == v || u < 0 || u >= n || removed[(size_t)u]) continue;
                unlink(u);
                deg[(size_t)u]--;
                push(u);
            }
            if (cur > 0) cur--;
        }
    }
    std::vector<int32_t> mark;
    int32_t ncol = 0;
    for (int32_t i = 0; i < n; i++) colors[i] = -1;
    for (int32_t k = 0; k < n; k++) {
        const int32_t i = seq[(size_t)k];
        for (int32_t jj = Ap[i]; jj < Ap[i + 1]; jj++) {
            const int32_t j = Aj[jj];
            if (j != i && j >= 0 && j < n && colors[j] >= 0) mark[(size_t)colors[j]] = i;
        }
        int32_t c = 0;
        while (c < ncol && mark[(size_t)c] == i) c++;
        if (c == ncol) { mark.push_back(-1); ncol++; }
        colors[i] = c;
    }
    return ncol;
}

// ---- smoothed aggregation (scalar problems) --------------------------------------------------------
// Symmetric strength (Vanek/Mandel/Brezina 1996): keep a_ij, i != j, iff |a_ij|^2 >= theta^2 |a_ii| |a_jj|;
// the diagonal is always kept.  Only the pattern is returned (aggregation uses nothing else).
//   <-> symmetric_strength_of_connection, pyamg/amg_core/smoothed_aggregation.h:56-108
int64_t amgb_setup_symmetric_strength(int32_t n, const int32_t *Ap, const int32_t *Aj, const double *Ax,
                                      double theta, int32_t *Sp, int32_t *Sj)
{
    std::vector<double> diag((size_t)n, 0.0);
    for (int32_t i = 0; i < n; i++) {
        double d = 0.0;
        for (int32_t jj = Ap[i]; jj < Ap[i + 1]; jj++)
            if (Aj[jj] == i) d += Ax[jj];                 // duplicates of the diagonal are summed
        diag[(size_t)i] = std::fabs(d);
    }
    int64_t nnz = 0;
    Sp[0] = 0;
    for (int32_t i = 0; i < n; i++) {
        const double eps_i = theta * theta * diag[(size_t)i];
        for (int32_t jj = Ap[i]; jj < Ap[i + 1]; jj++) {
            const int32_t j = Aj[jj];
            if (j == i || Ax[jj] * Ax[jj] >= eps_i * diag[(size_t)j]) Sj[nnz++] = j;
        }
        Sp[i + 1] = (int32_t)nnz;
    }
    return nnz;
}

// Standard (greedy, three-pass) aggregation on the strength pattern: pass 1 turns every node whose
// neighbourhood is still free into the root of an aggregate together with its neighbours; pass 2 attaches
// leftover nodes to an adjacent aggregate; pass 3 aggregates what remains.  agg[i] = aggregate id or -1
// (isolated node); roots[] receives the root node of every aggregate.  Returns the number of aggregates.
//   <-> standard_aggregation, pyamg/amg_core/smoothed_aggregation.h:138-236
int32_t amgb_setup_standard_aggregation(int32_t n, const int32_t *Sp, const int32_t *Sj, int32_t *agg,
                                        int32_t *roots)
{
    const int32_t FREE = -2, ISOLATED = -1;
    std::vector<int32_t> tentative((size_t)n, -1);       // pass-2 attachments, resolved after the pass
    for (int32_t i = 0; i < n; i++) agg[i] = FREE;
    int32_t count = 0;
    for (int32_t i = 0; i < n; i++) {                    // pass 1
        if (agg[i] != FREE) continue;
        bool any = false, taken = false;
        for (int32_t jj = Sp[i]; jj < Sp[i + 1] && !taken; jj++) {
            const int32_t j = Sj[jj];
            if (j == i) continue;
            any = true;
            if (agg[j] != FREE) taken = true;
        }
        if (!any) { agg[i] = ISOLATED; continue; }
        if (taken) continue;
        roots[count] = i;
        agg[i] = count;
        for (int32_t jj = Sp[i]; jj < Sp[i + 1]; jj++) agg[Sj[jj]] = count;
        count++;
    }
    for (int32_t i = 0; i < n; i++) {                    // pass 2: join a neighbour's PASS-1 aggregate
        if (agg[i] != FREE) continue;
        for (int32_t jj = Sp[i]; jj < Sp[i + 1]; jj++) {
            const int32_t a = agg[Sj[jj]];
            if (a >= 0) { tentative[(size_t)i] = a; break; }
        }
    }
    for (int32_t i = 0; i < n; i++)
        if (agg[i] == FREE && tentative[(size_t)i] >= 0) agg[i] = tentative[(size_t)i];
    for (int32_t i = 0; i < n; i++) {                    // pass 3
        if (agg[i] != FREE) continue;
        roots[count] = i;
        agg[i] = count;
        for (int32_t jj = Sp[i]; jj < Sp[i + 1]; jj++)
            if (agg[Sj[jj]] == FREE) agg[Sj[jj]] = count;
        count++;
    }
    return count;
}

// Sequential Gauss-Seidel sweeps on the host for candidate improvement during setup (A x = b, in place);
// `symmetric` != 0: forward then backward per iteration.  Zero diagonal leaves the row untouched.
void amgb_setup_gauss_seidel(int32_t n, const int32_t *Ap, const int32_t *Aj, const double *Ax, double *x,
                             const double *b, int32_t iterations, int32_t symmetric)
{
    auto sweep = [&](int32_t start, int32_t stop, int32_t step) {
        for (int32_t i = start; i != stop; i += step) {
            double rsum = 0.0, d = 0.0;
            for (int32_t jj = Ap[i]; jj < Ap[i + 1]; jj++) {
                if (Aj[jj] == i) d = Ax[jj];
                else rsum += Ax[jj] * x[Aj[jj]];
            }
            if (d != 0.0) x[i] = (b[i] - rsum) / d;
        }
    };
    for (int32_t it = 0; it < iterations; it++) {
        sweep(0, n, 1);
        if (symmetric) sweep(n - 1, -1, -1);
    }
}

// Sequential BLOCK Gauss-Seidel sweeps on the host (candidate improvement for vector problems): block row i of
// the BSR operator (blocks bs x bs, row-major) against x, then x_i = Dinv_i (b_i - sum_{j != i} A_ij x_j) with
// the (pseudo-)inverted diagonal blocks supplied by the caller.  `symmetric` != 0: forward then backward.
void amgb_setup_block_gauss_seidel(int32_t nb, int32_t bs, const int32_t *Ap, const int32_t *Aj, const double *Ax,
                                   const double *Dinv, double *x, const double *b, int32_t iterations,
                                   int32_t symmetric)
{
    const int64_t bb = (int64_t)bs * bs;
    std::vector<double> rsum((size_t)bs);
    auto sweep = [&](int32_t start, int32_t stop, int32_t step) {
        for (int32_t i = start; i != stop; i += step) {
            std::fill(rsum.begin(), rsum.end(), 0.0);
            for (int32_t jj = Ap[i]; jj < Ap[i + 1]; jj++) {
                const int32_t j = Aj[jj];
                if (j == i) continue;
                const double *blk = Ax + (int64_t)jj * bb;
                const double *xj = x + (int64_t)j * bs;
                for (int32_t r = 0; r < bs; r++) {
                    double v = 0.0;
                    for (int32_t c = 0; c < bs; c++) v += blk[(int64_t)r * bs + c] * xj[c];
                    rsum[(size_t)r] += v;
                }
            }
            for (int32_t r = 0; r < bs; r++) rsum[(size_t)r] = b[(int64_t)i * bs + r] - rsum[(size_t)r];
            const double *di = Dinv + (int64_t)i * bb;
            for (int32_t r = 0; r < bs; r++) {
                double v = 0.0;
                for (int32_t c = 0; c < bs; c++) v += di[(int64_t)r * bs + c] * rsum[(size_t)c];
                x[(int64_t)i * bs + r] = v;
            }
        }
    };
    for (int32_t it = 0; it < iterations; it++) {
        sweep(0, nb, 1);
        if (symmetric) sweep(nb - 1, -1, -1);
    }
}

// 1 if no stored off-diagonal entry joins two rows of equal colour (checks a colouring computed on a
// pattern assumed symmetric), else 0.
int32_t amgb_setup_coloring_is_valid(int32_t n, const int32_t *Ap, const int32_t *Aj, const int32_t *colors)
{
    for (int32_t i = 0; i < n; i++)
        for (int32_t jj = Ap[i]; jj < Ap[i + 1]; jj++) {
            const int32_t j = Aj[jj];
            if (j != i && j >= 0 && j < n && colors[j] == colors[i]) return 0;
        }
    return 1;
}

// ---------------------------------------------------------------------------------------------
// One round of modified-Gram-Schmidt Arnoldi on the host, multi-threaded (the spectral-radius estimates of the
// smoother setup: pyamg/util/linalg.py:90-252 _approximate_eigenvalues, non-symmetric branch :214-229).  Same
// contract as the device version amgb_arnoldi_run (csrc/abi_operator.cuh): V is (maxiter+1) x n row-major with the
// (unnormalised) start vector in row 0, the operator is x -> diag(row_scale) (A x) (row_scale may be NULL), H is
// (maxiter+1) x maxiter row-major; returns the number of valid steps m (0 for a zero start vector; a step whose
// new norm is below breakdown * max(1, max |H[:j+1,:j+1]|) ends the round at m = j + 1).
// Inner products are summed over fixed 8192-entry chunks whose partial sums are added in chunk order: the result
// does not depend on the number of threads.
// ---------------------------------------------------------------------------------------------
static double chunked_dot(const double *x, const double *y, int64_t n)
{
    const int64_t C = 8192, nch = (n + C - 1) / C;
    std::vector<double> part((size_t)std::max<int64_t>(nch, 1), 0.0);
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < nch; c++) {
        const int64_t i0 = c * C, i1 = std::min(n, i0 + C);
        double s = 0.0;
        for (int64_t i = i0; i < i1; i++) s += x[i] * y[i];
        part[(size_t)c] = s;
    }
    double t = 0.0;
    for (int64_t c = 0; c < nch; c++) t += part[(size_t)c];
    return t;
}

int32_t amgb_setup_arnoldi_round(int32_t n, const int32_t *Ap, const int32_t *Aj, const double *Ax,
                                 const double *row_scale, double *V, int32_t maxiter, double breakdown, double *H)
{
    const int64_t N = n;
    for (int64_t k = 0; k < (int64_t)(maxiter + 1) * maxiter; k++) H[k] = 0.0;
    double *v0 = V;
    const double nv = std::sqrt(chunked_dot(v0, v0, N));
    if (!(nv > 0.0)) return 0;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; i++) v0[i] /= nv;
    double hmax = 0.0;
    for (int32_t j = 0; j < maxiter; j++) {
        const double *vj = V + (size_t)j * N;
        double *w = V + (size_t)(j + 1) * N;
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < N; i++) {
            double sum = 0.0;
            for (int32_t jj = Ap[i]; jj < Ap[i + 1]; jj++) sum += Ax[jj] * vj[Aj[jj]];
            w[i] = row_scale ? row_scale[i] * sum : sum;
        }
        for (int32_t i = 0; i <= j; i++) {
            const double *vi = V + (size_t)i * N;
            const double h = chunked_dot(vi, w, N);
            H[(size_t)i * maxiter + j] = h;
#pragma omp parallel for schedule(static)
            for (int64_t k = 0; k < N; k++) w[k] -= h * vi[k];
        }
        const double hn = std::sqrt(chunked_dot(w, w, N));
        H[(size_t)(j + 1) * maxiter + j] = hn;
        for (int32_t i = 0; i <= j; i++) hmax = std::max(hmax, std::fabs(H[(size_t)i * maxiter + j]));
        for (int32_t jj = 0; jj < j; jj++) hmax = std::max(hmax, std::fabs(H[(size_t)j * maxiter + jj]));
        if (!(hn >= breakdown * std::max(1.0, hmax))) return j + 1;
#pragma omp parallel for schedule(static)
        for (int64_t k = 0; k < N; k++) w[k] /= hn;
    }
    return maxiter;
}

}  // extern "C"
