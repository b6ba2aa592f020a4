/*
 * oracle/lloyd_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C, single-threaded (optionally OpenMP for the order-independent argmin only)
 * CPU restatement of the Lloyd / centroidal-Voronoi hot path of jonathansick/tess.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may link or call this file.  The product (tess_b200/) never does.
 *
 * Parity pin: this restatement is checked bit-for-bit against the *compiled reference
 * itself* (tess/lloyd.pyx built unmodified into oracle/_ref/, see oracle/build_ref.py)
 * and against the committed fixtures in tests/golden/ that were generated from it
 * (tests/golden/make_golden.py).  See tests/test_oracle.py.
 *
 * What is restated, with the reference lines it follows:
 *   - assignment          tess/lloyd.pyx:54-55  (cKDTree(node_xy).query(xy,k=1)[1])
 *       SciPy's cKDTree is a third-party dependency (scipy, un-pinned in the reference;
 *       1.18.1 here).  Its 1-NN result on tie-free input equals
 *           argmin_j fl( fl(dx*dx) + fl(dy*dy) )      (fp64, NO fma)
 *       This file fixes the tie rule to LOWEST INDEX (cKDTree's is tree-order dependent,
 *       SURVEY.md section 0 fact 5) and adds the scale-weighted (WVT) extension
 *           argmin_j fl( fl(fl(dx*dx)+fl(dy*dy)) * inv_s2_j ),  inv_s2_j = fl(1/fl(s_j*s_j))
 *       which has no reference counterpart and degenerates bit-exactly at inv_s2 == NULL.
 *   - accumulate          tess/lloyd.pyx:58-65  (sequential, point order, products rounded)
 *   - node update         tess/lloyd.pyx:66-74  (population test, empty -> (0,0))
 *   - delta / stop rule   tess/lloyd.pyx:76-90
 *   - image point order   tess/voronoi.py:281-287 (x = column, y = row, row-major, finite only)
 *   - per-bin S/N         tess/pixel_accretion.py:501-517 (sum S / sqrt(sum N^2))
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared [-fopenmp] (oracle/Makefile).
 * -ffp-contract=off is REQUIRED: the reference is built without FMA contraction.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* d^2 under the pinned formula. */
static inline double d2_pinned(double px, double py, double gx, double gy)
{
    double dx = px - gx;
    double dy = py - gy;
    double dxx = dx * dx;
    double dyy = dy * dy;
    return dxx + dyy;
}

/*
 * Brute-force assignment of n points (AoS xy[n][2]) to k generators (AoS node[k][2]).
 * inv_s2 may be NULL (plain Euclidean, reference parity mode).
 * Lowest index wins ties (strict '<' scanning j ascending).
 * Restates tess/lloyd.pyx:54-55.
 */
void oracle_assign_points(long n, const double *xy, long k, const double *node,
                          const double *inv_s2, int64_t *idx)
{
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (long i = 0; i < n; ++i) {
        double px = xy[2 * i], py = xy[2 * i + 1];
        double best = INFINITY;
        long bj = 0;
        for (long j = 0; j < k; ++j) {
            double d = d2_pinned(px, py, node[2 * j], node[2 * j + 1]);
            if (inv_s2)
                d = d * inv_s2[j];
            if (d < best) {
                best = d;
                bj = j;
            }
        }
        idx[i] = bj;
    }
}

/*
 * Same for an H x W pixel grid with coordinates x = fl(x0 + col), y = fl(y0 + row)
 * (tess/voronoi.py:281-283 for from_image with x0=y0=0; tess/voronoi.py:107-114 for segmap
 * with x0=xlim[0], y0=ylim[0]).  labels is row-major [H][W].
 */
void oracle_assign_grid(long H, long W, double x0, double y0, long k, const double *node,
                        const double *inv_s2, int64_t *labels)
{
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (long r = 0; r < H; ++r) {
        double py = y0 + (double)r;
        for (long c = 0; c < W; ++c) {
            double px = x0 + (double)c;
            double best = INFINITY;
            long bj = 0;
            for (long j = 0; j < k; ++j) {
                double d = d2_pinned(px, py, node[2 * j], node[2 * j + 1]);
                if (inv_s2)
                    d = d * inv_s2[j];
                if (d < best) {
                    best = d;
                    bj = j;
                }
            }
            labels[r * W + c] = bj;
        }
    }
}

/*
 * Sequential accumulate in point-index order, products rounded before the add
 * (tess/lloyd.pyx:58-65).  mom is [k][6]: count, sum w, sum w*x, sum w*y, sum S, sum N^2.
 * signal / noise may be NULL (then columns 4,5 stay 0).  Points with idx < 0 are skipped
 * (masked pixels; the reference drops them before calling lloyd, tess/voronoi.py:285-287).
 */
void oracle_accumulate(long n, const double *xy, const double *w, const int64_t *idx,
                       const double *signal, const double *noise, long k, double *mom)
{
    memset(mom, 0, sizeof(double) * 6 * (size_t)k);
    for (long i = 0; i < n; ++i) {
        long l = idx[i];
        if (l < 0)
            continue;
        double *m = mom + 6 * l;
        double wx = w[i] * xy[2 * i];
        double wy = w[i] * xy[2 * i + 1];
        m[0] += 1.0;
        m[1] += w[i];
        m[2] += wx;
        m[3] += wy;
        if (signal)
            m[4] += signal[i];
        if (noise) {
            double n2 = noise[i] * noise[i];
            m[5] += n2;
        }
    }
}

/*
 * Node update + delta (tess/lloyd.pyx:66-81).  new_node[k][2] is written; returns delta.
 * Emptiness is tested on the population, not on the weight (lloyd.pyx:68).
 */
double oracle_update_nodes(long k, const double *mom, const double *old_node, double *new_node)
{
    for (long j = 0; j < k; ++j) {
        const double *m = mom + 6 * j;
        if (m[0] > 0.0) {
            new_node[2 * j] = m[2] / m[1];
            new_node[2 * j + 1] = m[3] / m[1];
        } else {
            new_node[2 * j] = 0.0;
            new_node[2 * j + 1] = 0.0;
        }
    }
    double delta = 0.0;
    for (long j = 0; j < k; ++j) {
        double dx = old_node[2 * j] - new_node[2 * j];
        double dy = old_node[2 * j + 1] - new_node[2 * j + 1];
        delta += dx * dx + dy * dy;
    }
    return delta;
}

/*
 * Full loop with the reference's stop rule (tess/lloyd.pyx:47-90):
 *   after iteration t (0-based): delta == 0 -> converged; else t > max_iters -> not converged.
 * node[k][2] is updated IN PLACE to the returned node table (the reference returns a fresh
 * array; the Python wrapper copies first).  idx[n] is the assignment against the node table
 * of the LAST executed iteration's input (lloyd.pyx:55 vs :86/:88).
 * Returns converged (1/0); *n_iters_out = number of loop bodies executed; deltas (nullable,
 * capacity max_iters+2) receives the per-iteration delta printed at lloyd.pyx:82.
 */
int oracle_lloyd(long n, const double *xy, const double *w, long k, double *node,
                 long max_iters, int64_t *idx, long *n_iters_out, double *deltas, double *mom_out)
{
    double *mom = (double *)malloc(sizeof(double) * 6 * (size_t)k);
    double *nn = (double *)malloc(sizeof(double) * 2 * (size_t)k);
    long n_iters = 0;
    long executed = 0;
    int converged = 0;
    for (;;) {
        oracle_assign_points(n, xy, k, node, NULL, idx);
        oracle_accumulate(n, xy, w, idx, NULL, NULL, k, mom);
        double delta = oracle_update_nodes(k, mom, node, nn);
        memcpy(node, nn, sizeof(double) * 2 * (size_t)k);
        if (deltas)
            deltas[executed] = delta;
        ++executed;
        if (delta == 0.0) {
            converged = 1;
            break;
        } else if (n_iters > max_iters) {
            converged = 0;
            break;
        } else {
            ++n_iters;
        }
    }
    if (mom_out)
        memcpy(mom_out, mom, sizeof(double) * 6 * (size_t)k);
    if (n_iters_out)
        *n_iters_out = executed;
    free(mom);
    free(nn);
    return converged;
}

int oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
