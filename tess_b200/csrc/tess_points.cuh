// tess_points.cuh -- assignment (+ moments) for arbitrary point sets.
//
// Replaces `tree.query(xy, k=1)[1]` + the accumulate loop of lloyd() for a general `xy`
// (reference tess/lloyd.pyx:55,58-65) and the KD-tree query of
// VoronoiTessellation.partition_points (reference tess/voronoi.py:171-172).
//
// One thread per point.  The point looks up its cell in the generator grid and scans square
// rings of cells outward; after ring r every unscanned generator is at least r*cs away, so the
// scan stops as soon as (r*cs)^2 * min_inv_s2 exceeds the best distance found -- the nearest
// generator (lowest index on exact ties) is guaranteed.  Moments go straight to the global table
// with fp64 reductions (points carry no spatial order, so there is nothing to aggregate).
#pragma once
#include "tess_common.cuh"

template <bool HAS_SCALE>
TESS_DEVFN int tess_nearest_ring(const TessGrid& g, const int* cell_start, const int* cell_items,
                                 const double2* cell_xy, const double* cell_inv, double min_inv, double px, double py) {
    const int ci = tess_cell_coord(px, g.ox, g.inv_cs, g.gw);
    const int cj = tess_cell_coord(py, g.oy, g.inv_cs, g.gh);
    double bd = INFINITY;
    int bi = 0x7fffffff;
    // generators of the cells i0..i1 of grid row j: one contiguous slot range (two index loads)
    auto scan = [&](int i0, int i1, int j) {
        const long row = (long)j * g.gw;
        const int s = cell_start[row + i0], e = cell_start[row + i1 + 1];
        for (int q = s; q < e; ++q) {
            const double2 p = cell_xy[q];
            double d = tess_d2_pinned(px, py, p.x, p.y);
            if (HAS_SCALE) d = __dmul_rn(d, cell_inv[q]);
            if (d <= bd) {
                const int gid = cell_items[q];
                if ((d < bd) | (gid < bi)) { bd = d; bi = gid; }
            }
        }
    };
    for (int ring = 0;; ++ring) {
        const int bi0 = ci - ring, bi1 = ci + ring, bj0 = cj - ring, bj1 = cj + ring;
        const int xi0 = max(bi0, 0), xi1 = min(bi1, g.gw - 1);
        if (ring == 0) {
            scan(ci, ci, cj);
        } else {
            if (bj0 >= 0) scan(xi0, xi1, bj0);
            if (bj1 < g.gh) scan(xi0, xi1, bj1);
            const int yj0 = max(bj0 + 1, 0), yj1 = min(bj1 - 1, g.gh - 1);
            for (int j = yj0; j <= yj1; ++j) {
                if (bi0 >= 0) scan(bi0, bi0, j);
                if (bi1 < g.gw) scan(bi1, bi1, j);
            }
        }
        if (bi0 <= 0 && bj0 <= 0 && bi1 >= g.gw - 1 && bj1 >= g.gh - 1) break;
        const double lb = fmax(g.cs * (double)ring - g.tol, 0.0);
        if (lb * lb * min_inv > bd * TESS_SLACK) break;
    }
    return bi;
}

// MODE 0: labels only (partition_points); MODE 1: labels + 4 moments (lloyd on a point set)
template <int MODE, bool HAS_SCALE>
__global__ void __launch_bounds__(256)
k_points_main(TessState* st, int honor, TessGrid g, const int* cell_start, const int* cell_items,
              const double2* cell_xy, const double* cell_inv, long n, const double* xy, const double* w,
              int32_t* labels_ext, int32_t* lab0, int32_t* lab1, double* mom) {
    if (TESS_STOPPED(st, honor)) return;
    int32_t* const labels = labels_ext ? labels_ext : ((st->iters_done & 1) ? lab1 : lab0);
    const double min_inv = HAS_SCALE ? st->min_inv_s2 : 1.0;
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double px = xy[2 * i], py = xy[2 * i + 1];
    const int bi = tess_nearest_ring<HAS_SCALE>(g, cell_start, cell_items, cell_xy, cell_inv, min_inv, px, py);
    labels[i] = bi;
    if (MODE == 1) {
        const double wv = w[i];
        double* m = mom + 4 * (long)bi;
        atomicAdd(m, 1.0);
        atomicAdd(m + 1, wv);
        atomicAdd(m + 2, __dmul_rn(wv, px));
        atomicAdd(m + 3, __dmul_rn(wv, py));
    }
}
