// tess_points.cuh -- assignment (+ moments) for arbitrary point sets.
//
// Replaces `tree.query(xy, k=1)[1]` + the accumulate loop of lloyd() for a general `xy`
// (reference tess/lloyd.pyx:55,58-65) and the KD-tree query of
// VoronoiTessellation.partition_points (reference tess/voronoi.py:171-172).
//
// One thread per point.  The point looks up its cell in the generator grid and scans square
// rings of cells outward; after ring r every unscanned generator is at least r*cs away, so the
// scan stops as soon as (r*cs)^2 * min_inv_s2 exceeds the best distance found -- the nearest
// generator (lowest index on exact ties) is guaranteed.  The Lloyd path walks the points in spatial
// order (sorted once per tess_set_points), so a warp shares its cells and its bins.
#pragma once
#include "tess_common.cuh"

template <bool HAS_SCALE>
TESS_DEVFN int tess_nearest_ring(const TessGrid& g, const int* cell_start, const int* cell_items,
                                 const double2* cell_xy, const double* cell_inv, double min_inv, double px, double py) {
    const int ci = tess_cell_coord(px, g.ox, g.inv_cs, g.gw);
    const int cj = tess_cell_coord(py, g.oy, g.inv_cs, g.gh);
    double bd = INFINITY;
    int bi = 0x7fffffff;
    // records [s, e) of the cell-ordered table; four per trip: the loads are independent, only the
    // compares are serial
    auto scan = [&](int s, int e) {
        for (int q = s; q < e; q += 4) {
            double d[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int qq = min(q + u, e - 1);
                const double2 p = cell_xy[qq];
                d[u] = tess_d2_pinned(px, py, p.x, p.y);
                if (HAS_SCALE) d[u] = __dmul_rn(d[u], cell_inv[qq]);
                if (q + u >= e) d[u] = INFINITY;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (d[u] <= bd && q + u < e) {
                    const int gid = cell_items[q + u];
                    if ((d[u] < bd) | (gid < bi)) { bd = d[u]; bi = gid; }
                }
            }
        }
    };
    // The cells i0..i1 of one grid row are one contiguous record range (two index loads).
    // Rings 0 and 1 together (three row ranges, their index loads issued up front): in a grid with a
    // few generators per cell most queries end here, and the loads of a ring no longer wait for the
    // scan of the previous one.
    {
        const int xi0 = max(ci - 1, 0), xi1 = min(ci + 1, g.gw - 1);
        int rs[3], re[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const int j = cj - 1 + r;
            const bool in = (j >= 0) && (j < g.gh);
            const long row = (long)(in ? j : cj) * g.gw;
            rs[r] = cell_start[row + xi0];
            re[r] = in ? cell_start[row + xi1 + 1] : rs[r];
        }
        scan(rs[1], re[1]);
        scan(rs[0], re[0]);
        scan(rs[2], re[2]);
    }
    for (int ring = 1;; ++ring) {
        const int bi0 = ci - ring, bi1 = ci + ring, bj0 = cj - ring, bj1 = cj + ring;
        if (ring > 1) {
            const int xi0 = max(bi0, 0), xi1 = min(bi1, g.gw - 1);
            int s0 = 0, e0 = 0, s1 = 0, e1 = 0;
            if (bj0 >= 0) { s0 = cell_start[(long)bj0 * g.gw + xi0]; e0 = cell_start[(long)bj0 * g.gw + xi1 + 1]; }
            if (bj1 < g.gh) { s1 = cell_start[(long)bj1 * g.gw + xi0]; e1 = cell_start[(long)bj1 * g.gw + xi1 + 1]; }
            scan(s0, e0);
            scan(s1, e1);
            const int yj0 = max(bj0 + 1, 0), yj1 = min(bj1 - 1, g.gh - 1);
            for (int j = yj0; j <= yj1; ++j) {
                const long row = (long)j * g.gw;
                int sl = 0, el = 0, sr = 0, er = 0;
                if (bi0 >= 0) { sl = cell_start[row + bi0]; el = cell_start[row + bi0 + 1]; }
                if (bi1 < g.gw) { sr = cell_start[row + bi1]; er = cell_start[row + bi1 + 1]; }
                scan(sl, el);
                scan(sr, er);
            }
        }
        if (bi0 <= 0 && bj0 <= 0 && bi1 >= g.gw - 1 && bj1 >= g.gh - 1) break;
        const double lb = fmax(g.cs * (double)ring - g.tol, 0.0);
        if (lb * lb * min_inv > bd * TESS_SLACK) break;
    }
    return bi;
}

// ---- the point set in spatial order (built once per tess_set_points: points never move).  The grid
// is the point set's own (about 16 points per cell); `perm[i]` is the caller's index of sorted point i.
__global__ void __launch_bounds__(256)
k_pts_cell(TessGrid g, const double* xy, long n, int* cell_of, int* cell_cnt) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const int c = tess_cell_coord(xy[2 * i + 1], g.oy, g.inv_cs, g.gh) * g.gw + tess_cell_coord(xy[2 * i], g.ox, g.inv_cs, g.gw);
        cell_of[i] = c;
        atomicAdd(&cell_cnt[c], 1);
    }
}
__global__ void __launch_bounds__(256)
k_pts_scatter(const int* cell_of, const int* cell_start, int* cell_cnt, const double* xy, const double* w, long n,
              int* perm, double2* xy_s, double* w_s) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const int c = cell_of[i];
        const int pos = cell_start[c] + atomicSub(&cell_cnt[c], 1) - 1;
        perm[pos] = (int)i;
        xy_s[pos] = make_double2(xy[2 * i], xy[2 * i + 1]);
        w_s[pos] = w[i];
    }
}

// MODE 0: labels only (partition_points; points in the caller's order, `perm` null);
// MODE 1: labels + 4 moments (lloyd on a point set; points in spatial order, labels[perm[i]]).
// Moments: neighbouring threads hold neighbouring points, so a warp sees a handful of bins: the lanes
// of one bin are merged with a shuffle tree and one lane issues the four reductions (single lanes and
// pairs go direct).  Before: four fp64 atomics per point on 4 k addresses -- the L2 atomic units, not
// memory, bounded the kernel (110 us for 28 MB at 1e6 points).
template <int MODE, bool HAS_SCALE>
__global__ void __launch_bounds__(256)
k_points_main(TessState* st, int honor, TessGrid g, const int* cell_start, const int* cell_items,
              const double2* cell_xy, const double* cell_inv, long n, const double* xy, const double* w, const int* perm,
              int32_t* labels_ext, int32_t* lab0, int32_t* lab1, double* mom) {
    if (TESS_STOPPED(st, honor)) return;
    int32_t* const labels = labels_ext ? labels_ext : ((st->iters_done & 1) ? lab1 : lab0);
    const double min_inv = HAS_SCALE ? st->min_inv_s2 : 1.0;
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = i < n;
    double px = 0.0, py = 0.0;
    int bi = -1;
    if (valid) {
        px = xy[2 * i]; py = xy[2 * i + 1];
        bi = tess_nearest_ring<HAS_SCALE>(g, cell_start, cell_items, cell_xy, cell_inv, min_inv, px, py);
        labels[perm ? (long)perm[i] : i] = bi;
    }
    if (MODE == 1) {
        const int lane = threadIdx.x & 31;
        const double wv = valid ? w[i] : 0.0;
        double v0 = wv, v1 = __dmul_rn(wv, px), v2 = __dmul_rn(wv, py);
        unsigned todo = __ballot_sync(TESS_FULL, valid);
        bool direct = false;
        while (todo) {
            const int leader = __ffs((int)todo) - 1;
            const int key = __shfl_sync(TESS_FULL, bi, leader);
            const bool mine = valid && (bi == key);
            const unsigned grp = __ballot_sync(TESS_FULL, mine);
            todo &= ~grp;
            if (__popc(grp) <= 2) { direct = direct || mine; continue; }
            double a0 = mine ? v0 : 0.0, a1 = mine ? v1 : 0.0, a2 = mine ? v2 : 0.0;
#pragma unroll
            for (int m = 16; m >= 1; m >>= 1) {
                a0 = __dadd_rn(a0, __shfl_xor_sync(TESS_FULL, a0, m));
                a1 = __dadd_rn(a1, __shfl_xor_sync(TESS_FULL, a1, m));
                a2 = __dadd_rn(a2, __shfl_xor_sync(TESS_FULL, a2, m));
            }
            if (lane == leader) {
                double* m = mom + 4 * (long)key;
                atomicAdd(m, (double)__popc(grp));
                atomicAdd(m + 1, a0);
                atomicAdd(m + 2, a1);
                atomicAdd(m + 3, a2);
            }
        }
        if (direct) {
            double* m = mom + 4 * (long)bi;
            atomicAdd(m, 1.0);
            atomicAdd(m + 1, v0);
            atomicAdd(m + 2, v1);
            atomicAdd(m + 3, v2);
        }
    }
}
