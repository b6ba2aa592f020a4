// tess_index.cuh -- per-iteration spatial index of the generators (replaces the reference's
// per-iteration `cKDTree(node_xy)` build, tess/lloyd.pyx:54) and the per-tile candidate lists.
//
//   k_cell_count / k_cell_scan / k_cell_scatter : counting sort of the k generators into a
//       uniform grid.  The scatter also writes the generator coordinates (and 1/scale^2) in
//       cell order, so that walking a cell is one contiguous read with no gather.
//   k_build_lists : one warp per 64x64-px tile.  Ring search outward over grid cells; the
//       search radius is bounded by  R = min_h maxd2(h,tile)*inv_s2_h  (divided by the smallest
//       inv_s2, i.e. scaled by the LARGEST scale length), so the nearest generator of every
//       point of the tile is guaranteed to be in the list.
#pragma once
#include "tess_common.cuh"

// Generator cell + histogram.  Also clears the moment vector of this iteration (grid-stride),
// because it is the first kernel of an iteration that honours the convergence latch.
// cell_cnt is all-zero on entry: k_cell_scatter drains it back to zero.
__global__ void __launch_bounds__(256)
k_cell_count(const TessState* st, int honor, TessGrid g, const double* node, int k, int* cell_cnt, int* cell_of,
             double* mom, long n_mom) {
    if (TESS_STOPPED(st, honor)) return;
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long stride = (long)gridDim.x * blockDim.x;
    for (long j = i; j < n_mom; j += stride) mom[j] = 0.0;
    for (long j = i; j < k; j += stride) {
        int ci = tess_cell_coord(node[2 * j], g.ox, g.inv_cs, g.gw);
        int cj = tess_cell_coord(node[2 * j + 1], g.oy, g.inv_cs, g.gh);
        int c = cj * g.gw + ci;
        cell_of[j] = c;
        atomicAdd(&cell_cnt[c], 1);
    }
}

// Exclusive scan of cell_cnt[0..n_cells) -> cell_start[0..n_cells] (single CTA, 1024 threads):
// 16 consecutive cells per thread in registers, warp-shuffle scans, one carry per 16384 cells.
// Thread 0 also resets the per-iteration scalars of the state block.
#define TESS_SCAN_ITEMS 16
__global__ void __launch_bounds__(1024)
k_cell_scan(TessState* st, int honor, const int* cell_cnt, int* cell_start, int n_cells) {
    if (TESS_STOPPED(st, honor)) return;
    __shared__ int s_w[32];
    __shared__ int s_tot;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    int carry = 0;
    for (int base = 0; base < n_cells; base += 1024 * TESS_SCAN_ITEMS) {
        const int i0 = base + t * TESS_SCAN_ITEMS;
        int v[TESS_SCAN_ITEMS];
#pragma unroll
        for (int j = 0; j < TESS_SCAN_ITEMS; ++j) v[j] = (i0 + j < n_cells) ? cell_cnt[i0 + j] : 0;
        int tot = 0;
#pragma unroll
        for (int j = 0; j < TESS_SCAN_ITEMS; ++j) { const int x = v[j]; v[j] = tot; tot += x; }   // local exclusive
        int inc = tot;                                                                           // warp inclusive
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int y = __shfl_up_sync(TESS_FULL, inc, d);
            if (lane >= d) inc += y;
        }
        if (lane == 31) s_w[w] = inc;
        __syncthreads();
        if (w == 0) {
            const int x = s_w[lane];
            int wi = x;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int y = __shfl_up_sync(TESS_FULL, wi, d);
                if (lane >= d) wi += y;
            }
            s_w[lane] = wi - x;                      // exclusive prefix of the warp totals
            if (lane == 31) s_tot = wi;
        }
        __syncthreads();
        const int off = carry + s_w[w] + (inc - tot);
#pragma unroll
        for (int j = 0; j < TESS_SCAN_ITEMS; ++j)
            if (i0 + j < n_cells) cell_start[i0 + j] = off + v[j];
        carry += s_tot;
        __syncthreads();
    }
    if (t == 0) {
        cell_start[n_cells] = carry;
        if (honor) {   // part of a Lloyd iteration (not a stand-alone assignment query)
            st->delta = 0.0;
            st->n_changed = 0ull;
            st->count_changed = 0;
            st->counted = 0;
            st->min_inv_next = 0x7ff0000000000000ull;   // +inf
        }
        st->list_cursor = 0u;
        st->tile_counter = 0;
        st->n_brute = 0;
    }
}

__global__ void __launch_bounds__(256)
k_cell_scatter(const TessState* st, int honor, const int* cell_of, int k, const int* cell_start, int* cell_cnt,
               const double* node, const double* inv_s2, int* cell_items, double2* cell_xy, double* cell_inv) {
    if (TESS_STOPPED(st, honor)) return;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= k) return;
    const int c = cell_of[j];
    const int pos = cell_start[c] + atomicSub(&cell_cnt[c], 1) - 1;
    cell_items[pos] = j;
    cell_xy[pos] = make_double2(node[2 * (long)j], node[2 * (long)j + 1]);
    if (inv_s2) cell_inv[pos] = inv_s2[j];
}

// ---------------------------------------------------------------------------------------------
// Tile description for k_build_lists: tile t = (t % tiles_x, t / tiles_x) covers pixels
// [64*tx, 64*tx+63] x [64*ty, 64*ty+63] (clipped to the image) of a grid whose pixel (r,c)
// sits at (x0 + c, y0 + r).
struct TessTiles {
    double x0, y0;
    long W, H;
    int tiles_x, tiles_y;
    int n_tiles;
};

TESS_DEVFN TessRect tess_tile_rect(const TessTiles& T, int t) {
    TessRect r;
    int tx = t % T.tiles_x, ty = t / T.tiles_x;
    long px0 = (long)tx * TESS_TS, py0 = (long)ty * TESS_TS;
    long px1 = min(px0 + TESS_TS - 1, T.W - 1), py1 = min(py0 + TESS_TS - 1, T.H - 1);
    r.x0 = T.x0 + (double)px0;
    r.x1 = T.x0 + (double)px1;
    r.y0 = T.y0 + (double)py0;
    r.y1 = T.y0 + (double)py1;
    return r;
}

// Visit, warp-cooperatively, every generator slot stored in the grid cells
//   base, base+stride, ..., base+(count-1)*stride.
// 32 cells are inspected per step; the slot ranges of the non-empty ones are then walked by
// all lanes.  `f(slot)` is called by individual lanes (divergent); no collectives inside f.
template <class F>
TESS_DEVFN void tess_visit_cells(const int* cell_start, long base, long stride, int count, int lane, F f) {
    for (int c0 = 0; c0 < count; c0 += 32) {
        int c = c0 + lane;
        int s = 0, e = 0;
        if (c < count) {
            long cell = base + (long)c * stride;
            s = cell_start[cell];
            e = cell_start[cell + 1];
        }
        if (stride == 1) {
            // consecutive cells hold consecutive slots: one flat range [s(lane 0), max e).
            // (valid lanes' e are non-decreasing; lanes beyond `count` contribute 0)
            int first = __shfl_sync(TESS_FULL, s, 0);
            int last = e;
#pragma unroll
            for (int m = 16; m >= 1; m >>= 1) last = max(last, __shfl_xor_sync(TESS_FULL, last, m));
            for (int i = first + lane; i < last; i += 32) f(i);
        } else {
            unsigned any = __ballot_sync(TESS_FULL, e > s);
            while (any) {
                int src = __ffs((int)any) - 1;
                any &= any - 1;
                int ss = __shfl_sync(TESS_FULL, s, src);
                int ee = __shfl_sync(TESS_FULL, e, src);
                for (int i = ss + lane; i < ee; i += 32) f(i);
            }
        }
    }
}

#define TESS_LIST_WARPS 4

// One warp per tile (grid-stride over tiles).
template <bool HAS_SCALE>
__global__ void __launch_bounds__(TESS_LIST_WARPS * 32)
k_build_lists(TessState* st, int honor, TessTiles T, TessGrid g, const int* cell_start, const int* cell_items,
              const double2* cell_xy, const double* cell_inv, int* list_off, int* list_cnt, int* list_items,
              unsigned list_capacity) {
    if (TESS_STOPPED(st, honor)) return;
    __shared__ int s_stage[TESS_LIST_WARPS][TESS_CAP];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int warp = blockIdx.x * TESS_LIST_WARPS + wib;
    const int n_warps = gridDim.x * TESS_LIST_WARPS;
    const double min_inv = HAS_SCALE ? st->min_inv_s2 : 1.0;
    int* stage = s_stage[wib];

    for (int t = warp; t < T.n_tiles; t += n_warps) {
        TessRect r = tess_tile_rect(T, t);
        // cell range overlapped by the rectangle
        const int ci0 = tess_cell_coord(r.x0, g.ox, g.inv_cs, g.gw), ci1 = tess_cell_coord(r.x1, g.ox, g.inv_cs, g.gw);
        const int cj0 = tess_cell_coord(r.y0, g.oy, g.inv_cs, g.gh), cj1 = tess_cell_coord(r.y1, g.oy, g.inv_cs, g.gh);

        // ---- phase 1: R = min_h maxd2(h, tile) [* inv_s2_h], rings outward until provably complete
        double rbest = INFINITY;
        auto upd = [&](int slot) {
            const double2 p = cell_xy[slot];
            double d = tess_maxd2(r, p.x, p.y);
            if (HAS_SCALE) d *= cell_inv[slot];
            rbest = fmin(rbest, d);
        };
        int ring = 0;
        for (;;) {
            const int bi0 = ci0 - ring, bi1 = ci1 + ring, bj0 = cj0 - ring, bj1 = cj1 + ring;
            const int xi0 = max(bi0, 0), xi1 = min(bi1, g.gw - 1);
            const int xj0 = max(bj0, 0), xj1 = min(bj1, g.gh - 1);
            if (ring == 0) {
                for (int j = xj0; j <= xj1; ++j)
                    tess_visit_cells(cell_start, (long)j * g.gw + xi0, 1, xi1 - xi0 + 1, lane, upd);
            } else {
                if (bj0 >= 0) tess_visit_cells(cell_start, (long)bj0 * g.gw + xi0, 1, xi1 - xi0 + 1, lane, upd);
                if (bj1 < g.gh) tess_visit_cells(cell_start, (long)bj1 * g.gw + xi0, 1, xi1 - xi0 + 1, lane, upd);
                const int yj0 = max(bj0 + 1, 0), yj1 = min(bj1 - 1, g.gh - 1);
                if (yj1 >= yj0) {
                    if (bi0 >= 0) tess_visit_cells(cell_start, (long)yj0 * g.gw + bi0, g.gw, yj1 - yj0 + 1, lane, upd);
                    if (bi1 < g.gw) tess_visit_cells(cell_start, (long)yj0 * g.gw + bi1, g.gw, yj1 - yj0 + 1, lane, upd);
                }
            }
            rbest = tess_warp_min(rbest);
            const bool whole = (bi0 <= 0 && bj0 <= 0 && bi1 >= g.gw - 1 && bj1 >= g.gh - 1);
            if (whole) break;
            // every generator outside the box is at least `ring` whole cells away from the tile
            double lb = fmax(g.cs * (double)ring - g.tol, 0.0);
            if (lb * lb * min_inv > rbest * TESS_SLACK) break;
            ++ring;
        }
        const double thr = rbest * TESS_SLACK;

        // ---- phase 2: keep every generator of the box with mind2 [* inv_s2] <= R
        int cnt = 0;   // warp-uniform
        {
            const int xi0 = max(ci0 - ring, 0), xi1 = min(ci1 + ring, g.gw - 1);
            const int xj0 = max(cj0 - ring, 0), xj1 = min(cj1 + ring, g.gh - 1);
            for (int j = xj0; j <= xj1; ++j) {
                // cells xi0..xi1 of one grid row are one contiguous slot range
                const int ss = cell_start[(long)j * g.gw + xi0];
                const int ee = cell_start[(long)j * g.gw + xi1 + 1];
                for (int i0 = ss; i0 < ee; i0 += 32) {
                    const int i = i0 + lane;
                    int gid = -1;
                    bool keep = false;
                    if (i < ee) {
                        const double2 p = cell_xy[i];
                        double d = tess_mind2(r, p.x, p.y);
                        if (HAS_SCALE) d *= cell_inv[i];
                        keep = d <= thr;
                        if (keep) gid = cell_items[i];
                    }
                    unsigned b = __ballot_sync(TESS_FULL, keep);
                    int pos = cnt + __popc(b & tess_lanemask_lt(lane));
                    if (keep && pos < TESS_CAP) stage[pos] = gid;
                    cnt += __popc(b);
                }
            }
        }
        __syncwarp();
        // ---- sort by generator index (bitonic, in shared memory): every list derived from this one
        // is an order-preserving compaction, so the evaluators implement "lowest index wins a tie"
        // with a strict '<'
        if (cnt > 1 && cnt <= TESS_CAP) {
            int n2 = 2;
            while (n2 < cnt) n2 <<= 1;
            for (int i = cnt + lane; i < n2; i += 32) stage[i] = 0x7fffffff;
            __syncwarp();
            for (int k2 = 2; k2 <= n2; k2 <<= 1) {
                for (int j = k2 >> 1; j > 0; j >>= 1) {
                    for (int i = lane; i < n2; i += 32) {
                        const int ixj = i ^ j;
                        if (ixj > i) {
                            const int va = stage[i], vb = stage[ixj];
                            const bool up = ((i & k2) == 0);
                            if ((va > vb) == up) { stage[i] = vb; stage[ixj] = va; }
                        }
                    }
                    __syncwarp();
                }
            }
        }
        // ---- publish
        unsigned off = 0;
        int ok = 1;
        if (lane == 0) {
            if (cnt > TESS_CAP) ok = 0;
            else {
                off = atomicAdd(&st->list_cursor, (unsigned)cnt);
                if (off + (unsigned)cnt > list_capacity) ok = 0;
            }
            list_off[t] = ok ? (int)off : 0;
            list_cnt[t] = ok ? cnt : TESS_BRUTE;
            if (!ok) atomicAdd(&st->n_brute, 1);
        }
        off = __shfl_sync(TESS_FULL, off, 0);
        ok = __shfl_sync(TESS_FULL, ok, 0);
        if (ok)
            for (int i = lane; i < cnt; i += 32) list_items[off + i] = stage[i];
        __syncwarp();
    }
}
