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
// because it is the first phase of an iteration that honours the convergence latch.
// cell_cnt is all-zero on entry: the scatter drains it back to zero.
// (Every O(k) step of the iteration is written as a grid-stride device function `tess_p_*` that works
//  for any grid of 256-thread CTAs: the stand-alone kernels below wrap them one to one, and
//  k_iteration_phases (tess_phases.cuh) runs several of them in ONE cooperative launch with grid-wide
//  barriers in between.)
TESS_DEVFN void tess_p_cell_count(TessGrid g, const double* node, int k, int* cell_cnt, int* cell_of, double* mom, long n_mom) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long stride = (long)gridDim.x * blockDim.x;
    for (long j = i; j < n_mom; j += stride) mom[j] = 0.0;
    for (long j0 = i; j0 < k; j0 += 4 * stride) {          // four independent generators per trip
        double2 xy[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long j = j0 + u * stride;
            xy[u] = (j < k) ? *reinterpret_cast<const double2*>(node + 2 * j) : make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long j = j0 + u * stride;
            if (j < k) {
                const int c = tess_cell_coord(xy[u].y, g.oy, g.inv_cs, g.gh) * g.gw + tess_cell_coord(xy[u].x, g.ox, g.inv_cs, g.gw);
                cell_of[j] = c;
                atomicAdd(&cell_cnt[c], 1);
            }
        }
    }
}
__global__ void __launch_bounds__(256)
k_cell_count(const TessState* st, int honor, TessGrid g, const double* node, int k, int* cell_cnt, int* cell_of,
             double* mom, long n_mom) {
    if (TESS_STOPPED(st, honor)) return;
    tess_p_cell_count(g, node, k, cell_cnt, cell_of, mom, n_mom);
}

// Largest number of generators in one cell of grid g (set-up only: decides how far the grid is
// refined, tess_b200.cu::image_geometry).  cell_cnt is all-zero on entry; the caller clears it again.
__global__ void __launch_bounds__(256)
k_cell_occupancy_max(TessState* st, TessGrid g, const double* node, int k, int* cell_cnt) {
    int m = 0;
    for (long j = (long)blockIdx.x * blockDim.x + threadIdx.x; j < k; j += (long)gridDim.x * blockDim.x) {
        const int ci = tess_cell_coord(node[2 * j], g.ox, g.inv_cs, g.gw);
        const int cj = tess_cell_coord(node[2 * j + 1], g.oy, g.inv_cs, g.gh);
        m = max(m, atomicAdd(&cell_cnt[cj * g.gw + ci], 1) + 1);
    }
    m = __reduce_max_sync(TESS_FULL, m);
    if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(&st->tile_counter, m);
}

// Exclusive scan of cell_cnt[0..n_cells) -> cell_start[0..n_cells], reduce-then-scan over blocks of
// TESS_SCAN_BLOCK cells: k_cell_blocksum writes one total per block, k_cell_scan adds the totals of
// the preceding blocks to a register/shuffle scan of its own block (16 consecutive cells per thread).
// Block 0 of k_cell_scan also resets the per-iteration scalars of the state block.
#define TESS_SCAN_ITEMS 16
#define TESS_SCAN_THREADS 256
#define TESS_SCAN_BLOCK (TESS_SCAN_ITEMS * TESS_SCAN_THREADS)

TESS_DEVFN void tess_p_cell_blocksum(const int* cell_cnt, int n_cells, int* block_sum) {
    __shared__ int s_w[TESS_SCAN_THREADS / 32];
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const int nsb = (n_cells + TESS_SCAN_BLOCK - 1) / TESS_SCAN_BLOCK;
    for (int b = blockIdx.x; b < nsb; b += gridDim.x) {
        const int base = b * TESS_SCAN_BLOCK;
        int tot = 0;
#pragma unroll
        for (int j = 0; j < TESS_SCAN_ITEMS; ++j) {
            const int i = base + j * TESS_SCAN_THREADS + t;      // coalesced
            tot += (i < n_cells) ? cell_cnt[i] : 0;
        }
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) tot += __shfl_xor_sync(TESS_FULL, tot, m);
        __syncthreads();                 // s_w of the previous trip has been read
        if (lane == 0) s_w[w] = tot;
        __syncthreads();
        if (t == 0) {
            int s = 0;
            for (int i = 0; i < TESS_SCAN_THREADS / 32; ++i) s += s_w[i];
            block_sum[b] = s;
        }
    }
}
__global__ void __launch_bounds__(TESS_SCAN_THREADS)
k_cell_blocksum(const TessState* st, int honor, const int* cell_cnt, int n_cells, int* block_sum) {
    if (TESS_STOPPED(st, honor)) return;
    tess_p_cell_blocksum(cell_cnt, n_cells, block_sum);
}

// Scalars of the state block that belong to one list build (done by one thread while the cells are
// scanned).  (flags_next is never cleared here: once raised the iteration is over until
// tess_set_generators resets the whole block, and other CTAs of a phased launch may be reading it.)
TESS_DEVFN void tess_reset_iteration_scalars(TessState* st, int honor) {
    (void)honor;
    st->list_cursor = 0u;
    st->tile_counter = 0;      // quadrant queues of the interior / edge image kernels
    st->spare_ = 0u;
    st->n_brute = 0;
    st->n_overflow = 0;
}
// Scalars that one accumulate + update pair produces (done by one thread of the first kernel of the
// accumulate): they stay readable -- tess_get_iter_stats -- until the next accumulate starts.
TESS_DEVFN void tess_reset_accumulate_scalars(TessState* st) {
    st->n_changed = 0ull;
    st->count_changed = 0;
    st->counted = 0;
    st->min_inv_next = 0x7ff0000000000000ull;   // +inf
}

TESS_DEVFN void tess_p_cell_scan(TessState* st, int honor, const int* cell_cnt, int* cell_start, int n_cells, const int* block_sum) {
    __shared__ int s_w[TESS_SCAN_THREADS / 32];
    __shared__ int s_base;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const int nsb = (n_cells + TESS_SCAN_BLOCK - 1) / TESS_SCAN_BLOCK;
    for (int b = blockIdx.x; b < nsb; b += gridDim.x) {
        // totals of the preceding blocks (a few hundred at most), or -- no block sums -- of the preceding cells
        int pre = 0;
        if (block_sum) {
            for (int i = t; i < b; i += TESS_SCAN_THREADS) pre += block_sum[i];
        } else {
            for (int i = t; i < b * TESS_SCAN_BLOCK; i += TESS_SCAN_THREADS) pre += cell_cnt[i];
        }
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) pre += __shfl_xor_sync(TESS_FULL, pre, m);
        __syncthreads();
        if (lane == 0) s_w[w] = pre;
        __syncthreads();
        if (t == 0) {
            int s = 0;
            for (int i = 0; i < TESS_SCAN_THREADS / 32; ++i) s += s_w[i];
            s_base = s;
        }
        __syncthreads();
        const int carry = s_base;
        __syncthreads();
        const int i0 = b * TESS_SCAN_BLOCK + t * TESS_SCAN_ITEMS;
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
        int woff = 0;
        for (int i = 0; i < w; ++i) woff += s_w[i];
        const int off = carry + woff + (inc - tot);
#pragma unroll
        for (int j = 0; j < TESS_SCAN_ITEMS; ++j)
            if (i0 + j < n_cells) cell_start[i0 + j] = off + v[j];
        if (i0 <= n_cells - 1 && n_cells - 1 < i0 + TESS_SCAN_ITEMS) cell_start[n_cells] = off + tot;   // owner of the last cell
    }
    if (blockIdx.x == 0 && t == 0) tess_reset_iteration_scalars(st, honor);
}
__global__ void __launch_bounds__(TESS_SCAN_THREADS)
k_cell_scan(TessState* st, int honor, const int* cell_cnt, int* cell_start, int n_cells, const int* block_sum) {
    if (TESS_STOPPED(st, honor)) return;
    tess_p_cell_scan(st, honor, cell_cnt, cell_start, n_cells, block_sum);
}

// Grids of up to TESS_SINGLE_SCAN_MAX cells are scanned WITHOUT the block-sum phase (and its grid
// barrier): tess_p_cell_scan with block_sum == nullptr lets every scan block add up the cells that
// precede it by itself (coalesced; at most 16 blocks of 4096 cells, i.e. 256 KB of reads per CTA).
#define TESS_SINGLE_SCAN_MAX 65536

TESS_DEVFN void tess_p_cell_scatter(const int* cell_of, int k, const int* cell_start, int* cell_cnt, const double* node,
                                    const double* inv_s2, int* cell_items, double2* cell_xy, double* cell_inv) {
    // four generators per thread and trip: the slot atomics (with return) and the loads behind them are
    // ~1 us round trips each; issued four at a time they overlap instead of queueing up
    const long stride = (long)gridDim.x * blockDim.x;
    for (long j0 = (long)blockIdx.x * blockDim.x + threadIdx.x; j0 < k; j0 += 4 * stride) {
        int c[4], pos[4];
        double2 xy[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long j = j0 + u * stride;
            c[u] = (j < k) ? cell_of[j] : 0;
            xy[u] = (j < k) ? *reinterpret_cast<const double2*>(node + 2 * j) : make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (j0 + u * stride < k) pos[u] = cell_start[c[u]] + atomicSub(&cell_cnt[c[u]], 1) - 1;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long j = j0 + u * stride;
            if (j < k) {
                cell_items[pos[u]] = (int)j;
                cell_xy[pos[u]] = xy[u];
                if (inv_s2) cell_inv[pos[u]] = inv_s2[j];
            }
        }
    }
}
__global__ void __launch_bounds__(256)
k_cell_scatter(const TessState* st, int honor, const int* cell_of, int k, const int* cell_start, int* cell_cnt,
               const double* node, const double* inv_s2, int* cell_items, double2* cell_xy, double* cell_inv) {
    if (TESS_STOPPED(st, honor)) return;
    tess_p_cell_scatter(cell_of, k, cell_start, cell_cnt, node, inv_s2, cell_items, cell_xy, cell_inv);
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

// ---------------------------------------------------------------------------------------------
// Per-quadrant candidate lists.  A quadrant is a 32x32-px quarter of a tile; its list holds every
// generator that can be nearest to some pixel of the quadrant, as arena records (coordinates,
// generator index[, 1/scale^2]) in no particular order (the evaluators break exact ties by index).
struct TessQLists {
    int* q_off;        // [nqx*nqy] first arena record of the quadrant's list
    int* q_cnt;        // [nqx*nqy] records; 0: quadrant outside the window; -1: overflow -> scan all generators
    double2* xy;       // arena
    int* gid;
    double* inv;       // arena (scaled metric only)
    int* ring_q;       // [nqx*nqy] the quadrants off the fast path (st->n_brute of them), any order
    unsigned capacity; // arena records
    int nqx, nqy;      // quadrant grid (2 tiles_x, 2 tiles_y)
};

#define TESS_LIST_WARPS 2
#define TESS_LEAF_MIN_BUILD 3   // tile lists up to this length are handed to every quadrant unfiltered

// One warp per 64x64 tile (grid-stride).
//  1. Box search: the generators of the grid cells within `ring` cells of the tile are scanned for
//     R = min_h maxd2(h, tile) [* inv_s2_h]; the ring doubles until every generator outside the box is
//     provably too far, (ring cs)^2 min_inv > R -- the search radius is bounded by the LARGEST scale.
//  2. The box generators with mind2 [* inv_s2] <= R are staged in shared memory: no other generator
//     can be nearest to any pixel of the tile.
//  3. Each of the four quadrants keeps those that pass the centre + radius test
//     u(c,g) <= min_h u(c,h) + 2 rho / s_min and writes them to the arena.
template <bool HAS_SCALE>
__global__ void __launch_bounds__(TESS_LIST_WARPS * 32)
k_build_qlists(TessState* st, int honor, TessTiles T, TessGrid g, const int* cell_start, const int* cell_items,
               const double2* cell_xy, const double* cell_inv, TessQLists L, int* tile_ring, int dense, int qcap,
               double* mom_clear, long n_clear) {
    if (TESS_STOPPED(st, honor)) return;
    // first kernel of an accumulate in image mode: the moment rows of the previous iteration go
    // (their tail slots were cleared with the scan of the generator grid)
    for (long j = (long)blockIdx.x * blockDim.x + threadIdx.x; j < n_clear; j += (long)gridDim.x * blockDim.x) mom_clear[j] = 0.0;
    if (honor && blockIdx.x == 0 && threadIdx.x == 0) tess_reset_accumulate_scalars(st);
    __shared__ __align__(16) double2 s_xy[TESS_LIST_WARPS][TESS_CAP];
    __shared__ double s_iv[HAS_SCALE ? TESS_LIST_WARPS : 1][HAS_SCALE ? TESS_CAP : 1];
    __shared__ unsigned s_key[TESS_LIST_WARPS][TESS_CAP];
    __shared__ int s_cnt[TESS_LIST_WARPS];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int warp = blockIdx.x * TESS_LIST_WARPS + wib;
    const int n_warps = gridDim.x * TESS_LIST_WARPS;
    const double min_inv = HAS_SCALE ? st->min_inv_s2 : 1.0;
    double2* const xy = s_xy[wib];
    double* const iv = HAS_SCALE ? s_iv[wib] : nullptr;
    unsigned* const key = s_key[wib];       // generator index

    for (int t = warp; t < T.n_tiles; t += n_warps) {
        const TessRect r = tess_tile_rect(T, t);
        const int tx = t % T.tiles_x, ty = t / T.tiles_x;
        // cell range overlapped by the rectangle
        const int ci0 = tess_cell_coord(r.x0, g.ox, g.inv_cs, g.gw), ci1 = tess_cell_coord(r.x1, g.ox, g.inv_cs, g.gw);
        const int cj0 = tess_cell_coord(r.y0, g.oy, g.inv_cs, g.gh), cj1 = tess_cell_coord(r.y1, g.oy, g.inv_cs, g.gh);

        const int ring_hint = tile_ring[t];     // requested before step 0 so that its latency overlaps
        // ---- 0. a tile that holds more generators than a list can (bins of a few pixels): every one
        //      of them is a candidate, so the tile overflows whatever R is -- hand its quadrants to the
        //      per-pixel ring search of the image kernel without walking any box.  Only asked on
        //      refined grids (`dense`: some tile is known to be that full); on the coarse grid the
        //      check would cost every tile a dependent load (7 us of 68 at C3) for nothing.
        if (dense) {
            int tot0 = 0;
            for (int j = cj0 + lane; j <= cj1; j += 32)
                tot0 += cell_start[(long)j * g.gw + ci1 + 1] - cell_start[(long)j * g.gw + ci0];
            tot0 = __reduce_add_sync(TESS_FULL, tot0);
            if (tot0 > TESS_CAP) {
                if (lane < 4) {
                    const int q = lane;
                    const long px0 = (long)tx * TESS_TS + 32 * (q & 1), py0 = (long)ty * TESS_TS + 32 * (q >> 1);
                    const bool in = (px0 < T.W) && (py0 < T.H);
                    const long qid = (long)(2 * ty + (q >> 1)) * L.nqx + (2 * tx + (q & 1));
                    L.q_off[qid] = 0;
                    L.q_cnt[qid] = in ? -1 : 0;
                    if (in) { L.ring_q[atomicAdd(&st->n_brute, 1)] = (int)qid; atomicAdd(&st->n_overflow, 1); }
                }
                continue;
            }
        }

        // ---- 1. box search for R.  Box rows are walked 32 at a time, one row per lane; a chunk whose
        //      rows hold few generators is walked row-parallel (every lane its own row), a dense one
        //      row by row with all lanes.  The ring starts near the one that sufficed last time.
        //      While walking, the box generators are also staged in shared memory (if they fit), so
        //      that step 2 needs no second pass over global memory.
        double rbest;
        int ring = max(1, (ring_hint * 3) / 4);
        int xi0, xi1, xj0, xj1;
        int n_box;
        for (;;) {
            xi0 = max(ci0 - ring, 0); xi1 = min(ci1 + ring, g.gw - 1);
            xj0 = max(cj0 - ring, 0); xj1 = min(cj1 + ring, g.gh - 1);
            rbest = INFINITY;
            if (lane == 0) s_cnt[wib] = 0;
            __syncwarp();
            auto look = [&](int i) {
                const double2 p = cell_xy[i];
                double pv = 1.0;
                double d = tess_maxd2(r, p.x, p.y);
                if (HAS_SCALE) { pv = cell_inv[i]; d *= pv; }
                rbest = fmin(rbest, d);
                const int pos = atomicAdd(&s_cnt[wib], 1);
                if (pos < TESS_CAP) {
                    xy[pos] = p;
                    if (HAS_SCALE) iv[pos] = pv;
                    key[pos] = (unsigned)cell_items[i];      // generator index (independent, coalesced load)
                }
            };
#pragma unroll 1
            for (int j0 = xj0; j0 <= xj1; j0 += 32) {
                const int j = j0 + lane;
                int s = 0, e = 0;
                if (j <= xj1) {
                    s = cell_start[(long)j * g.gw + xi0];
                    e = cell_start[(long)j * g.gw + xi1 + 1];
                }
                const int rows = min(32, xj1 - j0 + 1);
                const int tot = __reduce_add_sync(TESS_FULL, e - s);
                if (tot <= 6 * rows) {
#pragma unroll 1
                    for (int i = s; i < e; ++i) look(i);
                    __syncwarp();
                } else {
#pragma unroll 1
                    for (int src = 0; src < rows; ++src) {
                        const int ss = __shfl_sync(TESS_FULL, s, src), ee = __shfl_sync(TESS_FULL, e, src);
#pragma unroll 1
                        for (int i = ss + lane; i < ee; i += 32) look(i);
                    }
                    __syncwarp();
                }
            }
            rbest = tess_warp_min(rbest);
            n_box = s_cnt[wib];
            __syncwarp();     // every lane has read the count before lane 0 resets it for the next ring
            const bool whole = (ci0 - ring <= 0 && cj0 - ring <= 0 && ci1 + ring >= g.gw - 1 && cj1 + ring >= g.gh - 1);
            if (whole) break;
            // every generator outside the box is at least `ring` whole cells away from the tile
            const double lb = fmax(g.cs * (double)ring - g.tol, 0.0);
            if (lb * lb * min_inv > rbest * TESS_SLACK) break;
            ring *= 2;
        }
        if (lane == 0) tile_ring[t] = ring;
        const double thr = rbest * TESS_SLACK;

        // ---- 2. keep every generator of the box with mind2 [* inv_s2] <= R (any order)
        int cnt;   // warp-uniform
        if (n_box <= TESS_CAP) {
            // in shared memory: order-free compaction in place (write position <= read position)
            cnt = 0;
#pragma unroll 1
            for (int b0 = 0; b0 < n_box; b0 += 32) {
                const int i = b0 + lane;
                double2 p = make_double2(0.0, 0.0);
                double pv = 1.0;
                unsigned kv = 0u;
                bool keep = false;
                if (i < n_box) {
                    p = xy[i];
                    kv = key[i];
                    double d = tess_mind2(r, p.x, p.y);
                    if (HAS_SCALE) { pv = iv[i]; d *= pv; }
                    keep = d <= thr;
                }
                const unsigned bb = __ballot_sync(TESS_FULL, keep);
                __syncwarp();
                if (keep) {
                    const int pos = cnt + __popc(bb & tess_lanemask_lt(lane));
                    xy[pos] = p;
                    key[pos] = kv;
                    if (HAS_SCALE) iv[pos] = pv;
                }
                cnt += __popc(bb);
                __syncwarp();
            }
        } else {
            // the box does not fit: second pass over the grid cells
            if (lane == 0) s_cnt[wib] = 0;
            __syncwarp();
#pragma unroll 1
            for (int j0 = xj0; j0 <= xj1; j0 += 32) {
                const int j = j0 + lane;
                int s = 0, e = 0;
                if (j <= xj1) {
                    s = cell_start[(long)j * g.gw + xi0];
                    e = cell_start[(long)j * g.gw + xi1 + 1];
                }
                const int rows = min(32, xj1 - j0 + 1);
                const int tot = __reduce_add_sync(TESS_FULL, e - s);
                auto take = [&](int i) {
                    const double2 p = cell_xy[i];
                    double d = tess_mind2(r, p.x, p.y);
                    double pv = 1.0;
                    if (HAS_SCALE) { pv = cell_inv[i]; d *= pv; }
                    if (d <= thr) {
                        const int pos = atomicAdd(&s_cnt[wib], 1);
                        if (pos < TESS_CAP) {
                            xy[pos] = p;
                            if (HAS_SCALE) iv[pos] = pv;
                            key[pos] = (unsigned)cell_items[i];
                        }
                    }
                };
                if (tot <= 6 * rows) {
#pragma unroll 1
                    for (int i = s; i < e; ++i) take(i);
                    __syncwarp();
                } else {
#pragma unroll 1
                    for (int src = 0; src < rows; ++src) {
                        const int ss = __shfl_sync(TESS_FULL, s, src), ee = __shfl_sync(TESS_FULL, e, src);
#pragma unroll 1
                        for (int i = ss + lane; i < ee; i += 32) take(i);
                    }
                    __syncwarp();
                }
            }
            __syncwarp();
            cnt = s_cnt[wib];
        }
        const bool overflow = cnt > TESS_CAP;

        // ---- 4. the four quadrants: one pass for the four centre minima, one for the keep flags
        //      (entry i = 32 b + lane -> bit b of the lane's flag word), one arena reservation for
        //      the tile, then the writes
        const bool whole_tile = ((long)(tx + 1) * TESS_TS <= T.W) && ((long)(ty + 1) * TESS_TS <= T.H);
        double qcx[4], qcy[4], qthr[4];
        bool qin[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const long px0 = (long)tx * TESS_TS + 32 * (q & 1), py0 = (long)ty * TESS_TS + 32 * (q >> 1);
            const long px1 = min(px0 + 31, T.W - 1), py1 = min(py0 + 31, T.H - 1);
            qin[q] = (px0 < T.W) && (py0 < T.H);
            qcx[q] = T.x0 + 0.5 * (double)(px0 + px1);
            qcy[q] = T.y0 + 0.5 * (double)(py0 + py1);
            const double hx = 0.5 * (double)(px1 - px0), hy = 0.5 * (double)(py1 - py0);
            qthr[q] = whole_tile ? 21.920310216783 : sqrt(fmax(hx * hx + hy * hy, 0.0)) * 1.0000001;   // rho (15.5 sqrt 2)
        }
        unsigned bits[4] = {0u, 0u, 0u, 0u};
        int nq_[4] = {0, 0, 0, 0};
        if (!overflow && cnt > TESS_LEAF_MIN_BUILD) {
            double umin[4] = {INFINITY, INFINITY, INFINITY, INFINITY};
            double imax = 1.0;
#pragma unroll 1
            for (int i = lane; i < cnt; i += 32) {
                const double2 p = xy[i];
                double v = 1.0;
                if (HAS_SCALE) { v = iv[i]; imax = fmax(imax, v); }
                // the centre's x depends on the quadrant's column only, its y on the row only
                const double ax = qcx[0] - p.x, bx = qcx[1] - p.x, ay = qcy[0] - p.y, by = qcy[2] - p.y;
                const double axx = ax * ax, bxx = bx * bx, ayy = ay * ay, byy = by * by;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    double d = ((q & 1) ? bxx : axx) + ((q >> 1) ? byy : ayy);
                    if (HAS_SCALE) d *= v;
                    umin[q] = fmin(umin[q], d);
                }
            }
            double cs = 1.0;
            if (HAS_SCALE) {
                imax = -tess_warp_min(-imax);
                cs = (double)sqrtf((float)imax) * 1.000001 + 1e-18;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const double um = tess_warp_min(umin[q]);
                const double c = 2.0 * qthr[q] * cs;
                const double su = (double)sqrtf((float)um) * 1.000001 + 1e-18;      // >= sqrt(umin)
                qthr[q] = (um + 2.0 * c * su + c * c) * TESS_SLACK;
            }
#pragma unroll 1
            for (int b = 0; 32 * b < cnt; ++b) {
                const int i = 32 * b + lane;
                double2 p = make_double2(0.0, 0.0);
                double v = 1.0;
                if (i < cnt) { p = xy[i]; if (HAS_SCALE) v = iv[i]; }
                const double ax = qcx[0] - p.x, bx = qcx[1] - p.x, ay = qcy[0] - p.y, by = qcy[2] - p.y;
                const double axx = ax * ax, bxx = bx * bx, ayy = ay * ay, byy = by * by;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    double d = ((q & 1) ? bxx : axx) + ((q >> 1) ? byy : ayy);
                    if (HAS_SCALE) d *= v;
                    const bool keep = (i < cnt) && qin[q] && (d <= qthr[q]);
                    if (keep) bits[q] |= 1u << b;
                    nq_[q] += __popc(__ballot_sync(TESS_FULL, keep));
                }
            }
        } else if (!overflow) {
            // a handful of candidates: every quadrant takes them all
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (qin[q]) { bits[q] = (lane < cnt) ? 1u : 0u; nq_[q] = cnt; }
            }
        }
        const int n_all = nq_[0] + nq_[1] + nq_[2] + nq_[3];
        unsigned off = 0;
        int ok = overflow ? 0 : 1;
        if (lane == 0 && ok) {
            off = atomicAdd(&st->list_cursor, (unsigned)n_all);
            if (off + (unsigned)n_all > L.capacity) ok = 0;
        }
        off = __shfl_sync(TESS_FULL, off, 0);
        ok = __shfl_sync(TESS_FULL, ok, 0);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const long qid = (long)(2 * ty + (q >> 1)) * L.nqx + (2 * tx + (q & 1));
            if (lane == 0) {
                L.q_off[qid] = ok ? (int)off : 0;
                L.q_cnt[qid] = !qin[q] ? 0 : (ok ? nq_[q] : -1);
                // off the fast path of the image kernel (overflow, or a list longer than it stages):
                // queued for its tail phase
                if (qin[q] && (!ok || nq_[q] > qcap)) {
                    L.ring_q[atomicAdd(&st->n_brute, 1)] = (int)qid;
                    if (!ok) atomicAdd(&st->n_overflow, 1);     // a true overflow (not merely a long list)
                }
            }
            if (ok && qin[q]) {
                int wpos = 0;
#pragma unroll 1
                for (int b = 0; 32 * b < cnt; ++b) {
                    const bool keep = (bits[q] >> b) & 1u;
                    const unsigned bb = __ballot_sync(TESS_FULL, keep);
                    if (keep) {
                        const int i = 32 * b + lane;
                        const unsigned o = off + wpos + __popc(bb & tess_lanemask_lt(lane));
                        L.xy[o] = xy[i];
                        L.gid[o] = (int)key[i];
                        if (HAS_SCALE) L.inv[o] = iv[i];
                    }
                    wpos += __popc(bb);
                }
                off += (unsigned)nq_[q];
            }
        }
        __syncwarp();
    }
}
