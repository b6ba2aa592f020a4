// tess_image.cuh -- the fused assignment + segmented-reduction kernel for pixel grids.
//
// Replaces, per Lloyd iteration, `tree.query(xy, k=1)[1]` (reference tess/lloyd.pyx:55) and the
// sequential accumulate loop (tess/lloyd.pyx:58-65) for the point set that
// CVTessellation.from_image builds (tess/voronoi.py:281-287); in MODE 0 it is the nearest-
// generator rasterisation behind VoronoiTessellation.segmap (tess/voronoi.py:75-114).
//
// One persistent CTA (4 warps) pulls 64x64 super-tiles from a dynamic counter.  Per super-tile
// the candidate list (k_build_lists) is staged in shared memory; each warp walks four 32x8
// warp tiles, narrowing the list twice (warp tile, then 8x8 leaf) with the exact
// mind2 <= min maxd2 test, evaluates the surviving candidates with the pinned fp64 formula,
// writes int32 labels (one coalesced 128-byte line per row) and folds the pixel's moments into
// a two-entry most-recently-used cache held in registers.  A cache entry is flushed to the
// global moment table with fire-and-forget fp64 reductions when a third label shows up in the
// lane's pixel strip; at the end of every super-tile the lanes of a warp that hold the same bin
// are merged with shuffles and one reduction per (warp, bin) is issued.
//
// Algorithmic HBM traffic per pixel: MODE 1: 8 B (density) + 4 B (label) = 12 B;
// MODE 2: 8 + 8 B (signal, noise) + 4 B = 20 B; MODE 0: 4 B.  Pixel coordinates are implicit.
#pragma once
#include "tess_common.cuh"

struct TessImgArgs {
    const double* dens;    // MODE 1: [H][W] weights
    const double* sig;     // MODE 2
    const double* noise;   // MODE 2
    int32_t* labels_ext;   // [H][W] out (caller memory, MODE 0) or null: use lab0/lab1 by iteration parity
    int32_t* lab0;
    int32_t* lab1;
    double* mom;           // [k][M] global moment table (M = 4 or 6)
    const double* node;    // [k][2]
    const double* inv_s2;  // [k] or null
    const int* list_off;
    const int* list_cnt;
    const int* list_items;
    TessState* st;
    long H, W;
    double x0, y0;
    int tiles_x, tiles_y;
    int k;
    int uniform_weight;    // MODE 2: w = 1 instead of (S/N)^2
    int honor;             // honour the convergence latch (iteration kernels) or not (segmap)
};

// Two-entry MRU cache of per-bin partial moments, in registers.
template <int NV>
struct TessEntry {
    int id;      // generator index, -1 = empty
    int n;       // pixel count
    double v[NV];  // sum w, sum w*x, sum w*y [, sum S, sum N^2]
};

template <int NV>
TESS_DEVFN void tess_entry_clear(TessEntry<NV>& e) {
    e.id = -1;
    e.n = 0;
#pragma unroll
    for (int i = 0; i < NV; ++i) e.v[i] = 0.0;
}

// fire-and-forget fp64 reductions into the global table (REDG.E.ADD.F64 on sm_100a)
template <int NV>
TESS_DEVFN void tess_entry_flush(const TessEntry<NV>& e, double* mom) {
    if (e.n > 0) {
        double* m = mom + (long)e.id * (NV + 1);
        atomicAdd(m, (double)e.n);
#pragma unroll
        for (int i = 0; i < NV; ++i) atomicAdd(m + 1 + i, e.v[i]);
    }
}

// End-of-kernel flush: lanes holding the same bin are summed with shuffles first, so a bin that
// spans many warps receives one reduction per warp instead of one per lane.
template <int NV>
TESS_DEVFN void tess_entry_flush_warp(const TessEntry<NV>& e, double* mom, int lane) {
    int id = (e.n > 0) ? e.id : -1;
    unsigned todo = __ballot_sync(TESS_FULL, id >= 0);
    while (todo) {
        int leader = __ffs((int)todo) - 1;
        int lid = __shfl_sync(TESS_FULL, id, leader);
        bool mine = (id == lid);
        unsigned grp = __ballot_sync(TESS_FULL, mine);
        double cnt = tess_warp_sum(mine ? (double)e.n : 0.0);
        double s[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) s[i] = tess_warp_sum(mine ? e.v[i] : 0.0);
        if (lane == leader) {
            double* m = mom + (long)lid * (NV + 1);
            atomicAdd(m, cnt);
#pragma unroll
            for (int i = 0; i < NV; ++i) atomicAdd(m + 1 + i, s[i]);
        }
        todo &= ~grp;
    }
}

// Fold one pixel into the cache.  A is the most recently used entry.
template <int NV>
TESS_DEVFN void tess_cache_add(TessEntry<NV>& A, TessEntry<NV>& B, double* mom, int id, const double (&val)[NV]) {
    if (id != A.id) {
        if (id == B.id) {
            TessEntry<NV> t = A;
            A = B;
            B = t;
        } else {
            tess_entry_flush<NV>(B, mom);
            B = A;
            tess_entry_clear<NV>(A);
            A.id = id;
        }
    }
    A.n += 1;
#pragma unroll
    for (int i = 0; i < NV; ++i) A.v[i] = __dadd_rn(A.v[i], val[i]);
}

// bits of a full-warp ballot that belong to leaf `leaf` (lanes 4*leaf..4*leaf+3 and +16),
// packed in `sub` order (sub = (lane&3) | ((lane>>4)<<2))
TESS_DEVFN unsigned tess_leaf_bits(unsigned b, int leaf) {
    return ((b >> (4 * leaf)) & 0xFu) | (((b >> (16 + 4 * leaf)) & 0xFu) << 4);
}

// MODE 0: labels only; 1: density map (4 moments); 2: signal + noise maps (6 moments)
template <int MODE, bool HAS_SCALE, bool VEC>
__global__ void __launch_bounds__(TESS_MAIN_THREADS)
k_image_main(TessImgArgs a) {
    constexpr int NV = (MODE == 2) ? 5 : 3;
    if (TESS_STOPPED(a.st, a.honor)) return;
    int32_t* const labels = a.labels_ext ? a.labels_ext : ((a.st->iters_done & 1) ? a.lab1 : a.lab0);

    __shared__ __align__(16) double2 s_g[TESS_CAP];    // generator (x, y) of each list slot
    __shared__ int s_gid[TESS_CAP];                    // its global index
    __shared__ double s_inv[HAS_SCALE ? TESS_CAP : 1]; // 1/scale^2
    __shared__ unsigned short s_wl[TESS_MAIN_THREADS / 32][TESS_CAP];               // warp-tile lists
    __shared__ unsigned short s_ll[TESS_MAIN_THREADS / 32][4][TESS_LEAF_CAP];       // leaf lists
    __shared__ int s_tile;

    const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
    const int cx = lane & 15, ry = lane >> 4;          // patch: cols 2cx,2cx+1; rows 4ry..4ry+3
    const int leaf = cx >> 2;                          // 8-px-wide leaf of the warp tile
    const int sub = (lane & 3) | (ry << 2);            // rank inside the leaf's 8 lanes
    unsigned short* wl = s_wl[wib];
    unsigned short* ll = s_ll[wib][leaf];
    const int n_tiles = a.tiles_x * a.tiles_y;

    TessEntry<NV> A, B;
    tess_entry_clear<NV>(A);
    tess_entry_clear<NV>(B);

    for (;;) {
        if (tid == 0) s_tile = atomicAdd(&a.st->tile_counter, 1);
        __syncthreads();
        const int tile = s_tile;
        if (tile >= n_tiles) break;
        const int tx = tile % a.tiles_x, ty = tile / a.tiles_x;
        const int cnt_raw = a.list_cnt[tile];
        const bool brute = (cnt_raw == TESS_BRUTE);
        const int n_chunks = brute ? (a.k + TESS_CAP - 1) / TESS_CAP : 1;

        // warp-tile origin inside the super-tile: x half = wib&1, 32-row half = wib>>1, 4 steps down
        const long wx0 = (long)tx * TESS_TS + 32 * (wib & 1);
        const long wyb = (long)ty * TESS_TS + 32 * (wib >> 1);

        if (!brute) {
            const int off = a.list_off[tile];
            for (int i = tid; i < cnt_raw; i += TESS_MAIN_THREADS) {
                int gid = a.list_items[off + i];
                s_gid[i] = gid;
                s_g[i] = make_double2(a.node[2 * (long)gid], a.node[2 * (long)gid + 1]);
                if (HAS_SCALE) s_inv[i] = a.inv_s2[gid];
            }
            __syncthreads();
        }

        for (int step = 0; step < 4; ++step) {
            const long wy0 = wyb + 8 * step;
            const bool wt_in = (wx0 < a.W) && (wy0 < a.H);   // warp-uniform
            const long px = wx0 + 2 * cx;                    // first column of this lane
            const long py = wy0 + 4 * ry;                    // first row of this lane
            const double X0 = a.x0 + (double)px, X1 = a.x0 + (double)(px + 1);

            // ---- issue the pixel loads early; they are consumed after the candidate evaluation
            double w_[4][2], s_[4][2], n_[4][2];
            if (MODE != 0) {
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const bool rin = wt_in && (py + r < a.H);
                    const long base = (py + r) * a.W + px;
#pragma unroll
                    for (int c = 0; c < 2; ++c) { w_[r][c] = 0.0; s_[r][c] = 0.0; n_[r][c] = 0.0; }
                    if (VEC) {
                        if (rin && px < a.W) {
                            if (MODE == 1) {
                                double2 v = *reinterpret_cast<const double2*>(a.dens + base);
                                w_[r][0] = v.x; w_[r][1] = v.y;
                            } else {
                                double2 v = *reinterpret_cast<const double2*>(a.sig + base);
                                double2 u = *reinterpret_cast<const double2*>(a.noise + base);
                                s_[r][0] = v.x; s_[r][1] = v.y; n_[r][0] = u.x; n_[r][1] = u.y;
                            }
                        }
                    } else {
#pragma unroll
                        for (int c = 0; c < 2; ++c)
                            if (rin && px + c < a.W) {
                                if (MODE == 1) w_[r][c] = a.dens[base + c];
                                else { s_[r][c] = a.sig[base + c]; n_[r][c] = a.noise[base + c]; }
                            }
                    }
                }
            }

            // ---- best (distance, index) per pixel
            double bd[4][2];
            int bi[4][2];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 2; ++c) { bd[r][c] = INFINITY; bi[r][c] = 0x7fffffff; }

            for (int chunk = 0; chunk < n_chunks; ++chunk) {
                int cnt = cnt_raw;
                if (brute) {   // exhaustive fallback: stage generators [chunk*CAP, ...) ; CTA-uniform
                    __syncthreads();
                    const int g0 = chunk * TESS_CAP;
                    cnt = min(TESS_CAP, a.k - g0);
                    for (int i = tid; i < cnt; i += TESS_MAIN_THREADS) {
                        s_gid[i] = g0 + i;
                        s_g[i] = make_double2(a.node[2 * (long)(g0 + i)], a.node[2 * (long)(g0 + i) + 1]);
                        if (HAS_SCALE) s_inv[i] = a.inv_s2[g0 + i];
                    }
                    __syncthreads();
                }

                // ---- level 1: super-tile list -> warp-tile list (wl, slots into s_g)
                int c_w;
                bool wl_identity;   // true: the warp list is the staged list itself (slots 0..cnt-1)
                if (cnt == 1 || brute || !wt_in) {
                    c_w = cnt;
                    wl_identity = true;
                } else {
                    TessRect wr;
                    wr.x0 = a.x0 + (double)wx0;
                    wr.x1 = a.x0 + (double)min(wx0 + TESS_WT_W - 1, a.W - 1);
                    wr.y0 = a.y0 + (double)wy0;
                    wr.y1 = a.y0 + (double)min(wy0 + TESS_WT_H - 1, a.H - 1);
                    double rmin = INFINITY;
                    for (int i = lane; i < cnt; i += 32) {
                        double d = tess_maxd2(wr, s_g[i].x, s_g[i].y);
                        if (HAS_SCALE) d *= s_inv[i];
                        rmin = fmin(rmin, d);
                    }
                    rmin = tess_warp_min(rmin);
                    const double thr = rmin * TESS_SLACK;
                    c_w = 0;
                    for (int b0 = 0; b0 < cnt; b0 += 32) {
                        int i = b0 + lane;
                        bool keep = false;
                        if (i < cnt) {
                            double d = tess_mind2(wr, s_g[i].x, s_g[i].y);
                            if (HAS_SCALE) d *= s_inv[i];
                            keep = d <= thr;
                        }
                        unsigned b = __ballot_sync(TESS_FULL, keep);
                        if (keep) wl[c_w + __popc(b & tess_lanemask_lt(lane))] = (unsigned short)i;
                        c_w += __popc(b);
                    }
                    wl_identity = false;
                    __syncwarp();
                }

                // ---- level 2: warp-tile list -> leaf list (ll), per 8-lane group
                const unsigned short* lst = wl;
                int len = c_w;
                bool identity = wl_identity;
                if (!wl_identity && c_w > 1) {
                    TessRect lr;
                    const long lx0 = wx0 + 8 * leaf;
                    lr.x0 = a.x0 + (double)lx0;
                    lr.x1 = a.x0 + (double)min(lx0 + 7, a.W - 1);
                    lr.y0 = a.y0 + (double)wy0;
                    lr.y1 = a.y0 + (double)min(wy0 + TESS_WT_H - 1, a.H - 1);
                    double rmin = INFINITY;
                    for (int i = sub; i < c_w; i += 8) {
                        int sl = wl[i];
                        double d = tess_maxd2(lr, s_g[sl].x, s_g[sl].y);
                        if (HAS_SCALE) d *= s_inv[sl];
                        rmin = fmin(rmin, d);
                    }
                    rmin = fmin(rmin, __shfl_xor_sync(TESS_FULL, rmin, 1));
                    rmin = fmin(rmin, __shfl_xor_sync(TESS_FULL, rmin, 2));
                    rmin = fmin(rmin, __shfl_xor_sync(TESS_FULL, rmin, 16));
                    const double thr = rmin * TESS_SLACK;
                    int c_l = 0;
                    for (int b0 = 0; b0 < c_w; b0 += 8) {
                        int i = b0 + sub;
                        bool keep = false;
                        int sl = 0;
                        if (i < c_w) {
                            sl = wl[i];
                            double d = tess_mind2(lr, s_g[sl].x, s_g[sl].y);
                            if (HAS_SCALE) d *= s_inv[sl];
                            keep = d <= thr;
                        }
                        unsigned gb = tess_leaf_bits(__ballot_sync(TESS_FULL, keep), leaf);
                        int pos = c_l + __popc(gb & ((1u << sub) - 1u));
                        if (keep && pos < TESS_LEAF_CAP) ll[pos] = (unsigned short)sl;
                        c_l += __popc(gb);
                    }
                    __syncwarp();
                    if (c_l <= TESS_LEAF_CAP) { lst = ll; len = c_l; }
                }

                // ---- evaluate the surviving candidates with the pinned formula
                if (wt_in) {
                    if (len == 1 && !brute) {
                        const int sl = identity ? 0 : lst[0];
                        const int gid = s_gid[sl];
#pragma unroll
                        for (int r = 0; r < 4; ++r) { bi[r][0] = gid; bi[r][1] = gid; }
                    } else {
                        for (int j = 0; j < len; ++j) {
                            const int sl = identity ? j : lst[j];
                            const double2 g = s_g[sl];
                            const int gid = s_gid[sl];
                            double inv = 1.0;
                            if (HAS_SCALE) inv = s_inv[sl];
                            const double dx0 = __dsub_rn(X0, g.x), dx1 = __dsub_rn(X1, g.x);
                            const double dxx0 = __dmul_rn(dx0, dx0), dxx1 = __dmul_rn(dx1, dx1);
#pragma unroll
                            for (int r = 0; r < 4; ++r) {
                                const double Y = a.y0 + (double)(py + r);
                                const double dy = __dsub_rn(Y, g.y);
                                const double dyy = __dmul_rn(dy, dy);
                                double d0 = __dadd_rn(dxx0, dyy), d1 = __dadd_rn(dxx1, dyy);
                                if (HAS_SCALE) { d0 = __dmul_rn(d0, inv); d1 = __dmul_rn(d1, inv); }
                                if ((d0 < bd[r][0]) | ((d0 == bd[r][0]) & (gid < bi[r][0]))) { bd[r][0] = d0; bi[r][0] = gid; }
                                if ((d1 < bd[r][1]) | ((d1 == bd[r][1]) & (gid < bi[r][1]))) { bd[r][1] = d1; bi[r][1] = gid; }
                            }
                        }
                    }
                }
            }   // chunks

            // ---- labels out: 16 lanes x 8 B = one 128-byte line per row
            if (wt_in) {
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    if (py + r < a.H) {
                        const long base = (py + r) * a.W + px;
                        if (VEC) {
                            if (px < a.W) *reinterpret_cast<int2*>(labels + base) = make_int2(bi[r][0], bi[r][1]);
                        } else {
                            if (px < a.W) labels[base] = bi[r][0];
                            if (px + 1 < a.W) labels[base + 1] = bi[r][1];
                        }
                    }
                }
            }

            // ---- moments: fold the lane's 8 pixels into its register cache
            if (MODE != 0 && wt_in) {
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const double Y = a.y0 + (double)(py + r);
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        const bool in = (py + r < a.H) && (px + c < a.W);
                        const double X = c ? X1 : X0;
                        double val[NV];
                        bool ok;
                        if (MODE == 1) {
                            const double w = w_[r][c];
                            ok = in && isfinite(w);
                            val[0] = w;
                            val[1] = __dmul_rn(w, X);
                            val[2] = __dmul_rn(w, Y);
                        } else {
                            const double S = s_[r][c], N = n_[r][c];
                            double w = 1.0;
                            if (!a.uniform_weight) {
                                const double q = __ddiv_rn(S, N);
                                w = __dmul_rn(q, q);
                            }
                            ok = in && isfinite(w) && isfinite(S) && isfinite(N);
                            val[0] = w;
                            val[1] = __dmul_rn(w, X);
                            val[2] = __dmul_rn(w, Y);
                            val[3] = S;
                            val[4] = __dmul_rn(N, N);
                        }
                        if (ok) tess_cache_add<NV>(A, B, a.mom, bi[r][c], val);
                    }
                }
            }
        }   // steps

        // ---- end of super-tile: lanes holding the same bin are merged with shuffles, one
        // reduction per (warp, bin) goes to the global table
        if (MODE != 0) {
            tess_entry_flush_warp<NV>(A, a.mom, lane);
            tess_entry_flush_warp<NV>(B, a.mom, lane);
            tess_entry_clear<NV>(A);
            tess_entry_clear<NV>(B);
        }
    }       // tiles
}

// ---------------------------------------------------------------------------------------------
// Exhaustive O(n*k) assignment, validation aid only (include/tess_b200.h).  One thread per pixel
// or point; generators streamed through shared memory in chunks.
template <bool HAS_SCALE, bool GRID>
__global__ void __launch_bounds__(256)
k_assign_exhaustive(const double* node, const double* inv_s2, int k, long n, long W, double x0, double y0,
                    const double* xy, int32_t* out) {
    __shared__ __align__(16) double2 s_g[512];
    __shared__ double s_inv[HAS_SCALE ? 512 : 1];
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    double px = 0.0, py = 0.0;
    if (i < n) {
        if (GRID) { px = x0 + (double)(i % W); py = y0 + (double)(i / W); }
        else { px = xy[2 * i]; py = xy[2 * i + 1]; }
    }
    double bd = INFINITY;
    int bi = 0x7fffffff;
    for (int g0 = 0; g0 < k; g0 += 512) {
        const int cnt = min(512, k - g0);
        __syncthreads();
        for (int j = threadIdx.x; j < cnt; j += blockDim.x) {
            s_g[j] = make_double2(node[2 * (long)(g0 + j)], node[2 * (long)(g0 + j) + 1]);
            if (HAS_SCALE) s_inv[j] = inv_s2[g0 + j];
        }
        __syncthreads();
        for (int j = 0; j < cnt; ++j) {
            double d = tess_d2_pinned(px, py, s_g[j].x, s_g[j].y);
            if (HAS_SCALE) d = __dmul_rn(d, s_inv[j]);
            if (d < bd) { bd = d; bi = g0 + j; }   // ascending j: strict '<' keeps the lowest index
        }
    }
    if (i < n) out[i] = bi;
}
