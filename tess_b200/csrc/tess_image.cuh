// tess_image.cuh -- the fused assignment + segmented-reduction kernel for pixel grids.
//
// Replaces, per Lloyd iteration, `tree.query(xy, k=1)[1]` (reference tess/lloyd.pyx:55) and the
// sequential accumulate loop (tess/lloyd.pyx:58-65) for the point set that
// CVTessellation.from_image builds (tess/voronoi.py:281-287); in MODE 0 it is the nearest-
// generator rasterisation behind VoronoiTessellation.segmap (tess/voronoi.py:75-114).
//
// Work unit: one WARP per 32x32-px quadrant, warps fully independent (no CTA barrier anywhere).
// A warp's first chunk of quadrants is fixed (its global warp index), the following ones are drawn
// from a device queue (one atomic per chunk of TESS_CHUNK quadrants, requested a whole chunk before
// its value is needed and never touched in between), so no warp idles while others still have work.
// The quadrant's candidate list (k_build_qlists: arena records in no particular order) is copied
// into the warp's shared-memory buffer with cp.async one quadrant AHEAD, so list latency is hidden
// behind the previous quadrant's arithmetic; the next quadrant's pixel tile is requested from DRAM
// into L2 by one TMA prefetch.
// Once per quadrant the candidate sets of its sixteen 8x8 leaves are formed -- lane pair (2L, 2L+1)
// measures the list (lists longer than 64: the list of the leaf's 16x16 block) against leaf L's centre
// with the centre + radius test
//     keep g  iff  u(c,g) <= min_h u(c,h) + 2 rho / s_min        (u = distance / scale)
// which can never drop the true nearest generator of any pixel of the rectangle -- and kept as 64-bit
// masks over the list.  For each of its four 32x8 warp tiles the warp then evaluates the set bits of
// every lane's leaf mask with the pinned fp64 formula (an exact tie is detected and then resolved to
// the lowest generator index) and writes labels as one coalesced 128-byte line per row.  A quadrant
// whose leaves all end with the same single candidate is one bin: constant labels, closed-form sums.
//
// Moments ("segmented reduction"): every lane owns a 2-column x 4-row pixel patch per warp tile and
// keeps ONE running same-bin sum in registers, keyed by generator id, across warp tiles AND
// quadrants (labels are spatially coherent, so most patches extend the current run); a patch is
// summed in patch-local coordinates and shifted once.  When a lane's run ends, its sums go straight
// to the bin's row of the global moment vector as fire-and-forget fp64 reductions; when many lanes
// retire at once (a warp leaving a large bin) they are merged with a shuffle tree first, because
// same-address reductions serialise in L2.  The fold is written without lane-divergent branches.
//
// Tail phase: quadrants off this fast path (list longer than the staged capacity, or a tile that
// overflowed the builder) are queued by the builder; after their own quadrants all warps of the grid take
// 2-row strips of them (list scan from L2, or the per-pixel ring search over the generator grid).
//
// Algorithmic HBM traffic per pixel: MODE 1: 8 B (density) + 4 B (label) = 12 B;
// MODE 2: 8 + 8 B (signal, noise) + 4 B = 20 B; MODE 0: 4 B.  Pixel coordinates are implicit.
#pragma once
#include "tess_common.cuh"
#include "tess_index.cuh"
#include "tess_points.cuh"

#ifndef TESS_BCAP
#define TESS_BCAP 64        // 16x16 block list capacity
#endif
#ifndef TESS_BLOCK_MIN
#define TESS_BLOCK_MIN 64   // quadrant lists longer than this (the leaf masks hold 64 entries) get the intermediate 16x16 level
#endif
#ifndef TESS_LEAF_MIN
#define TESS_LEAF_MIN 3     // parent lists up to this length are evaluated without a leaf filter
#endif
// quadrant list capacity of the fast path (records staged in the warp's shared memory)
#define TESS_QCAP(MODE) 128
#ifndef TESS_CHUNK
#define TESS_CHUNK 2        // consecutive quadrants per draw from the device queue (C3 / C4 kernel: 1: 0.233 / 3.88 ms, 2: 0.234 / 3.06, 4: 0.247 / 3.07)
#endif
#ifndef TESS_RING_ROWS
#define TESS_RING_ROWS 2
#endif
#ifndef TESS_MERGE_ENTER
#define TESS_MERGE_ENTER 16 // retiring lanes from which the same-bin groups are merged with shuffles first
#endif
#ifndef TESS_MERGE_MIN
#define TESS_MERGE_MIN 12   // smallest group worth a shuffle tree
#endif
#ifndef TESS_EARLY_LOAD
#define TESS_EARLY_LOAD 0   // 1: the next warp tile's pixels are requested before the fold of the current one
#endif
#ifndef TESS_PREFETCH
#define TESS_PREFETCH 1     // the next quadrant's pixel tile is prefetched into L2 by the TMA unit
#endif

struct TessImgArgs {
    const double* dens;    // MODE 1: [H][W] weights
    const double* sig;     // MODE 2
    const double* noise;   // MODE 2
    int32_t* labels_ext;   // [H][W] out (caller memory, MODE 0) or null: use lab0/lab1 by iteration parity
    int32_t* lab0;
    int32_t* lab1;
    double* mom;           // [k][M] global moment table (M = 4 or 6)
    TessQLists L;          // per-quadrant candidate lists
    TessGrid g;            // generator grid (ring-search fallback of overflowed tiles)
    const int* cell_start;
    const int* cell_items;
    const double2* cell_xy;
    const double* cell_inv;
    TessState* st;
    long H, W;
    double x0, y0;
    int uniform_weight;    // MODE 2: w = 1 instead of (S/N)^2
    int honor;             // honour the convergence latch (iteration kernels) or not (segmap)
};

// A run of same-bin pixels being summed by one lane (registers).
// v = sum w, sum w*x, sum w*y [, sum S, sum N^2]
template <int NV>
struct TessRun {
    int key;    // generator id of the bin being summed; -1 = none yet
    int n;      // pixel count
    double v[NV];
};

TESS_DEVFN void tess_gmem_add(double* p, double v) {
#ifndef TESS_EMU
    asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
#else
    atomicAdd(p, v);
#endif
}

// 16 / 8 / 4-byte asynchronous global -> shared copies (LDGSTS); plain copies under the emulator
TESS_DEVFN void tess_cp_async16(void* dst, const void* src) {
#ifndef TESS_EMU
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
#else
    memcpy(dst, src, 16);
#endif
}
TESS_DEVFN void tess_cp_async8(void* dst, const void* src) {
#ifndef TESS_EMU
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
#else
    memcpy(dst, src, 8);
#endif
}
TESS_DEVFN void tess_cp_async4(void* dst, const void* src) {
#ifndef TESS_EMU
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
#else
    memcpy(dst, src, 4);
#endif
}
// TMA prefetch of the 32x32-px tile whose first pixel is (col, row) into L2: ONE instruction for the
// whole 8 KB tile (cp.async.bulk.prefetch.tensor -> UTMAPF), no register or shared-memory destination,
// no per-row address arithmetic.  Issued by one lane.
TESS_DEVFN void tess_prefetch_tile(const TessTileMap* tm, int col, int row) {
#ifndef TESS_EMU
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(tm), "r"(col), "r"(row) : "memory");
#else
    (void)tm; (void)col; (void)row;
#endif
}
TESS_DEVFN void tess_cp_async_wait() {
#ifndef TESS_EMU
    asm volatile("cp.async.wait_all;" ::: "memory");
#endif
}

// Lanes with `flush` set retire their run: its pixel count and sums go straight to the bin's row of
// the global moment vector as fire-and-forget fp64 reductions (red.global.add.f64; the moment table
// is a few MB and lives in L2).  Runs are keyed by GENERATOR id, so a run outlives the quadrant it
// started in: a lane inside a large bin keeps adding to its registers quadrant after quadrant and
// issues no reduction at all until its bin changes or the kernel ends.  No shared-memory table, no
// election between lanes, no divergence beyond the predicate.
template <int NV>
TESS_DEVFN void tess_run_flush(bool flush, const TessRun<NV>& cur, double* mom, int lane) {
    // Many lanes retiring at once usually retire the SAME bin (a warp that leaves a large bin): their
    // sums are merged with shuffles first and one lane issues the reductions -- same-address fp64
    // reductions serialise in L2, 32 lanes x thousands of warps on a handful of rows is a hot spot.
    // A handful of retiring lanes (a bin boundary inside the warp tile) go straight to L2.
    unsigned todo = __ballot_sync(TESS_FULL, flush);
    if (__popc(todo) >= TESS_MERGE_ENTER) {
#pragma unroll 1
        for (int turn = 0; turn < 2; ++turn) {
            const int leader = __ffs((int)todo) - 1;
            const int key = __shfl_sync(TESS_FULL, cur.key, leader);
            const bool mine = flush && (cur.key == key);
            const unsigned grp = __ballot_sync(TESS_FULL, mine);
            if (__popc(grp) < TESS_MERGE_MIN) break;             // small groups: every lane goes direct below
            int n = mine ? cur.n : 0;
            double v[NV];
#pragma unroll
            for (int i = 0; i < NV; ++i) v[i] = mine ? cur.v[i] : 0.0;
#pragma unroll
            for (int m = 16; m >= 1; m >>= 1) {
                n += __shfl_xor_sync(TESS_FULL, n, m);
#pragma unroll
                for (int i = 0; i < NV; ++i) v[i] = __dadd_rn(v[i], __shfl_xor_sync(TESS_FULL, v[i], m));
            }
            if (lane == leader) {
                double* m = mom + (long)key * (NV + 1);
                tess_gmem_add(m, (double)n);
#pragma unroll
                for (int i = 0; i < NV; ++i) tess_gmem_add(m + 1 + i, v[i]);
            }
            flush = flush && !mine;
            todo &= ~grp;
            if (__popc(todo) < TESS_MERGE_MIN) break;
        }
    }
    if (flush) {
        double* m = mom + (long)cur.key * (NV + 1);
        tess_gmem_add(m, (double)cur.n);
#pragma unroll
        for (int i = 0; i < NV; ++i) tess_gmem_add(m + 1 + i, cur.v[i]);
    }
}

// The lane's run moves to bin `key` where `sw` is set (after its old contents were flushed).
template <int NV>
TESS_DEVFN void tess_run_switch(bool sw, int key, TessRun<NV>& cur) {
    cur.key = sw ? key : cur.key;
    cur.n = sw ? 0 : cur.n;
#pragma unroll
    for (int i = 0; i < NV; ++i) cur.v[i] = sw ? 0.0 : cur.v[i];
}

// upper bound of sqrt(x) for x >= 0 from the fp32 unit (exact value is not needed, only >=)
TESS_DEVFN double tess_sqrt_up(double x) {
    return (double)sqrtf((float)x) * 1.000001 + 1e-18;
}

// squared threshold of the centre + radius test: (sqrt(umin) + c)^2 with slack, c = 2 rho / s_min
TESS_DEVFN double tess_center_thr(double umin, double c) {
    return (umin + 2.0 * c * tess_sqrt_up(umin) + c * c) * TESS_SLACK;
}

// A strip (TESS_RING_ROWS rows x 32 columns, one column per lane) of a quadrant whose candidate list
// is longer than the fast path stages in shared memory: every pixel scans the whole list (n records
// at gxy / gids / ginv in global memory, i.e. L2; four records per trip so that the loads overlap)
// and sends its moments straight to the global vector.  Part of the tail phase of k_image_main.
template <int MODE, bool HAS_SCALE>
TESS_NOINLINE void tess_rows_scan(const TessImgArgs& a, int32_t* labels, long qx0, long r0, const double2* gxy,
                                  const int* gids, const double* ginv, int n) {
    constexpr int NV = (MODE == 2) ? 5 : 3;
    constexpr int R = TESS_RING_ROWS;
    const int lane = threadIdx.x & 31;
    const long c = qx0 + lane;
    const double X = a.x0 + (double)c;
    double Y[R], bd[R];
    int bi[R];
#pragma unroll
    for (int j = 0; j < R; ++j) { Y[j] = a.y0 + (double)(r0 + j); bd[j] = INFINITY; bi[j] = 0x7fffffff; }
    for (int i = 0; i < n; i += 4) {
        double2 g[4];
        double iv[4];
        int gid[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int ii = min(i + u, n - 1);
            g[u] = gxy[ii];
            iv[u] = HAS_SCALE ? ginv[ii] : 1.0;
            gid[u] = gids[ii];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (i + u >= n) break;
            const double dx = __dsub_rn(X, g[u].x);
            const double dxx = __dmul_rn(dx, dx);
#pragma unroll
            for (int j = 0; j < R; ++j) {
                const double dy = __dsub_rn(Y[j], g[u].y);
                double d = __dadd_rn(dxx, __dmul_rn(dy, dy));
                if (HAS_SCALE) d = __dmul_rn(d, iv[u]);
                if ((d < bd[j]) || ((d == bd[j]) && (gid[u] < bi[j]))) { bd[j] = d; bi[j] = gid[u]; }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < R; ++j) {
        const long r = r0 + j;
        if (c >= a.W || r >= a.H) continue;
        const long idx = r * a.W + c;
        labels[idx] = bi[j];
        if (MODE != 0) {
            double val[NV];
            bool ok;
            if (MODE == 1) {
                const double w = a.dens[idx];
                ok = isfinite(w);
                val[0] = w; val[1] = __dmul_rn(w, X); val[2] = __dmul_rn(w, Y[j]);
            } else {
                const double S = a.sig[idx], N = a.noise[idx];
                double w = 1.0;
                if (!a.uniform_weight) w = tess_sn_weight(S, N);
                ok = isfinite(w) && isfinite(S) && isfinite(N);
                val[0] = w; val[1] = __dmul_rn(w, X); val[2] = __dmul_rn(w, Y[j]);
                val[NV - 2] = S; val[NV - 1] = __dmul_rn(N, N);
            }
            if (ok) {
                double* m = a.mom + (long)bi[j] * (NV + 1);
                tess_gmem_add(m, 1.0);
#pragma unroll
                for (int i = 0; i < NV; ++i) tess_gmem_add(m + 1 + i, val[i]);
            }
        }
    }
    __syncwarp();
}

// Fallback for the quadrants off the fast path -- the tile overflowed the list builder (bins of a few
// pixels, q_cnt = -1: ring search) or the quadrant's list is longer than the fast path stages
// (q_cnt > QCAP: tess_rows_scan over that list).  Ring search: every
// pixel runs the exact ring search of the point-set path over the generator grid (tess_points.cuh),
// so the cost follows the LOCAL generator density, not k.  The builder collects those quadrants in
// L.ring_q; after its own quadrants every warp of the grid takes strips of them (TESS_RING_ROWS rows of
// 32 pixels per warp and turn, one column per lane: walking down a column keeps the cells in L1), so
// that a few slow quadrants do not become the kernel's tail.
// Moments go straight to the global vector.
template <int MODE, bool HAS_SCALE>
TESS_NOINLINE void tess_rows_ring(const TessImgArgs& a, int32_t* labels, long qx0, long r0) {
    constexpr int NV = (MODE == 2) ? 5 : 3;
    const int lane = threadIdx.x & 31;
    const long c = qx0 + lane;
    const double min_inv = HAS_SCALE ? a.st->min_inv_s2 : 1.0;
    for (long r = r0; r < r0 + TESS_RING_ROWS; ++r) {
    if (c < a.W && r < a.H) {
        const double X = a.x0 + (double)c, Y = a.y0 + (double)r;
        const int bi = tess_nearest_ring<HAS_SCALE>(a.g, a.cell_start, a.cell_items, a.cell_xy, a.cell_inv, min_inv, X, Y);
        const long idx = r * a.W + c;
        labels[idx] = bi;
        if (MODE != 0) {
            double val[NV];
            bool ok;
            if (MODE == 1) {
                const double w = a.dens[idx];
                ok = isfinite(w);
                val[0] = w; val[1] = __dmul_rn(w, X); val[2] = __dmul_rn(w, Y);
            } else {
                const double S = a.sig[idx], N = a.noise[idx];
                double w = 1.0;
                if (!a.uniform_weight) w = tess_sn_weight(S, N);
                ok = isfinite(w) && isfinite(S) && isfinite(N);
                val[0] = w; val[1] = __dmul_rn(w, X); val[2] = __dmul_rn(w, Y);
                val[NV - 2] = S; val[NV - 1] = __dmul_rn(N, N);
            }
            if (ok) {
                double* m = a.mom + (long)bi * (NV + 1);
                tess_gmem_add(m, 1.0);
#pragma unroll
                for (int i = 0; i < NV; ++i) tess_gmem_add(m + 1 + i, val[i]);
            }
        }
    }
    }
    __syncwarp();
}

// ---- 16x16 block lists (quadrant lists longer than TESS_BLOCK_MIN only).  Centre + radius filter of
// the quadrant list for two 16x16 blocks side by side, each done by 16 lanes (the lanes of eight
// adjacent pixel-column pairs, both row halves):
//   keep g  iff  u(c,g) <= min_h u(c,h) + 2 rho / s_min      (u = distance / scale)
// Order preserving; the survivors' quadrant-list positions go to `out` (at most `cap` of them).
// Returns the number of survivors of the lane's block (which may exceed `cap`).
template <bool HAS_SCALE>
TESS_DEVFN int tess_block_filter(const double2* qxy, const double* qinv, int plen, double ccx, double ccy, double rho,
                                 bool active, int lane, unsigned short* out, int cap) {
    const int rank = (lane & 7) | ((lane >> 4) << 3);
    const int half = (lane >> 3) & 1;
    double umin = INFINITY, imax = 1.0;
#pragma unroll 2
    for (int i = rank; i < plen; i += 16) {
        const double2 g = qxy[i];
        const double dx = ccx - g.x, dy = ccy - g.y;
        double d = dx * dx + dy * dy;
        if (HAS_SCALE) { const double iv = qinv[i]; d *= iv; imax = fmax(imax, iv); }
        umin = fmin(umin, d);
    }
    umin = fmin(umin, __shfl_xor_sync(TESS_FULL, umin, 1));
    umin = fmin(umin, __shfl_xor_sync(TESS_FULL, umin, 2));
    umin = fmin(umin, __shfl_xor_sync(TESS_FULL, umin, 4));
    umin = fmin(umin, __shfl_xor_sync(TESS_FULL, umin, 16));
    double c = 2.0 * rho;
    if (HAS_SCALE) {
        imax = fmax(imax, __shfl_xor_sync(TESS_FULL, imax, 1));
        imax = fmax(imax, __shfl_xor_sync(TESS_FULL, imax, 2));
        imax = fmax(imax, __shfl_xor_sync(TESS_FULL, imax, 4));
        imax = fmax(imax, __shfl_xor_sync(TESS_FULL, imax, 16));
        c *= tess_sqrt_up(imax);
    }
    const double thr = tess_center_thr(umin, c);
    int n = 0;
#pragma unroll 2
    for (int b0 = 0; b0 < plen; b0 += 16) {            // plen is warp-uniform (ballots inside)
        const int i = b0 + rank;
        bool keep = false;
        if (i < plen) {
            const double2 g = qxy[i];
            const double dx = ccx - g.x, dy = ccy - g.y;
            double d = dx * dx + dy * dy;
            if (HAS_SCALE) d *= qinv[i];
            keep = active && (d <= thr);
        }
        const unsigned bb = __ballot_sync(TESS_FULL, keep);
        const unsigned gb = ((bb >> (8 * half)) & 0xFFu) | (((bb >> (16 + 8 * half)) & 0xFFu) << 8);
        const int pos = n + __popc(gb & ((1u << rank) - 1u));
        if (keep && pos < cap) out[pos] = (unsigned short)i;
        n += __popc(gb);
    }
    __syncwarp();
    return n;
}

// ---- the candidate sets of all sixteen 8x8 leaves of a quadrant at once, as bit masks over a parent
// list of at most 64 entries (the quadrant list itself, or the 16x16 block list of the leaf's block:
// `pl`, positions in the quadrant list).  Leaf L = lane >> 1 is done by the lane pair (2L, 2L+1): each
// lane measures every other entry of the parent list against the leaf's centre,
//   keep g  iff  u(c,g) <= min_h u(c,h) + 2 rho / s_min      (u = distance / scale)
// which can never drop the true nearest generator of any pixel of the leaf.  Bit i of (lo, hi) =
// parent entry i survives; both lanes of the pair return the leaf's whole mask.  One shuffle for the
// minimum and one for the mask per QUADRANT (the per-warp-tile filter this replaces cost four
// fp64 shuffle rounds, a compaction with ballots and a rewrite of the survivors' records per tile).
// The test is a bound, not a result, so it is done on fp32 copies of the fp64 centre distances that
// err on the safe side only (distances rounded down, the minimum and the threshold rounded up): a
// borderline candidate too many is evaluated exactly like every other survivor.
template <bool HAS_SCALE>
TESS_DEVFN float tess_center_d2_lo(const double2* qxy, const double* qinv, int e, double ccx, double ccy, double& imax) {
    const double2 g = qxy[e];
    const double dx = ccx - g.x, dy = ccy - g.y;
    double d = dx * dx + dy * dy;
    if (HAS_SCALE) { const double iv = qinv[e]; d *= iv; imax = fmax(imax, iv); }
#ifndef TESS_EMU
    return __double2float_rd(d);
#else
    float f = (float)d;
    if ((double)f > d) f = nextafterf(f, -INFINITY);
    return f;
#endif
}
template <bool HAS_SCALE>
TESS_DEVFN void tess_leaf_masks(const double2* qxy, const double* qinv, const unsigned short* pl, bool ident, int plen,
                                double ccx, double ccy, double rho, int lane, unsigned& lo, unsigned& hi) {
    const int h = lane & 1;
    const int last = max(plen - 1, 0);
    double imax = 1.0;
    // the first sixteen parent entries (eight per lane): loads issued together (entries beyond the
    // list re-read its last one and are discarded), distances kept in registers for the second pass
    float dd[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        const int i = 2 * t + h;
        int e = min(i, last);
        if (!ident) e = (plen > 0) ? (int)pl[e] : 0;
        const float d = tess_center_d2_lo<HAS_SCALE>(qxy, qinv, e, ccx, ccy, imax);
        dd[t] = (i < plen) ? d : INFINITY;
    }
    float fmn = fminf(fminf(fminf(dd[0], dd[1]), fminf(dd[2], dd[3])), fminf(fminf(dd[4], dd[5]), fminf(dd[6], dd[7])));
    // longer parent lists: the rest is measured twice
#pragma unroll 2
    for (int i = 16 + h; i < plen; i += 2) {
        const int e = ident ? i : (int)pl[i];
        fmn = fminf(fmn, tess_center_d2_lo<HAS_SCALE>(qxy, qinv, e, ccx, ccy, imax));
    }
    fmn = fminf(fmn, __shfl_xor_sync(TESS_FULL, fmn, 1));
    // an upper bound of the true minimum: the next float up (fmn is a rounded-down, non-negative value)
    const double umin = (double)__int_as_float(__float_as_int(fmn) + 1);
    double c = 2.0 * rho;
    if (HAS_SCALE) {
        imax = fmax(imax, __shfl_xor_sync(TESS_FULL, imax, 1));
        c *= tess_sqrt_up(imax);
    }
#ifndef TESS_EMU
    const float thr = __double2float_ru(tess_center_thr(umin, c));
#else
    const float thr = nextafterf((float)tess_center_thr(umin, c), INFINITY);
#endif
    unsigned m16 = 0u;
#pragma unroll
    for (int t = 0; t < 8; ++t) m16 |= (dd[t] <= thr) ? (1u << (2 * t)) : 0u;     // (inf beyond the list)
    unsigned long long mm = (unsigned long long)(m16 << h);
#pragma unroll 2
    for (int i = 16 + h; i < plen; i += 2) {
        const int e = ident ? i : (int)pl[i];
        double dummy = 1.0;
        const float d = tess_center_d2_lo<HAS_SCALE>(qxy, qinv, e, ccx, ccy, dummy);
        mm |= (d <= thr) ? (1ull << i) : 0ull;
    }
    const unsigned mlo = (unsigned)mm, mhi = (unsigned)(mm >> 32);
    lo = mlo | __shfl_xor_sync(TESS_FULL, mlo, 1);
    hi = mhi | __shfl_xor_sync(TESS_FULL, mhi, 1);
}

// Takes the lowest set bit out of the 64-bit mask (lo, hi): its parent-list entry is translated to
// the quadrant-list position `e` (0 when the mask was empty: a harmless read).  Returns whether
// there was a bit.
TESS_DEVFN bool tess_mask_pop(unsigned& lo, unsigned& hi, const unsigned short* pl, bool ident, int& e) {
    const bool have = (lo | hi) != 0u;
    const bool in_lo = lo != 0u;
    const unsigned w = in_lo ? lo : hi;
    int idx = (__ffs((int)w) - 1) + (in_lo ? 0 : 32);
    const unsigned w2 = w & (w - 1u);
    lo = in_lo ? w2 : 0u;
    hi = in_lo ? hi : w2;
    idx = have ? idx : 0;
    int ee = idx;
    if (!ident) ee = (int)pl[idx];
    e = have ? ee : 0;
    return have;
}

// Loads the lane's 2x4 pixel patch whose first pixel is (px, py): weight (MODE 1) or signal and
// noise (MODE 2).  Interior quadrants: four 128-bit loads; else guarded scalar loads, NaN outside.
template <int MODE>
TESS_DEVFN void tess_load_patch(const TessImgArgs& a, const double* src, bool interior, long px, long py,
                                double (&w)[4][2], double (&nz)[4][2]) {
    const long base0 = py * a.W + px;
    if (interior) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const double2 v = *reinterpret_cast<const double2*>(src + base0 + r * a.W);
            w[r][0] = v.x; w[r][1] = v.y;
            if (MODE == 2) {
                const double2 u = *reinterpret_cast<const double2*>(a.noise + base0 + r * a.W);
                nz[r][0] = u.x; nz[r][1] = u.y;
            }
        }
    } else {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const bool in = (py + r < a.H) && (px + c < a.W);
                w[r][c] = in ? src[base0 + r * a.W + c] : TESS_NAN;
                if (MODE == 2) nz[r][c] = in ? a.noise[base0 + r * a.W + c] : TESS_NAN;
            }
    }
}

// per-warp shared memory of the main kernel
template <int MODE, bool HAS_SCALE>
struct TessWarpSmem {
    static constexpr int NV = (MODE == 2) ? 5 : 3;
    static constexpr int QC = TESS_QCAP(MODE);
    double2 qxy[2][QC];                                      // quadrant list (double buffered): coordinates,
    double qinv[HAS_SCALE ? 2 : 1][HAS_SCALE ? QC : 1];     //   1/scale^2,
    int qgid[2][QC];                                         //   generator index
    unsigned short bl[4][TESS_BCAP];                         // 16x16 block lists (quadrant-list positions)
};

// A quadrant whose pixels all belong to generator `gid` (interior, MODE 1): constant labels and
// closed-form patch sums that simply extend the lane's run.  Returns false (and leaves the run alone)
// if some pixel is masked: the caller then takes the general path, which rewrites the labels.
template <int NV>
TESS_DEVFN bool tess_one_bin_quadrant(const TessImgArgs& a, int32_t* labels, const double* src, int gid, int px, int qy0,
                                      int ry, double X0, double X1, int lane, const double (&w0)[4][2], TessRun<NV>& cur) {
    // the patches of warp tiles 1..3 are requested together (tile 0 was loaded by the caller)
    double2 v[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r) v[0][r] = make_double2(w0[r][0], w0[r][1]);
#pragma unroll
    for (int step = 1; step < 4; ++step) {
        const long base0 = (long)(qy0 + 8 * step + 4 * ry) * a.W + px;
#pragma unroll
        for (int r = 0; r < 4; ++r) v[step][r] = *reinterpret_cast<const double2*>(src + base0 + r * a.W);
    }
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    bool clean = true;
#pragma unroll
    for (int step = 0; step < 4; ++step) {
        const int py = qy0 + 8 * step + 4 * ry;
        const long base0 = (long)py * a.W + px;
#pragma unroll
        for (int r = 0; r < 4; ++r) *reinterpret_cast<int2*>(labels + base0 + r * a.W) = make_int2(gid, gid);
        const double c0 = __dadd_rn(__dadd_rn(v[step][0].x, v[step][1].x), __dadd_rn(v[step][2].x, v[step][3].x));
        const double c1 = __dadd_rn(__dadd_rn(v[step][0].y, v[step][1].y), __dadd_rn(v[step][2].y, v[step][3].y));
        double sy = 0.0;
#pragma unroll
        for (int r = 0; r < 4; ++r)
            sy = __dadd_rn(sy, __dmul_rn(__dadd_rn(v[step][r].x, v[step][r].y), a.y0 + (double)(py + r)));
        const double t = __dadd_rn(c0, c1);
        clean = clean && (fabs(t) <= 1.7976931348623157e308);
        s0 = __dadd_rn(s0, t);
        s1 = __dadd_rn(s1, __dadd_rn(__dmul_rn(c0, X0), __dmul_rn(c1, X1)));
        s2 = __dadd_rn(s2, sy);
    }
    if (!__all_sync(TESS_FULL, clean)) return false;
    // the whole quadrant joins the lane's run (no shuffles, no reductions while the bin lasts)
    const bool sw = (cur.key != gid);
    tess_run_flush<NV>(sw && cur.n > 0, cur, a.mom, lane);
    tess_run_switch<NV>(sw, gid, cur);
    cur.n += 32;
    cur.v[0] = __dadd_rn(cur.v[0], s0);
    cur.v[1] = __dadd_rn(cur.v[1], s1);
    cur.v[2] = __dadd_rn(cur.v[2], s2);
    return true;
}

// One quadrant on the fast path (list of at most QCAP candidates staged in sm.q*[buf]).  INTERIOR:
// every pixel access is in bounds and 16-byte aligned (128-bit loads, 64-bit label stores, constant
// rectangle radii); the general variant guards every access and is kept out of line.
template <int MODE, bool HAS_SCALE, bool INTERIOR>
TESS_DEVFN void tess_quadrant_body(const TessImgArgs& a, TessWarpSmem<MODE, HAS_SCALE>& sm, int buf, int32_t* labels,
                                   const double* src, int c_q, int qx0, int qy0, int lane,
                                   TessRun<(MODE == 2) ? 5 : 3>& cur) {
    constexpr int NV = (MODE == 2) ? 5 : 3;
    const int cx = lane & 15, ry = lane >> 4;          // patch: cols 2cx,2cx+1; rows 4ry..4ry+3
    const int leaf = cx >> 2;                          // 8-px-wide leaf of the warp tile
    const int blk = cx >> 3;                           // 16-px-wide block
    const int Wi = (int)a.W, Hi = (int)a.H;
    const double2* const qxy = sm.qxy[buf];
    const double* const qinv = HAS_SCALE ? sm.qinv[buf] : nullptr;
    const int* const gidv = sm.qgid[buf];
    constexpr bool interior = INTERIOR;   // every pixel access in bounds and 16-byte aligned
    const int px = qx0 + 2 * cx;                     // first column of this lane
    const double X0 = a.x0 + (double)px, X1 = a.x0 + (double)(px + 1);

    int one_gid = (c_q == 1) ? gidv[0] : -1;      // the quadrant's only bin, if it has one (warp-uniform)

    // ---- pixel data of the first warp tile: requested before the candidate masks are computed, so that
    //      its latency hides behind them.  Pixels outside the image read as NaN and drop out with the
    //      masked ones.  The following tiles' data is requested at the end of the tile before.
    double w_[4][2], n_[4][2];                     // MODE 1: weight; MODE 2: signal, noise
    if (MODE != 0) tess_load_patch<MODE>(a, src, interior, px, qy0 + 4 * ry, w_, n_);

    // ---- candidate sets of the quadrant's sixteen 8x8 leaves, once per quadrant: lane pair
    //      (2L, 2L+1) holds the mask of leaf L = 4 (leaf row) + (leaf column) over its parent list --
    //      the quadrant list itself, or (long lists) the list of the leaf's 16x16 block
    unsigned mlo = 0u, mhi = 0u;          // mask of leaf (lane >> 1)
    bool ident = true;                    // parent lists are the quadrant list (warp-uniform)
    bool slowq = false;                   // a block list overflowed: every pixel scans the quadrant list (rare)
    if (c_q > 1) {
        const int L = lane >> 1, lxi = L & 3, lyi = L >> 2;
        int plen = c_q;
        const unsigned short* plm = nullptr;
        if (c_q > TESS_BLOCK_MIN) {
            int nb[2];
#pragma unroll
            for (int hb = 0; hb < 2; ++hb) {              // top / bottom pair of blocks
                const int bx0 = qx0 + 16 * blk, by0 = qy0 + 16 * hb;
                const int bx1 = interior ? bx0 + 15 : min(bx0 + 15, Wi - 1), by1 = interior ? by0 + 15 : min(by0 + 15, Hi - 1);
                const double hx = 0.5 * (double)(bx1 - bx0), hy = 0.5 * (double)(by1 - by0);
                const double rho = interior ? 10.606602778 : sqrt(fmax(hx * hx + hy * hy, 0.0)) * 1.0000001;   // 7.5 sqrt 2
                nb[hb] = tess_block_filter<HAS_SCALE>(qxy, qinv, c_q, a.x0 + 0.5 * (double)(bx0 + bx1),
                                                      a.y0 + 0.5 * (double)(by0 + by1), rho,
                                                      interior || (bx0 < Wi && by0 < Hi), lane, sm.bl[2 * hb + blk], TESS_BCAP);
            }
            // lengths of the four block lists (lane 0 sits in block column 0, lane 8 in column 1)
            const int n00 = __shfl_sync(TESS_FULL, nb[0], 0), n01 = __shfl_sync(TESS_FULL, nb[0], 8);
            const int n10 = __shfl_sync(TESS_FULL, nb[1], 0), n11 = __shfl_sync(TESS_FULL, nb[1], 8);
            slowq = max(max(n00, n01), max(n10, n11)) > TESS_BCAP;
            ident = false;
            const int b = 2 * (lyi >> 1) + (lxi >> 1);
            plen = (b == 0) ? n00 : (b == 1) ? n01 : (b == 2) ? n10 : n11;
            plm = sm.bl[b];
        }
        if (!slowq) {
            const int lx0 = qx0 + 8 * lxi, ly0 = qy0 + 8 * lyi;
            const int lx1 = interior ? lx0 + 7 : min(lx0 + 7, Wi - 1), ly1 = interior ? ly0 + 7 : min(ly0 + 7, Hi - 1);
            const double hx = 0.5 * (double)(lx1 - lx0), hy = 0.5 * (double)(ly1 - ly0);
            const double rho = interior ? 4.9497479632 : sqrt(fmax(hx * hx + hy * hy, 0.0)) * 1.0000001;   // 3.5 sqrt 2
            const bool active = interior || (lx0 < Wi && ly0 < Hi);
            tess_leaf_masks<HAS_SCALE>(qxy, qinv, plm, ident, active ? plen : 0, a.x0 + 0.5 * (double)(lx0 + lx1),
                                       a.y0 + 0.5 * (double)(ly0 + ly1), rho, lane, mlo, mhi);
            // ---- all sixteen leaves ended with the same single candidate: the quadrant is one bin
            if (MODE == 1 && interior && ident) {
                const unsigned m0 = __shfl_sync(TESS_FULL, mlo, 0);
                if (__all_sync(TESS_FULL, (mlo == m0) && (mhi == 0u)) && __popc(m0) == 1) one_gid = gidv[__ffs((int)m0) - 1];
            }
        }
    }
    // ---- one-bin quadrant (interior, MODE 1): constant labels, closed-form patch sums
    if (MODE == 1 && interior && one_gid >= 0) {
        if (tess_one_bin_quadrant<NV>(a, labels, src, one_gid, px, qy0, ry, X0, X1, lane, w_, cur)) return;
        // a masked pixel somewhere: the general path below rewrites the labels
    }

    for (int step = 0; step < 4; ++step) {
        const int wy0 = qy0 + 8 * step;
        if (!interior && wy0 >= a.H) break;            // warp-uniform
        const int py = wy0 + 4 * ry;                   // first row of this lane
        const long base0 = (long)py * a.W + px;              // flat index of the patch's first pixel
        double Y[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) Y[r] = a.y0 + (double)(py + r);

        // ---- candidates of this lane's 8x8 leaf: the mask computed above, from the lane pair that
        //      holds leaf 4 step + leaf
        const unsigned short* const pl = ident ? nullptr : sm.bl[2 * (step >> 1) + blk];
        const unsigned m_lo = __shfl_sync(TESS_FULL, mlo, 2 * (4 * step + leaf));
        const unsigned m_hi = __shfl_sync(TESS_FULL, mhi, 2 * (4 * step + leaf));
        bool uniform = (c_q == 1);                     // the whole warp tile belongs to one bin
        int ukey = 0;
        if (c_q > 1 && !slowq) {
            // all (in-image) leaves ended with the same single candidate?
            const int len = __popc(m_lo) + __popc(m_hi);
            int e0 = (m_lo != 0u) ? (__ffs((int)m_lo) - 1) : ((m_hi != 0u) ? 31 + __ffs((int)m_hi) : 0);
            if (!ident) e0 = (int)pl[e0];
            const int one = (len == 1) ? e0 : ((interior || qx0 + 8 * leaf < Wi) ? -1 : -3);
            const int ref = __reduce_max_sync(TESS_FULL, one);
            uniform = (ref >= 0) && __all_sync(TESS_FULL, one == ref || one == -3);
            ukey = ref;
        }

        int bs[4][2];                                  // winner per pixel: quadrant-list position, then generator id
        if (uniform) {
            // ---- one bin for all 256 pixels: constant labels
            const int gid = gidv[ukey];
            if (interior) {
#pragma unroll
                for (int r = 0; r < 4; ++r) *reinterpret_cast<int2*>(labels + base0 + r * a.W) = make_int2(gid, gid);
            } else {
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int c = 0; c < 2; ++c)
                        if ((py + r < a.H) && (px + c < a.W)) labels[base0 + r * a.W + c] = gid;
            }
#pragma unroll
            for (int r = 0; r < 4; ++r) { bs[r][0] = gid; bs[r][1] = gid; }
            ukey = gid;                                // from here on keys are generator ids
        } else {
            // ---- evaluate the surviving candidates with the pinned formula.  Lists are in no
            //      particular order, so an exact tie (d == best so far) is only RECORDED here; the
            //      rare warp that saw one re-evaluates with the full (distance, index) order below.
            double bd[4][2];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 2; ++c) { bd[r][c] = INFINITY; bs[r][c] = 0; }
            bool tied = slowq;         // (a quadrant with an overflowed block list is done by the exact loop below)
            // (the record of the NEXT candidate is read while this one is evaluated)
            unsigned lo = slowq ? 0u : m_lo, hi = slowq ? 0u : m_hi;
            int en;
            bool more = tess_mask_pop(lo, hi, pl, ident, en);
            double2 gn = qxy[en];
            double invn = HAS_SCALE ? qinv[en] : 1.0;
#pragma unroll 1     // the compiler's own 2x unrolling costs 1.3 % (instruction cache)
            while (more) {
                const double2 g = gn;
                const int e = en;
                const double inv = invn;
                more = tess_mask_pop(lo, hi, pl, ident, en);
                gn = qxy[en];
                if (HAS_SCALE) invn = qinv[en];
                const double dx0 = __dsub_rn(X0, g.x), dx1 = __dsub_rn(X1, g.x);
                const double dxx0 = __dmul_rn(dx0, dx0), dxx1 = __dmul_rn(dx1, dx1);
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const double dy = __dsub_rn(Y[r], g.y);
                    const double dyy = __dmul_rn(dy, dy);
                    double d0 = __dadd_rn(dxx0, dyy), d1 = __dadd_rn(dxx1, dyy);
                    if (HAS_SCALE) { d0 = __dmul_rn(d0, inv); d1 = __dmul_rn(d1, inv); }
                    tied = tied || (d0 == bd[r][0]) || (d1 == bd[r][1]);
                    // (select pairs, not fmin: DMNMX measured 9 % slower on the whole kernel)
                    if (d0 < bd[r][0]) { bd[r][0] = d0; bs[r][0] = e; }
                    if (d1 < bd[r][1]) { bd[r][1] = d1; bs[r][1] = e; }
                }
            }
            if (__any_sync(TESS_FULL, tied)) {
                // exact ties: lowest generator index wins
                int bg[4][2];
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int c = 0; c < 2; ++c) { bd[r][c] = INFINITY; bs[r][c] = 0; bg[r][c] = 0x7fffffff; }
                lo = m_lo; hi = m_hi;
                int j = 0;
#pragma unroll 1
                for (;;) {
                    int e;
                    if (slowq) { if (j >= c_q) break; e = j++; }
                    else if (!tess_mask_pop(lo, hi, pl, ident, e)) break;
                    const double2 g = qxy[e];
                    const double inv = HAS_SCALE ? qinv[e] : 1.0;
                    const int gid = gidv[e];
#pragma unroll
                    for (int r = 0; r < 4; ++r)
#pragma unroll
                        for (int c = 0; c < 2; ++c) {
                            double d = tess_d2_pinned(c ? X1 : X0, Y[r], g.x, g.y);
                            if (HAS_SCALE) d = __dmul_rn(d, inv);
                            if ((d < bd[r][c]) || ((d == bd[r][c]) && (gid < bg[r][c]))) {
                                bd[r][c] = d; bs[r][c] = e; bg[r][c] = gid;
                            }
                        }
                }
            }
            __syncwarp();                              // the leaves' lists differ in length
            // ---- winners as generator ids (they key the runs and are the labels)
#pragma unroll
            for (int r = 0; r < 4; ++r) { bs[r][0] = gidv[bs[r][0]]; bs[r][1] = gidv[bs[r][1]]; }
            // ---- labels out: 16 lanes x 8 B = one 128-byte line per row
            if (interior) {
#pragma unroll
                for (int r = 0; r < 4; ++r)
                    *reinterpret_cast<int2*>(labels + base0 + r * a.W) = make_int2(bs[r][0], bs[r][1]);
            } else {
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int c = 0; c < 2; ++c)
                        if ((py + r < a.H) && (px + c < a.W)) labels[base0 + r * a.W + c] = bs[r][c];
            }
        }
        if (MODE == 0) continue;
#if TESS_EARLY_LOAD
        // the next warp tile's pixel data is requested before this tile's fold
        double wn_[4][2], nn_[4][2];
        const bool more_tiles = step < 3 && (interior || wy0 + 8 < a.H);
        if (more_tiles) tess_load_patch<MODE>(a, src, interior, px, py + 8, wn_, nn_);
#endif

        // ---- moments.  Weights first.  The sum of the patch's weights is finite iff all of them
        //      are (up to a harmless false alarm on overflow), so one test per lane decides whether
        //      any pixel of the warp tile is masked (non-finite weight, signal or noise; outside
        //      the image); only then are the masked pixels marked consumed (key -2) one by one.
        double wv[4][2];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                double w = w_[r][c];
                if (MODE == 2) {
                    const double S = w_[r][c], N = n_[r][c];
                    w = 1.0;
                    if (!a.uniform_weight) w = tess_sn_weight(S, N);
                    // a non-finite signal or noise poisons the weight
                    if (!((fabs(S) <= 1.7976931348623157e308) && (fabs(N) <= 1.7976931348623157e308))) w = TESS_NAN;
                }
                wv[r][c] = w;
            }
        const double c0 = __dadd_rn(__dadd_rn(wv[0][0], wv[1][0]), __dadd_rn(wv[2][0], wv[3][0]));
        const double c1 = __dadd_rn(__dadd_rn(wv[0][1], wv[1][1]), __dadd_rn(wv[2][1], wv[3][1]));
        const double wtot = __dadd_rn(c0, c1);
        const bool all_clean = __all_sync(TESS_FULL, fabs(wtot) <= 1.7976931348623157e308);
        if (!all_clean) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 2; ++c)
                    bs[r][c] = (fabs(wv[r][c]) <= 1.7976931348623157e308) ? bs[r][c] : -2;
        }
        if (uniform && all_clean) {
            // closed form for a one-bin patch: sum w x = X0 col0 + X1 col1, sum w y = sum_r Y_r row_r
            const bool sw = (cur.key != ukey);
            tess_run_flush<NV>(sw && cur.n > 0, cur, a.mom, lane);
            tess_run_switch<NV>(sw, ukey, cur);
            double sy = 0.0;
#pragma unroll
            for (int r = 0; r < 4; ++r) sy = __dadd_rn(sy, __dmul_rn(__dadd_rn(wv[r][0], wv[r][1]), Y[r]));
            cur.n += 8;
            cur.v[0] = __dadd_rn(cur.v[0], wtot);
            cur.v[1] = __dadd_rn(cur.v[1], __dadd_rn(__dmul_rn(c0, X0), __dmul_rn(c1, X1)));
            cur.v[2] = __dadd_rn(cur.v[2], sy);
            if (MODE == 2) {
                double ss = 0.0, nn = 0.0;
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        ss = __dadd_rn(ss, w_[r][c]);
                        nn = __dadd_rn(nn, __dmul_rn(n_[r][c], n_[r][c]));
                    }
                cur.v[NV - 2] = __dadd_rn(cur.v[NV - 2], ss);
                cur.v[NV - 1] = __dadd_rn(cur.v[NV - 1], nn);
            }
        } else {
            // Rounds (no lane-divergent branch): a lane whose run's bin is no longer among its
            // unconsumed pixels ends the run and starts the run of one of the remaining bins;
            // then every lane folds all its pixels of its run's bin.
            for (;;) {
                const int rest = max(max(max(bs[0][0], bs[0][1]), max(bs[1][0], bs[1][1])),
                                     max(max(bs[2][0], bs[2][1]), max(bs[3][0], bs[3][1])));
                if (!__any_sync(TESS_FULL, rest >= 0)) break;
                bool present = false;
#pragma unroll
                for (int r = 0; r < 4; ++r) present = present || (bs[r][0] == cur.key) || (bs[r][1] == cur.key);
                const bool sw = (rest >= 0) && !present;
                tess_run_flush<NV>(sw && cur.n > 0, cur, a.mom, lane);
                tess_run_switch<NV>(sw, rest, cur);
                // the lane's pixels of this bin, summed in patch-local coordinates (column offset
                // 0/1, row offset 0..3) and shifted once: sum w x = X0 sum w + sum_{col 1} w,
                // sum w y = Y0 sum w + sum_r r (w_r0 + w_r1).  Predicated adds: a pixel of another
                // bin costs its compare only.
                const int key = cur.key;
                double sA = 0.0, sB = 0.0, c1 = 0.0, sy = 0.0, ss = 0.0, nn = 0.0;
                int nh = 0;
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    if (bs[r][0] == key) {
                        sA = __dadd_rn(sA, wv[r][0]);
                        if (r == 1) sy = __dadd_rn(sy, wv[r][0]);
                        if (r > 1) sy = __fma_rn((double)r, wv[r][0], sy);
                        if (MODE == 2) { ss = __dadd_rn(ss, w_[r][0]); nn = __fma_rn(n_[r][0], n_[r][0], nn); }
                        ++nh;
                        bs[r][0] = -2;
                    }
                    if (bs[r][1] == key) {
                        sB = __dadd_rn(sB, wv[r][1]);
                        if (r == 1) sy = __dadd_rn(sy, wv[r][1]);
                        if (r > 1) sy = __fma_rn((double)r, wv[r][1], sy);
                        if (MODE == 2) { ss = __dadd_rn(ss, w_[r][1]); nn = __fma_rn(n_[r][1], n_[r][1], nn); }
                        ++nh;
                        bs[r][1] = -2;
                    }
                }
                c1 = sB;
                const double wsum = __dadd_rn(sA, sB);
                cur.n += nh;
                cur.v[0] = __dadd_rn(cur.v[0], wsum);
                cur.v[1] = __dadd_rn(cur.v[1], __fma_rn(X0, wsum, c1));
                cur.v[2] = __dadd_rn(cur.v[2], __fma_rn(Y[0], wsum, sy));
                if (MODE == 2) {
                    cur.v[NV - 2] = __dadd_rn(cur.v[NV - 2], ss);
                    cur.v[NV - 1] = __dadd_rn(cur.v[NV - 1], nn);
                }
            }
        }
        // ---- the next warp tile's pixel data (into the registers this tile's fold has just released)
#if TESS_EARLY_LOAD
        if (more_tiles) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 2; ++c) { w_[r][c] = wn_[r][c]; if (MODE == 2) n_[r][c] = nn_[r][c]; }
        }
#else
        if (step < 3 && (interior || wy0 + 8 < a.H)) tess_load_patch<MODE>(a, src, interior, px, py + 8, w_, n_);
#endif
    }   // steps

    // (the lanes' runs stay live: the next quadrant may continue them)
}

// One ticket from the quadrant queue.  The ticket is not needed before the next chunk boundary, but
// both the compiler and ptxas "warp-aggregate" an atomic add of a provably uniform value under a
// divergent predicate (vote, leader election, ONE atomic, shuffle of its result) -- and the shuffle
// right behind the atomic makes the warp wait a whole round trip (~900 cycles per chunk, 2.5-3.8 % of
// the kernel's stall samples).  The addend therefore comes from a per-lane slot of a constant-memory
// table of ones: nothing to prove, nothing aggregated, the result register is simply scoreboarded.
#ifndef TESS_EMU
__constant__ int c_tess_ones[32] = {1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1,
                                    1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1};
#endif
TESS_DEVFN int tess_queue_take(int* counter, int one) {
#ifndef TESS_EMU
    int t;
    asm volatile("atom.global.add.u32 %0, [%1], %2;" : "=r"(t) : "l"(counter), "r"(one) : "memory");
    return t;
#else
    return atomicAdd(counter, one);
#endif
}

// MODE 0: labels only; 1: density map (4 moments); 2: signal + noise maps (6 moments).
// EDGE = false: the interior quadrants (every access in bounds and aligned; `vec` must hold);
// EDGE = true: the remaining ones (image edges; all of them when vec == 0), with guarded accesses.
template <int MODE, bool HAS_SCALE, bool EDGE>
__global__ void __launch_bounds__(TESS_MAIN_THREADS, (MODE == 2) ? TESS_MAIN_MINBLOCKS_SN : TESS_MAIN_MINBLOCKS)
k_image_main(const TESS_GRID_CONSTANT TessImgArgs a, int vec, int use_tma, const TESS_GRID_CONSTANT TessTileMap tm0,
             const TESS_GRID_CONSTANT TessTileMap tm1) {
    constexpr int NV = (MODE == 2) ? 5 : 3;
    constexpr int NW = TESS_MAIN_THREADS / 32;
    constexpr int QC = TESS_QCAP(MODE);
    if (TESS_STOPPED(a.st, a.honor)) return;
    int32_t* const labels = a.labels_ext ? a.labels_ext : ((a.st->iters_done & 1) ? a.lab1 : a.lab0);

    __shared__ __align__(16) TessWarpSmem<MODE, HAS_SCALE> s_w[NW];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    TessWarpSmem<MODE, HAS_SCALE>& sm = s_w[wib];
    const double* const src = (MODE == 2) ? a.sig : a.dens;   // MODE 1: weight map; MODE 2: signal map
    const int n_q = a.L.nqx * a.L.nqy;               // < 2^31 (checked by the host)
    const int stride = (int)gridDim.x * NW;

    TessRun<NV> cur;                 // the lane's running same-bin sum (registers), alive across quadrants
    cur.key = -1;
    cur.n = 0;
#pragma unroll
    for (int i = 0; i < NV; ++i) cur.v[i] = 0.0;
    // ---- software pipeline over this warp's quadrants: the list of quadrant i+1 is copied to shared
    //      memory (cp.async) while quadrant i is processed; (offset, count) of quadrant i+2 are loaded
    //      at the same time.  Quadrants are handed out in CHUNKS of TESS_CHUNK consecutive ones: the
    //      first chunk of a warp is fixed (its global warp index), the following ones are drawn from a
    //      device counter, so that no warp idles while others still have a queue (with fixed strides
    //      the per-warp sums of quadrant costs differed by ~9 %).  One atomic per chunk, requested a
    //      whole chunk before its value is needed: a counter asked once per quadrant by 2 400 warps
    //      was itself a hot spot (5 % of all warp stalls waited for its answer), and neighbouring
    //      quadrants let a lane's run live on.
    constexpr int G = TESS_CHUNK;
    int* const counter = EDGE ? reinterpret_cast<int*>(&a.st->spare_) : &a.st->tile_counter;
    int tok = 0;                                   // lane 0: the chunk after the current one
    #ifndef TESS_EMU
    const int one = c_tess_ones[lane];
#else
    const int one = 1;
#endif
    if (lane == 0) tok = tess_queue_take(counter, one);      // (raw ticket: nothing may depend on it yet)
    auto next_q = [&](int qq) {
        if (qq >= n_q) return n_q;
        int nq = qq + 1;
        if ((nq % G) == 0) {                       // chunk exhausted
            const int t = stride + __shfl_sync(TESS_FULL, tok, 0);
            const bool more = t < (n_q + G - 1) / G;
            if (lane == 0 && more) tok = tess_queue_take(counter, one);
            nq = more ? t * G : n_q;
        }
        return min(nq, n_q);
    };
    // (row, column) of quadrant qq without an integer division: estimate by a reciprocal multiply
    // (at most one too small for qq < 2^31), one correction
    const unsigned nqx_u = (unsigned)a.L.nqx, q_magic = 0xffffffffu / nqx_u;
    auto row_col = [&](int qq, int& col, int& row) {
        unsigned r = __umulhi((unsigned)qq, q_magic);
        unsigned c = (unsigned)qq - r * nqx_u;
        if (c >= nqx_u) { c -= nqx_u; ++r; }
        col = (int)c; row = (int)r;
    };
    int q = min(((int)blockIdx.x * NW + wib) * G, n_q);
    if ((long)((int)blockIdx.x * NW + wib) * G >= n_q) q = n_q;
    int cnt = (q < n_q) ? a.L.q_cnt[q] : 0, off = (q < n_q) ? a.L.q_off[q] : 0;
    int q1 = next_q(q);
    int cnt1 = (q1 < n_q) ? a.L.q_cnt[q1] : 0, off1 = (q1 < n_q) ? a.L.q_off[q1] : 0;
    int buf = 0;
    auto stage = [&](int b, int o, int n) {
        if (n > 0 && n <= QC) {
            for (int i = lane; i < n; i += 32) {
                tess_cp_async16(&sm.qxy[b][i], a.L.xy + o + i);
                tess_cp_async4(&sm.qgid[b][i], a.L.gid + o + i);
                if (HAS_SCALE) tess_cp_async8(&sm.qinv[b][i], a.L.inv + o + i);
            }
        }
    };
    stage(0, off, cnt);

    while (q < n_q) {
        tess_cp_async_wait();          // this quadrant's list has landed (waited for BEFORE any new load
        __syncwarp();                  //  is issued: the wait covers every load in flight)
        if (q1 < n_q) stage(buf ^ 1, off1, cnt1);
        // (offset, count) of the quadrant after the next: in flight during this quadrant
        const int q2 = next_q(q1);
        int cnt2 = 0, off2 = 0;
        if (q2 < n_q) { cnt2 = a.L.q_cnt[q2]; off2 = a.L.q_off[q2]; }
        int qx, qy;
        row_col(q, qx, qy);
        const int qx0 = 32 * qx, qy0 = 32 * qy;            // pixel coordinates fit 32 bits (host-checked)
        const bool interior = vec && (qx0 + 32 <= a.W) && (qy0 + 32 <= a.H);
#if TESS_PREFETCH
        // the NEXT quadrant's pixels start their trip from DRAM now: one 256-byte row per lane
        if (MODE != 0 && !EDGE && use_tma && q1 < n_q && cnt1 > 0 && lane == 0) {
            int qx1, qy1;
            row_col(q1, qx1, qy1);
            tess_prefetch_tile(&tm0, 32 * qx1, 32 * qy1);          // (clipped at the image edge by the unit)
            if (MODE == 2) tess_prefetch_tile(&tm1, 32 * qx1, 32 * qy1);
        }
#endif
        const int cnt_cur = (interior != EDGE) ? cnt : 0, off_cur = off;   // the other kernel's quadrants are skipped
        // cnt < 0 (builder overflow) and cnt > QC (list too long for the fast path) are done in the
        // tail below, spread over all warps
        if (cnt_cur > 0 && cnt_cur <= QC)
            tess_quadrant_body<MODE, HAS_SCALE, !EDGE>(a, sm, buf, labels, src, cnt_cur, qx0, qy0, lane, cur);
        // ---- rotate the pipeline
        q = q1; cnt = cnt1; off = off1;
        q1 = q2; cnt1 = cnt2; off1 = off2;
        buf ^= 1;
    }
    if (MODE != 0) tess_run_flush<NV>(cur.n > 0, cur, a.mom, lane);      // the runs still live
    // ---- tail: rows of the overflowed quadrants, spread over all warps of the grid (only in the
    //      instantiation that is always launched: the interior one for vector-friendly images)
    if (EDGE != (vec != 0)) {
        constexpr int PER_Q = 32 / TESS_RING_ROWS;
        const int n_items = a.st->n_brute * PER_Q;
        for (int item = (int)blockIdx.x * NW + wib; item < n_items; item += stride) {
            const int qq = a.L.ring_q[item / PER_Q];
            const int qy = qq / a.L.nqx, qx = qq - qy * a.L.nqx;
            const long r0 = 32L * qy + TESS_RING_ROWS * (item % PER_Q);
            const int cn = a.L.q_cnt[qq];
            if (cn < 0) {
                tess_rows_ring<MODE, HAS_SCALE>(a, labels, 32L * qx, r0);
            } else {       // a list beyond the fast path: every pixel scans it (from L2)
                const int o = a.L.q_off[qq];
                tess_rows_scan<MODE, HAS_SCALE>(a, labels, 32L * qx, r0, a.L.xy + o, a.L.gid + o,
                                                HAS_SCALE ? a.L.inv + o : nullptr, cn);
            }
        }
    }
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
