// tess_common.cuh -- shared definitions for the sm_100a kernels of libtess_b200.so.
//
// Geometry of the pruning hierarchy (image mode):
//   tile        64 x 64 px   one warp of k_build_qlists: box search over the generator grid, then
//                            the candidate lists of its four quadrants
//   quadrant    32 x 32 px   one warp of k_image_main: list staged in shared memory (cp.async)
//   block       16 x 16 px   16 lanes; intermediate filter for long quadrant lists
//   warp tile   32 x  8 px   one warp step; each lane owns a 2-column x 4-row patch (8 px),
//                            so every row of a warp tile is 16 lanes x 16 B = 256 B contiguous
//   leaf         8 x  8 px   8 lanes (4 across x 2 down); shortest candidate list
// A generator g can be the nearest one somewhere in a rectangle T only if
//   mind2(g,T) <= min_h maxd2(h,T)        (times the scale weights in WVT mode);
// for a disc of radius rho around c the sharper form is |c-g| <= min_h |c-h| + 2 rho.
// Each level keeps exactly those (plus a 2^-39 relative slack that covers every fp64 rounding
// difference between these bounds and the pinned per-pixel formula).
#pragma once
#include <stdint.h>
#include <math.h>

static long g_tess_launches = 0;   // kernels launched by the library (host-side count)

#ifdef TESS_EMU
#define TESS_LAUNCH(kern, grid, block, stream, ...) (++g_tess_launches, emu::launch(kern, dim3(grid), dim3(block), __VA_ARGS__))
#else
#include <cuda_runtime.h>
#define TESS_LAUNCH(kern, grid, block, stream, ...) (++g_tess_launches, kern<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__))
#endif

#define TESS_TS 64            // tile edge, px
#define TESS_CAP 512           // candidates of one tile staged in shared memory
#ifndef TESS_LEAF_CAP
#define TESS_LEAF_CAP 24      // candidates of one 8x8 leaf
#endif
#define TESS_MAIN_THREADS 64  // image main kernel: 2 independent warps per CTA
#ifndef TESS_MAIN_MINBLOCKS
#define TESS_MAIN_MINBLOCKS 8 // CTAs per SM the main kernel is compiled for (<= 128 registers)
#endif
#ifndef TESS_MAIN_MINBLOCKS_SN
#define TESS_MAIN_MINBLOCKS_SN 6 // same for the signal/noise mode (two maps in flight: <= 168 registers)
#endif
#define TESS_FULL 0xffffffffu

// A TMA tensor map (CUtensorMap: 128 opaque bytes, 64-byte aligned) of a pixel map [H][W] fp64 with a
// 32x32-px box -- the quadrant tile.  Encoded on the host (tess_b200.cu::make_tile_map), passed to the
// image kernel as a __grid_constant__ parameter.
struct alignas(64) TessTileMap {
    unsigned long long opaque[16];
};
#ifdef TESS_EMU
#define TESS_GRID_CONSTANT
#else
#define TESS_GRID_CONSTANT __grid_constant__
#endif
#define TESS_NAN (__longlong_as_double(0x7ff8000000000000LL))

// relative slack on bound comparisons, 1 + 2^-39
#define TESS_SLACK 1.0000000000018189894035458564758300781250

struct TessRect {
    double x0, x1, y0, y1;   // closed ranges of point coordinates
};

// Uniform grid of generators (counting sort).  Cell (i,j) holds generators with
// floor((gx-ox)/cs) == i (clamped to [0,gw-1]: border cells extend to infinity).
struct TessGrid {
    double ox, oy, cs, inv_cs, tol;
    int gw, gh;
};

// Device-resident iteration state (one per context).
struct TessState {
    double delta;                    // sum |old-new|^2 of the last update   (lloyd.pyx:76-81)
    double min_inv_s2;               // min_j 1/scale_j^2 (1 when unscaled): bounds the search radius
    unsigned long long min_inv_next; // bit pattern of the running min while scales are being updated
    unsigned long long n_changed;    // labels changed in the last accumulate (valid if counted)
    long long iters_done;            // loop bodies executed since set_generators
    long long conv_iter;             // iteration index (0-based) at which the latch closed
    int count_changed;               // 1 if some bin's population differs from the previous iteration
    int counted;                     // 1 if n_changed was actually counted this iteration
    int converged;                   // latch: once set every kernel of later iterations is a no-op
    int flags;                       // TESS_FLAG_*; non-zero also stops the iteration kernels
    int tile_counter;                // quadrant queue of k_image_main (interior); also the scratch of k_cell_occupancy_max
    int have_prev;                   // previous-iteration labels exist
    int n_brute;                     // quadrants off the fast path of the image kernel (entries of ring_q)
    int n_overflow;                  // of those: tiles that overflowed the list builder (per-pixel ring search)
    unsigned list_cursor;            // bump allocator of the candidate-list arena
    unsigned spare_;                 // quadrant queue of k_image_main (edge instantiation)
    unsigned update_ran;             // k_update_nodes did work this iteration (read by k_update_finish)
    int last_parity;                 // label buffer (0/1) written by the most recent accumulate
    int flags_next;                  // flags raised by k_update_nodes; k_update_finish folds them into `flags`
};
#define TESS_FLAG_NONFINITE 1
// iteration kernels are no-ops once the latch closed or an error flag is up (unless forced)
#define TESS_STOPPED(st, honor) ((honor) && ((st)->converged | (st)->flags))

// max over the rectangle of the squared distance to (gx,gy) -- an upper bound (to rounding)
__host__ __device__ __forceinline__ double tess_maxd2(const TessRect& r, double gx, double gy) {
    double ax = fmax(fabs(gx - r.x0), fabs(gx - r.x1));
    double ay = fmax(fabs(gy - r.y0), fabs(gy - r.y1));
    return ax * ax + ay * ay;
}
// min over the rectangle of the squared distance to (gx,gy)
__host__ __device__ __forceinline__ double tess_mind2(const TessRect& r, double gx, double gy) {
    double ax = fmax(fmax(r.x0 - gx, gx - r.x1), 0.0);
    double ay = fmax(fmax(r.y0 - gy, gy - r.y1), 0.0);
    return ax * ax + ay * ay;
}

__host__ __device__ __forceinline__ int tess_cell_coord(double g, double o, double inv_cs, int n) {
    double f = floor((g - o) * inv_cs);
    f = fmin(fmax(f, 0.0), (double)(n - 1));   // NaN -> 0
    return (int)f;
}

#ifndef TESS_EMU
#define TESS_DEVFN __device__ __forceinline__
#define TESS_NOINLINE __device__ __noinline__
#else
#define TESS_DEVFN inline
#define TESS_NOINLINE static __attribute__((noinline))
#endif

TESS_DEVFN double tess_warp_min(double v) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v = fmin(v, __shfl_xor_sync(TESS_FULL, v, m));
    return v;
}
TESS_DEVFN double tess_warp_sum(double v) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(TESS_FULL, v, m);
    return v;
}
TESS_DEVFN unsigned tess_lanemask_lt(int lane) { return (1u << lane) - 1u; }

// Weight of a signal / noise pixel, w = (S / N)^2.  The quotient comes from a reciprocal seed of the
// special-function unit refined by two Newton steps (~1 ulp): a correctly rounded fp64 division is
// ~30 instructions, eight of them per lane and warp tile were a fifth of the S/N-mode kernel.  The
// weight enters the per-bin SUMS only (tolerance 1e-12), never the labels.  N = 0, Inf or NaN (and
// denormal N, flushed to zero by the seed) give a non-finite weight: the pixel is masked.
TESS_DEVFN double tess_sn_weight(double S, double N) {
#ifndef TESS_EMU
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(N));
    r = fma(fma(-N, r, 1.0), r, r);
    r = fma(fma(-N, r, 1.0), r, r);
    const double q = S * r;
    return q * q;
#else
    const double q = S / N;
    return q * q;
#endif
}

// The pinned squared distance: fl(fl(dx*dx) + fl(dy*dy)), never contracted to an FMA.
TESS_DEVFN double tess_d2_pinned(double px, double py, double gx, double gy) {
    double dx = __dsub_rn(px, gx);
    double dy = __dsub_rn(py, gy);
    return __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
}

