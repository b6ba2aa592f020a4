/*
 * include/tess_b200.h -- C ABI of libtess_b200.so, the B200 (sm_100a) implementation of the
 * Lloyd / centroidal-Voronoi (CVT, WVT) hot path of jonathansick/tess.
 *
 * The reference has no C ABI for this path: it is a Python-callable Cython function,
 *     cpdef lloyd(double[:, :] xy, double[:] w, double[:, :] node_xy, long max_iters)
 *                                                       (reference tess/lloyd.pyx:10-11)
 * called from exactly one site (reference tess/voronoi.py:301-302), plus SciPy nearest-
 * neighbour calls for the segmentation map (tess/voronoi.py:113-114) and point partition
 * (tess/voronoi.py:171-172).  Its historic ctypes binding (reference
 * examples/test_clloyds.py:21-39: `cvt.lloyd(n, x, y, w, nNode, xNode, yNode, vBinNum)`,
 * plain pointers, status return) is the shape this header echoes.  Each entry point below
 * names the reference lines it replaces.  INTEGRATION.md shows the ctypes stub a tess
 * maintainer would add.
 *
 * Conventions
 *   - every function (bar tess_version, tess_last_error, tess_launch_count) returns 0 (TESS_OK)
 *     or a negative TESS_ERR_* code; never throws; tess_last_error() gives a thread-local
 *     message for the last failure.
 *   - ALL data pointers are DEVICE pointers on the context's device (e.g. torch CUDA tensor
 *     data_ptr()).  Input arrays are BORROWED: they must stay alive and unchanged until they
 *     are replaced by another tess_set_* call or the context is destroyed.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Calls are
 *     asynchronous on that stream unless stated (tess_lloyd_run and tess_get_iter_stats
 *     synchronise it because they return host values).
 *   - one context per device; a context is not thread-safe, distinct contexts are.
 *   - labels are int32 (the reference returns int64, widened by the Python layer);
 *     coordinates, weights and moments are fp64.
 *
 * Arithmetic contract (bit-exact with the reference on tie-free input; see oracle/oracle.py):
 *   d2(i,j) = fl( fl(dx*dx) + fl(dy*dy) )   fp64, no FMA        [cKDTree via lloyd.pyx:54-55]
 *   label_i = argmin_j d2(i,j) [* inv_s2_j when scales are given], ties -> lowest j
 *   moments per bin: count, sum w, sum w*x, sum w*y (+ sum S, sum N*N in signal/noise mode),
 *     products rounded separately; the summation ORDER differs from the reference's
 *     sequential loop (lloyd.pyx:58-65), so sums agree to ~1e-15 relative, not bitwise.
 */
#ifndef TESS_B200_H
#define TESS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tess_ctx tess_ctx;

enum {
    TESS_OK = 0,
    TESS_ERR_ARG = -1,       /* bad argument (null pointer, non-positive size, ...)            */
    TESS_ERR_CUDA = -2,      /* CUDA runtime failure; text in tess_last_error()                */
    TESS_ERR_STATE = -3,     /* call order violated (no image/points or no generators set)     */
    TESS_ERR_NONFINITE = -4, /* a generator coordinate / scale is NaN or Inf.  The reference   */
                             /* raises ValueError from cKDTree in that case (bin with count>0  */
                             /* and sum w == 0, lloyd.pyx:66-70 -> :54 next iteration)         */
    TESS_ERR_NOMEM = -5      /* device allocation failed                                       */
};

/* Moment columns of tess_get_moments / tess_moments_ptr. */
enum {
    TESS_MOM_COUNT = 0, /* pixels/points in the bin   (lloyd.pyx:62 node_population)          */
    TESS_MOM_W = 1,     /* sum w                      (lloyd.pyx:63 weight_sum)               */
    TESS_MOM_WX = 2,    /* sum fl(w*x)                (lloyd.pyx:64-65)                       */
    TESS_MOM_WY = 3,    /* sum fl(w*y)                (lloyd.pyx:64-65)                       */
    TESS_MOM_S = 4,     /* sum signal                 (pixel_accretion.py:513)                */
    TESS_MOM_N2 = 5     /* sum fl(noise*noise)        (pixel_accretion.py:514)                */
};

/* Options for tess_set_option. */
enum {
    TESS_OPT_WVT_SCALE_UPDATE = 1, /* 0 (default): scales stay as given.  1: after each node   */
                                   /* update set scale_j = sqrt(count_j / (S/N)_j) (Diehl &   */
                                   /* Statler 2006 eq. 4); needs signal/noise mode.  No        */
                                   /* reference counterpart.              */
    TESS_OPT_UNIFORM_WEIGHT = 2,   /* signal/noise mode only. 0 (default): w = fl(fl(S/N)^2)  */
                                   /* (scripts/demo_iso_sn_voronoi.py:26,37). 1: w = 1         */
                                   /* (geometric centroids, WVT).                              */
    TESS_OPT_KERNEL_TIMING = 3,    /* 1: bracket the dominant kernel of tess_lloyd_accumulate  */
                                   /* (pixel / point assignment + sums) with CUDA events on    */
                                   /* its stream; read with tess_get_kernel_timing.            */
    TESS_OPT_OWNER_UPDATE = 4      /* multi-GPU, tess_peer_allreduce_sparse. 1: the rank that    */
                                   /* owns a slice of the moment vector forms the new nodes of  */
                                   /* its bins from the totals (lloyd.pyx:66-74) and broadcasts */
                                   /* 16 bytes per bin instead of the 32-byte row; the totals   */
                                   /* stay in the owner's copy of its slice; tess_lloyd_update  */
                                   /* takes the broadcast nodes.  Needs 2 k more doubles in the */
                                   /* caller's buffer (>= 6 k + 2 + 34); applies to 4-moment    */
                                   /* rows without WVT scale updates, else the rows travel.     */
};

int tess_version(void);
const char* tess_last_error(void);

int tess_ctx_create(int device, tess_ctx** out);
int tess_ctx_destroy(tess_ctx* ctx);
int tess_set_option(tess_ctx* ctx, int option, long value);

/*
 * Image mode -- replaces the point-set construction of CVTessellation.from_image
 * (reference tess/voronoi.py:281-287): pixel (row r, col c) is the point (x0 + c, y0 + r);
 * it is never materialised.  H x W row-major fp64 maps.  Give EITHER `density` (CVT parity
 * mode: w = density, 4 moments) OR `signal` and `noise` (S/N mode: w = (S/N)^2, 6 moments).
 * Pixels whose weight (or signal/noise) is non-finite are excluded from every sum, exactly
 * like voronoi.py:285 drops them, but still receive a label (so the label map doubles as the
 * segmap of voronoi.py:75-81).  For a row-sharded multi-GPU run each rank passes its own
 * rows with y0 = first global row.
 */
int tess_set_image(tess_ctx* ctx, long H, long W, double x0, double y0, const double* density,
                   const double* signal, const double* noise, void* stream);

/*
 * Optional: the rectangle [x0, x1] x [y0, y1] that contains (nearly) all generators.  The uniform
 * grid that indexes the generators covers the union of this rectangle and the pixel window;
 * generators outside it are clamped into its border cells (still exact, but slow if many).  A
 * row-sharded rank passes the FULL image here so that the replicated generator table is binned
 * properly although the rank's own window is only a slab.  Default: the pixel window itself.
 * x1 < x0 clears it.
 */
int tess_set_domain(tess_ctx* ctx, double x0, double y0, double x1, double y1);

/*
 * Point mode -- the `xy`, `w` arguments of lloyd() (reference tess/lloyd.pyx:10-11):
 * xy is AoS [n][2], w is [n], both BORROWED for as long as they are set (no copy is made); the
 * call computes their bounding box (one pass, one host synchronisation).  Labels are returned in
 * the caller's point order.
 */
int tess_set_points(tess_ctx* ctx, long n, const double* xy, const double* w, void* stream);

/*
 * Generators -- the `node_xy` argument of lloyd() ([k][2], (x, y) order, voronoi.py:273-277).
 * COPIED into the context (the reference does not mutate its input either, lloyd.pyx:49,59).
 * `scale` ([k], nullable) are WVT scale lengths; NULL = plain Euclidean (reference parity).
 * Resets the iteration counter and convergence state.  TESS_ERR_NONFINITE if any is NaN/Inf.
 */
int tess_set_generators(tess_ctx* ctx, long k, const double* node_xy, const double* scale,
                        void* stream);

/*
 * One Lloyd iteration = tess_lloyd_accumulate + tess_lloyd_update.
 *
 * accumulate: rebuild the generator grid index, assign every pixel/point to its nearest
 *   generator (lloyd.pyx:54-55) and form the per-bin moments (lloyd.pyx:58-65) -- LOCAL to
 *   this context's pixels.  Also evaluates "did any label change since the last iteration".
 * tess_moments_ptr: device pointer to the packed vector [k*M moments | tail], M = 4 or 6;
 *   `n_doubles` is its full length.  A multi-GPU host sums THIS vector across ranks
 *   (one allreduce) between accumulate and update; nothing else crosses GPUs.
 * update: nodes = (sum wx, sum wy)/sum w where count > 0 else (0,0) (lloyd.pyx:66-74),
 *   delta = sum |old-new|^2 (lloyd.pyx:76-81), convergence latch (see tess_lloyd_run).
 */
int tess_lloyd_accumulate(tess_ctx* ctx, void* stream);
int tess_moments_ptr(tess_ctx* ctx, double** dev_ptr, long* n_doubles, int* n_moments);
int tess_lloyd_update(tess_ctx* ctx, void* stream);

/*
 * The exchange step without NCCL (multi-GPU, one NVSwitch node).  tess_set_moment_buffer makes the
 * context keep its moment vector in caller memory -- a SYMMETRIC allocation of the same size on
 * every rank (>= 6 k + 2 doubles, 16-byte aligned; NULL returns to a library-owned vector).
 * tess_peer_allreduce then sums that vector over the ranks in place, in one kernel of this
 * library: with `multicast_ptr` (an NVLS multicast mapping of the allocation) the switch does the
 * reduction (multimem.ld_reduce / multimem.st); otherwise `peer_ptrs[0..world-1]` (HOST array of
 * the peer-mapped device addresses, own rank included) are read and written over NVLink.  All
 * ranks end with bit-identical sums.  The caller provides the cross-rank ordering: a device-side
 * barrier over all ranks before the call (stream-ordered after tess_lloyd_accumulate) and after it
 * (before tess_lloyd_update).  Replaces: nothing in the reference (single process, lloyd.pyx:58-65
 * sums in one loop); it is the collective of SURVEY.md section 8e.
 */
int tess_set_moment_buffer(tess_ctx* ctx, double* buf, long n_doubles);
int tess_peer_allreduce(tess_ctx* ctx, int world, int rank, double* const* peer_ptrs, double* multicast_ptr,
                        long n_doubles, void* stream);

/*
 * Sparse form of the same exchange, for generator tables numbered in spatial order (rows of a rank's
 * pixel shard then touch one short range of bins).  The caller's buffer must hold 34 more doubles
 * than the vector (>= k*M + 2 + 34, the same size on every rank): 32 barrier words and the range
 * [first, last + 1) of bins with a non-zero local count, which tess_lloyd_accumulate publishes there
 * (tess_peer_publish_range does it again on request).  tess_peer_allreduce_sparse is ONE kernel that
 * carries its own cross-rank barriers (so it needs no barrier from the caller; every rank must launch
 * it once per iteration): it reads a row only from the peers whose published range holds it, adds in
 * rank order and stores the totals into every rank's copy -- bit-identical to the dense sum, since
 * the skipped rows are exact zeros.  The range is reset by tess_lloyd_update.  `n_doubles` is the
 * summed length (tess_moments_ptr).
 */
int tess_peer_publish_range(tess_ctx* ctx, void* stream);
/*
 * Optional: `multicast_ptr` = an NVLS multicast mapping of the caller's moment buffer (same offset 0 as
 * the buffer given to tess_set_moment_buffer).  The broadcast half of tess_peer_allreduce_sparse then
 * goes through the switch (one multimem.st per pair of totals reaches every rank's copy) instead of
 * one bulk copy per destination rank; the reduce half is unchanged, the sums are the same bits.
 * NULL (the default) returns to peer-to-peer copies.
 */
int tess_peer_set_multicast(tess_ctx* ctx, double* multicast_ptr);
int tess_peer_allreduce_sparse(tess_ctx* ctx, int world, int rank, double* const* peer_ptrs, long n_doubles, void* stream);

/* n_steps iterations back to back, no host synchronisation (benchmarks, fixed-step use). */
int tess_lloyd_step(tess_ctx* ctx, long n_steps, void* stream);

/*
 * The reference loop and stop rule (lloyd.pyx:47-90): iterate until the node table is a
 * fixed point (reference: delta == 0; here, equivalently and robustly to summation order:
 * no label changed since the previous iteration, or delta == 0) -> *converged = 1, or until
 * max_iters + 2 loop bodies have run -> *converged = 0.  *n_iters = loop bodies executed.
 * Synchronises `stream`.  Afterwards tess_get_generators gives the reference's returned
 * node_xy and tess_get_labels its returned idx (assignment w.r.t. the previous node table,
 * lloyd.pyx:55 vs :86/:88).
 */
int tess_lloyd_run(tess_ctx* ctx, long max_iters, int* converged, long* n_iters, void* stream);

/* Results (device-to-device copies into caller memory, asynchronous on `stream`). */
int tess_get_labels(tess_ctx* ctx, int32_t* out, void* stream);      /* [H*W] or [n]          */
int tess_get_moments(tess_ctx* ctx, double* out, void* stream);      /* [k][6], unused cols 0 */
int tess_get_generators(tess_ctx* ctx, double* out, void* stream);   /* [k][2]                */
int tess_get_scales(tess_ctx* ctx, double* out, void* stream);       /* [k]                   */
/*
 * delta of the last update (lloyd.pyx:82's printed value), number of labels that changed in
 * the last accumulate (-1 = "some changed, not counted": counting is skipped when a per-bin
 * population already differs), iterations done since tess_set_generators.  Synchronises.
 */
int tess_get_iter_stats(tess_ctx* ctx, double* delta, long* n_changed, long* n_iters_done,
                        void* stream);

/* Latch and error state: *converged = 1 once the fixed point was reached, *nonfinite = 1 if a
 * generator became NaN/Inf (see TESS_ERR_NONFINITE).  Synchronises `stream`. */
int tess_get_status(tess_ctx* ctx, int* converged, int* nonfinite, void* stream);

/*
 * Measurement aids (bench.py).  tess_get_kernel_timing synchronises the device and returns the
 * summed device time (ms) and number of launches of the dominant kernel recorded since the last
 * call (needs TESS_OPT_KERNEL_TIMING; at most 512 launches are kept between calls).
 * tess_launch_count returns how many kernels this library has launched in this process.
 */
int tess_get_kernel_timing(tess_ctx* ctx, double* ms_total, long* launches);
/* Diagnostics of the last list build: out[0] = quadrants handed to the per-pixel ring search (their
 * tile overflowed the list builder: bins of a few pixels), out[1] = arena records used, out[2] = arena capacity, out[3] = quadrants whose
 * list exceeded the fast path (scan of the list).  Synchronises the device. */
int tess_get_list_stats(tess_ctx* ctx, long* out4);
long tess_launch_count(void);

/*
 * Assignment only -- the segmentation map of VoronoiTessellation.segmap /
 * render_voronoi_field (reference tess/voronoi.py:75-114): labels_out[r*W + c] = nearest
 * generator of (x0 + c, y0 + r).  Uses the context's generators; no image needs to be set.
 */
int tess_assign_grid(tess_ctx* ctx, double x0, double y0, long H, long W, int32_t* labels_out,
                     void* stream);
/* partition_points (reference tess/voronoi.py:154-173): out[i] = nearest generator of xy[i]. */
int tess_assign_points(tess_ctx* ctx, long m, const double* xy, int32_t* out, void* stream);

/*
 * Per-bin sums over a label array -- np.bincount(segmap.ravel()) of compute_cell_areas
 * (reference tess/voronoi.py:150) and np.bincount(idx, weights=mass) of sum_cell_point_mass
 * (tess/voronoi.py:195): out[j] = sum of weights[i] (or of 1 when weights is NULL) over labels[i]
 * == j.  Labels outside [0, k) are skipped (flagged / masked pixels).  out is [k].
 */
int tess_label_sums(tess_ctx* ctx, long n, const int32_t* labels, const double* weights, long k,
                    double* out, void* stream);

/*
 * Per-label moments of a row-major [H][W] label image in one pass, pixel coordinates implicit:
 * out[l] = (count, sum w, sum w*row, sum w*col) for labels 0 <= l < k (others ignored).  A pixel
 * always counts; its weight only if finite; weights == NULL means w = 1.  This is the sum behind
 * PixelAccretor.centroids (reference tess/pixel_accretion.py:212-240) -- no product array w*row /
 * w*col is ever built.  out is [k][4] device memory.
 */
int tess_label_moments(tess_ctx* ctx, long H, long W, const int32_t* labels, const double* weights, long k,
                       double* out, void* stream);

/*
 * Validation aids: exhaustive O(n*k) assignment on the GPU with the same pinned arithmetic,
 * no index, no pruning.  Used by the tests to check the tiled kernels at sizes the CPU oracle
 * cannot reach.  Not a fallback: nothing in the library calls them.
 */
int tess_assign_grid_exhaustive(tess_ctx* ctx, double x0, double y0, long H, long W,
                                int32_t* labels_out, void* stream);
int tess_assign_points_exhaustive(tess_ctx* ctx, long m, const double* xy, int32_t* out,
                                  void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TESS_B200_H */
