// tess_b200.cu -- host side of libtess_b200.so: context, workspace and the C ABI declared in
// include/tess_b200.h.  Built for sm_100a only (tess_b200/build.py); there is no CPU path.
//
// One Lloyd iteration (reference tess/lloyd.pyx:47-90) is the launch sequence
//   k_cell_count -> k_cell_blocksum -> k_cell_scan -> k_cell_scatter   generator grid (replaces cKDTree build, :54)
//   k_build_qlists                                       per-quadrant candidate lists (arena records, any order)
//   k_image_main | k_points_main                         assignment (:55) fused with the sums (:58-65)
//   k_count_prefilter -> k_compare_labels                "did any label change" (convergence, :85)
//   [multi-GPU: the host all-reduces the moment vector here]
//   k_update_nodes                                       node update, delta, latch (:66-90)
// with no host synchronisation in between; the convergence latch lives on the device.
#include "../../include/tess_b200.h"

#include "tess_common.cuh"
#include "tess_index.cuh"
#include "tess_image.cuh"
#include "tess_points.cuh"
#include "tess_update.cuh"
#include "tess_phases.cuh"
#include "tess_peer.cuh"

#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define TESS_VERSION_NUM 100

static thread_local char g_err[512] = "";

static int tess_fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CU_TRY(expr)                                                                              \
    do {                                                                                          \
        cudaError_t e_ = (expr);                                                                  \
        if (e_ != cudaSuccess)                                                                    \
            return tess_fail(e_ == cudaErrorMemoryAllocation ? TESS_ERR_NOMEM : TESS_ERR_CUDA,    \
                             "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

#define TRY(expr)               \
    do {                        \
        int r_ = (expr);        \
        if (r_ != TESS_OK) return r_; \
    } while (0)

enum { SRC_NONE = 0, SRC_IMAGE = 1, SRC_POINTS = 2 };

struct tess_ctx {
    int device;
    int sm_count;
    // ---- borrowed inputs
    int source;
    int img_mode;   // 1: density, 2: signal + noise
    long H, W;
    double x0, y0;
    const double *dens, *sig, *noise;
    long n_pts;
    const double *pts_xy, *pts_w;
    // the point set in spatial order (owned; built by tess_set_points)
    int* pts_perm;
    double2* pts_xy_s;
    double* pts_w_s;
    long pts_cap;
    double pb[4];   // bounding box of the points (x0, x1, y0, y1)
    double gb[4];   // bounding box of the generator table at tess_set_generators
    int gb_valid;
    int has_domain;
    double dom[4];  // generator domain (x0, y0, x1, y1)
    // ---- generators (owned)
    int k, k_cap;
    double* node;     // [k][2]
    double* scale;    // [k] or unused
    double* inv_s2;   // [k]
    int has_scale;
    // ---- options
    int wvt_update, uniform_weight;
    // ---- workspace (owned)
    TessState* st;
    int *cell_cnt, *cell_start, *cell_of, *cell_items, *block_sum;
    double2* cell_xy;
    double* cell_inv;
    long cells_cap;
    TessQLists ql;      // per-quadrant candidate lists (arena)
    int* tile_ring;     // per tile: box-search ring that sufficed last time (a hint, not state)
    long quads_cap;
    int arena_has_inv;
    int32_t* lab[2];
    long lab_cap;
    double* mom;
    long mom_cap;
    int mom_external;   // mom is caller memory (tess_set_moment_buffer): never freed or grown here
    double* prev_count;
    double* scratch;   // 8 doubles (bounding-box reduction)
    double* delta_part; // per-CTA partial deltas of k_update_nodes
    int cell_shift;     // image grids: cell = 64 px >> cell_shift; -1 = not decided for geo_key yet
    int regrid_tried;   // the grid was re-decided after an overflow since the last tess_set_generators
    double geo_key[4];  // window (x0, y0, H, W) the shift was decided for
    TessTileMap tmap[2];   // TMA descriptors of the pixel map(s) (32x32-px box), see make_tile_map
    int tmap_ok;
    const double* tmap_src[2];
    long tmap_hw[2];
    unsigned long long exchange_epoch;   // sparse exchanges launched so far (the barrier word of the next one)
    double* sparse_mc;                   // multicast mapping of the caller's moment buffer (tess_peer_set_multicast) or null
    int owner_update;                    // TESS_OPT_OWNER_UPDATE: the sparse exchange broadcasts owner-computed nodes
    int staged_valid;                    // the node area of the caller's buffer holds this iteration's new nodes
    int index_valid;   // the generator grid in the workspace is that of the CURRENT nodes for index_geo
    TessGrid index_geo;
    int phase_grid;    // CTAs of k_iteration_phases (all resident: cooperative launch); 0 = not decided
    int coop_ok;
    int timing;        // TESS_OPT_KERNEL_TIMING
#ifndef TESS_EMU
    cudaEvent_t ev[2 * 512];
    int n_ev_alloc, n_ev_used;
#endif
    TessState* h_st;   // pinned mirror
};

static cudaStream_t S(void* s) { return (cudaStream_t)s; }

static int bind(tess_ctx* c) {
    if (!c) return tess_fail(TESS_ERR_ARG, "null context");
    CU_TRY(cudaSetDevice(c->device));
    return TESS_OK;
}

// ------------------------------------------------------------------------------------ misc
extern "C" int tess_version(void) { return TESS_VERSION_NUM; }
extern "C" const char* tess_last_error(void) { return g_err; }

extern "C" int tess_ctx_create(int device, tess_ctx** out) {
    if (!out) return tess_fail(TESS_ERR_ARG, "tess_ctx_create: out is null");
    *out = nullptr;
    int ndev = 0;
    CU_TRY(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return tess_fail(TESS_ERR_ARG, "tess_ctx_create: no CUDA device %d", device);
    CU_TRY(cudaSetDevice(device));
    tess_ctx* c = new tess_ctx();
    memset(c, 0, sizeof(*c));
    c->device = device;
    int r = TESS_OK;
    do {
        cudaError_t e = cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device);
        if (e != cudaSuccess) { r = tess_fail(TESS_ERR_CUDA, "device attribute: %s", cudaGetErrorString(e)); break; }
        e = cudaMalloc((void**)&c->st, sizeof(TessState));
        if (e != cudaSuccess) { r = tess_fail(TESS_ERR_NOMEM, "cudaMalloc(state): %s", cudaGetErrorString(e)); break; }
        cudaMemset(c->st, 0, sizeof(TessState));
        e = cudaMalloc((void**)&c->scratch, 8 * sizeof(double));
        if (e != cudaSuccess) { r = tess_fail(TESS_ERR_NOMEM, "cudaMalloc(scratch): %s", cudaGetErrorString(e)); break; }
        cudaMemset(c->scratch, 0, 8 * sizeof(double));      // [6]: "CTAs done" counter of the sparse exchange
        e = cudaMalloc((void**)&c->delta_part, sizeof(double) * (size_t)(c->sm_count * 16 + 16));
        if (e != cudaSuccess) { r = tess_fail(TESS_ERR_NOMEM, "cudaMalloc(delta_part): %s", cudaGetErrorString(e)); break; }
        e = cudaMallocHost((void**)&c->h_st, sizeof(TessState) + 8 * sizeof(double));
        if (e != cudaSuccess) { r = tess_fail(TESS_ERR_NOMEM, "cudaMallocHost: %s", cudaGetErrorString(e)); break; }
    } while (0);
    if (r != TESS_OK) {
        tess_ctx_destroy(c);
        return r;
    }
    *out = c;
    return TESS_OK;
}

extern "C" int tess_ctx_destroy(tess_ctx* c) {
    if (!c) return TESS_OK;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    void* ptrs[] = {c->node, c->scale, c->inv_s2, c->st, c->cell_cnt, c->cell_start, c->block_sum, c->cell_of, c->cell_items,
                    c->cell_xy, c->cell_inv, c->ql.q_off, c->ql.q_cnt, c->ql.ring_q, c->ql.xy, c->ql.gid, c->ql.inv, c->tile_ring, c->lab[0], c->lab[1],
                    c->mom_external ? nullptr : c->mom, c->prev_count, c->scratch, c->delta_part, c->pts_perm, c->pts_xy_s,
                    c->pts_w_s};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    if (c->h_st) cudaFreeHost(c->h_st);
#ifndef TESS_EMU
    for (int i = 0; i < c->n_ev_alloc; ++i) cudaEventDestroy(c->ev[i]);
#endif
    delete c;
    return TESS_OK;
}

extern "C" int tess_set_option(tess_ctx* c, int option, long value) {
    TRY(bind(c));
    switch (option) {
        case TESS_OPT_WVT_SCALE_UPDATE: c->wvt_update = value ? 1 : 0; return TESS_OK;
        case TESS_OPT_UNIFORM_WEIGHT: c->uniform_weight = value ? 1 : 0; return TESS_OK;
        case TESS_OPT_KERNEL_TIMING: c->timing = value ? 1 : 0; return TESS_OK;
        case TESS_OPT_OWNER_UPDATE: c->owner_update = value ? 1 : 0; return TESS_OK;
        default: return tess_fail(TESS_ERR_ARG, "tess_set_option: unknown option %d", option);
    }
}

static int n_moments(const tess_ctx* c) { return (c->source == SRC_IMAGE && c->img_mode == 2) ? 6 : 4; }

static int reset_iteration_state(tess_ctx* c, cudaStream_t s) {
    // iteration counter, latch and "previous labels" start over; error flags are kept by the caller
    TessState z;
    memset(&z, 0, sizeof(z));
    z.min_inv_s2 = 1.0;
    z.conv_iter = -1;
    memcpy(c->h_st, &z, sizeof(z));
    CU_TRY(cudaMemcpyAsync(c->st, c->h_st, sizeof(TessState), cudaMemcpyHostToDevice, s));
    CU_TRY(cudaStreamSynchronize(s));   // h_st is reused
    return TESS_OK;
}

__global__ void __launch_bounds__(256) k_fill(double* p, long n, double v) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// prev_count <- -1 and min 1/scale^2 (after the state block was reset)
static int restart_tail(tess_ctx* c, cudaStream_t s) {
    if (c->k > 0) {
        TESS_LAUNCH((k_fill), (c->k + 255) / 256, 256, s, c->prev_count, (long)c->k, -1.0);
        if (c->has_scale) TESS_LAUNCH((k_min_inv_s2), 1, 1024, s, c->st, c->inv_s2, c->k);
        CU_TRY(cudaGetLastError());
    }
    return TESS_OK;
}

// new inputs: the iteration starts over
static int restart(tess_ctx* c, cudaStream_t s) {
    TRY(reset_iteration_state(c, s));
    return restart_tail(c, s);
}

static int ensure_labels(tess_ctx* c, long n) {
    if (n <= c->lab_cap) return TESS_OK;
    for (int i = 0; i < 2; ++i) {
        if (c->lab[i]) cudaFree(c->lab[i]);
        c->lab[i] = nullptr;
    }
    c->lab_cap = 0;
    CU_TRY(cudaMalloc((void**)&c->lab[0], sizeof(int32_t) * (size_t)n));
    CU_TRY(cudaMalloc((void**)&c->lab[1], sizeof(int32_t) * (size_t)n));
    c->lab_cap = n;
    return TESS_OK;
}

// ------------------------------------------------------------------------------------ inputs
extern "C" int tess_set_image(tess_ctx* c, long H, long W, double x0, double y0, const double* density,
                              const double* signal, const double* noise, void* stream) {
    TRY(bind(c));
    if (H <= 0 || W <= 0) return tess_fail(TESS_ERR_ARG, "tess_set_image: H and W must be positive");
    if ((double)H * (double)W > 2.0e9 * 64.0) return tess_fail(TESS_ERR_ARG, "tess_set_image: image too large");
    const bool cvt = density && !signal && !noise, sn = !density && signal && noise;
    if (!cvt && !sn) return tess_fail(TESS_ERR_ARG, "tess_set_image: give either density or signal+noise");
    c->source = SRC_IMAGE;
    c->index_valid = 0;
    c->img_mode = cvt ? 1 : 2;
    c->H = H; c->W = W; c->x0 = x0; c->y0 = y0;
    c->dens = density; c->sig = signal; c->noise = noise;
    c->n_pts = 0; c->pts_xy = c->pts_w = nullptr;
    TRY(ensure_labels(c, H * W));
    TRY(restart(c, S(stream)));
    return TESS_OK;
}

extern "C" int tess_set_domain(tess_ctx* c, double x0, double y0, double x1, double y1) {
    TRY(bind(c));
    c->index_valid = 0;
    if (!(x1 >= x0) || !(y1 >= y0)) { c->has_domain = 0; c->cell_shift = -1; return TESS_OK; }
    if (!isfinite(x0) || !isfinite(y0) || !isfinite(x1) || !isfinite(y1))
        return tess_fail(TESS_ERR_ARG, "tess_set_domain: non-finite rectangle");
    c->has_domain = 1;
    c->cell_shift = -1;
    c->dom[0] = x0; c->dom[1] = y0; c->dom[2] = x1; c->dom[3] = y1;
    return TESS_OK;
}

// bounding box of a point set: block partial min/max -> atomics on the bit patterns is avoided by
// a single-CTA pass (n is at most a few million in point mode)
__global__ void __launch_bounds__(1024) k_points_bbox(const double* xy, long n, double* out) {
    __shared__ double s_v[4][32];
    __shared__ int s_bad;
    if (threadIdx.x == 0) s_bad = 0;
    __syncthreads();
    double x0 = INFINITY, x1 = -INFINITY, y0 = INFINITY, y1 = -INFINITY;
    bool bad = false;
    for (long i = threadIdx.x; i < n; i += blockDim.x) {
        double x = xy[2 * i], y = xy[2 * i + 1];
        if (isfinite(x) && isfinite(y)) {
            x0 = fmin(x0, x); x1 = fmax(x1, x); y0 = fmin(y0, y); y1 = fmax(y1, y);
        } else {
            bad = true;
        }
    }
    if (bad) s_bad = 1;
    x0 = tess_warp_min(x0); y0 = tess_warp_min(y0);
    x1 = -tess_warp_min(-x1); y1 = -tess_warp_min(-y1);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { s_v[0][w] = x0; s_v[1][w] = x1; s_v[2][w] = y0; s_v[3][w] = y1; }
    __syncthreads();
    if (w == 0) {
        const int nw = blockDim.x >> 5;
        x0 = l < nw ? s_v[0][l] : INFINITY; x1 = l < nw ? s_v[1][l] : -INFINITY;
        y0 = l < nw ? s_v[2][l] : INFINITY; y1 = l < nw ? s_v[3][l] : -INFINITY;
        x0 = tess_warp_min(x0); y0 = tess_warp_min(y0);
        x1 = -tess_warp_min(-x1); y1 = -tess_warp_min(-y1);
        if (l == 0) { out[0] = x0; out[1] = x1; out[2] = y0; out[3] = y1; out[4] = s_bad ? 1.0 : 0.0; }
    }
}

static int points_bbox(tess_ctx* c, long n, const double* xy, double* box, cudaStream_t s, bool must_be_finite = true) {
    TESS_LAUNCH((k_points_bbox), 1, 1024, s, xy, n, c->scratch);
    CU_TRY(cudaGetLastError());
    double* h = (double*)(c->h_st + 1);
    CU_TRY(cudaMemcpyAsync(h, c->scratch, 5 * sizeof(double), cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaStreamSynchronize(s));
    // the reference's cKDTree.query raises ValueError("'x' must be finite") here (lloyd.pyx:55)
    if (must_be_finite && h[4] != 0.0)
        return tess_fail(TESS_ERR_NONFINITE, "point coordinates must be finite (NaN or Inf found)");
    memcpy(box, h, 4 * sizeof(double));
    if (!(box[0] <= box[1]) || !(box[2] <= box[3])) { box[0] = box[2] = 0.0; box[1] = box[3] = 1.0; }
    return TESS_OK;
}

static int ensure_cells(tess_ctx* c, int n_cells);

// The point set in spatial order (counting sort over a grid of its own, ~16 points per cell): done
// once, the points never move.  The Lloyd kernel then walks neighbouring points with neighbouring
// threads (shared generator cells, shared bins).  Uses the generator-grid workspace as scratch.
static int sort_points(tess_ctx* c, cudaStream_t s) {
    const long n = c->n_pts;
    if (n > c->pts_cap) {
        void** ps[] = {(void**)&c->pts_perm, (void**)&c->pts_xy_s, (void**)&c->pts_w_s};
        for (void** p : ps) { if (*p) cudaFree(*p); *p = nullptr; }
        c->pts_cap = 0;
        CU_TRY(cudaMalloc((void**)&c->pts_perm, sizeof(int) * (size_t)n));
        CU_TRY(cudaMalloc((void**)&c->pts_xy_s, sizeof(double2) * (size_t)n));
        CU_TRY(cudaMalloc((void**)&c->pts_w_s, sizeof(double) * (size_t)n));
        c->pts_cap = n;
    }
    TessGrid g;
    const double w = fmax(c->pb[1] - c->pb[0], 0.0), h = fmax(c->pb[3] - c->pb[2], 0.0);
    double cs = (w > 0.0 && h > 0.0) ? sqrt(w * h * 16.0 / (double)n) : fmax(w, h) * 16.0 / (double)n;
    if (!(cs > 0.0) || !isfinite(cs)) cs = 1.0;
    cs = fmax(cs, fmax(w, h) / 2048.0);
    g.ox = c->pb[0]; g.oy = c->pb[2]; g.cs = cs; g.inv_cs = 1.0 / cs; g.tol = 0.0;
    g.gw = (int)fmin(floor(w / cs) + 1.0, 2049.0); g.gh = (int)fmin(floor(h / cs) + 1.0, 2049.0);
    if (g.gw < 1) g.gw = 1;
    if (g.gh < 1) g.gh = 1;
    const int n_cells = g.gw * g.gh;
    TRY(ensure_cells(c, n_cells));
    int* cell_of = nullptr;
    CU_TRY(cudaMalloc((void**)&cell_of, sizeof(int) * (size_t)n));
    int nb = (int)((n + 255) / 256);
    if (nb > c->sm_count * 16) nb = c->sm_count * 16;
    TESS_LAUNCH((k_pts_cell), nb, 256, s, g, c->pts_xy, n, cell_of, c->cell_cnt);
    const int nsb = (n_cells + TESS_SCAN_BLOCK - 1) / TESS_SCAN_BLOCK;
    TESS_LAUNCH((k_cell_blocksum), nsb, TESS_SCAN_THREADS, s, c->st, 0, c->cell_cnt, n_cells, c->block_sum);
    TESS_LAUNCH((k_cell_scan), nsb, TESS_SCAN_THREADS, s, c->st, 0, c->cell_cnt, c->cell_start, n_cells, c->block_sum);
    TESS_LAUNCH((k_pts_scatter), nb, 256, s, cell_of, c->cell_start, c->cell_cnt, c->pts_xy, c->pts_w, n, c->pts_perm,
                c->pts_xy_s, c->pts_w_s);
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaStreamSynchronize(s));
    cudaFree(cell_of);
    c->index_valid = 0;
    return TESS_OK;
}

extern "C" int tess_set_points(tess_ctx* c, long n, const double* xy, const double* w, void* stream) {
    TRY(bind(c));
    if (n <= 0 || !xy || !w) return tess_fail(TESS_ERR_ARG, "tess_set_points: need n > 0, xy and w");
    if (n > 2000000000L) return tess_fail(TESS_ERR_ARG, "tess_set_points: n too large");
    c->source = SRC_POINTS;
    c->index_valid = 0;
    c->img_mode = 0;
    c->n_pts = n; c->pts_xy = xy; c->pts_w = w;
    c->dens = c->sig = c->noise = nullptr;
    TRY(points_bbox(c, n, xy, c->pb, S(stream)));
    TRY(ensure_labels(c, n));
    TRY(sort_points(c, S(stream)));
    TRY(restart(c, S(stream)));
    return TESS_OK;
}

extern "C" int tess_set_generators(tess_ctx* c, long k, const double* node_xy, const double* scale, void* stream) {
    TRY(bind(c));
    if (k <= 0 || !node_xy) return tess_fail(TESS_ERR_ARG, "tess_set_generators: need k > 0 and node_xy");
    if (k > 1000000000L) return tess_fail(TESS_ERR_ARG, "tess_set_generators: k too large");
    cudaStream_t s = S(stream);
    cudaStream_t s_early = s;
    if (k > c->k_cap) {
        void** ps[] = {(void**)&c->node, (void**)&c->scale, (void**)&c->inv_s2, (void**)&c->cell_of,
                       (void**)&c->cell_items, (void**)&c->cell_xy, (void**)&c->cell_inv, (void**)&c->prev_count};
        for (void** p : ps) { if (*p) cudaFree(*p); *p = nullptr; }
        c->k_cap = 0;
        CU_TRY(cudaMalloc((void**)&c->node, sizeof(double) * 2 * k));
        CU_TRY(cudaMalloc((void**)&c->scale, sizeof(double) * k));
        CU_TRY(cudaMalloc((void**)&c->inv_s2, sizeof(double) * k));
        CU_TRY(cudaMalloc((void**)&c->cell_of, sizeof(int) * k));
        CU_TRY(cudaMalloc((void**)&c->cell_items, sizeof(int) * k));
        CU_TRY(cudaMalloc((void**)&c->cell_xy, sizeof(double2) * k));
        CU_TRY(cudaMalloc((void**)&c->cell_inv, sizeof(double) * k));
        CU_TRY(cudaMalloc((void**)&c->prev_count, sizeof(double) * k));
        c->k_cap = (int)k;
    }
    c->k = (int)k;
    if (c->cell_cnt && c->cells_cap > 0)   // (an update that raised NONFINITE leaves a partial histogram behind)
        CU_TRY(cudaMemsetAsync(c->cell_cnt, 0, sizeof(int) * (size_t)c->cells_cap, s_early));
    c->index_valid = 0;
    c->cell_shift = -1;
    c->regrid_tried = 0;
    c->gb_valid = 0;
    // WVT scale updates need a scale table: start from unit scales when none is given
    c->has_scale = (scale || c->wvt_update) ? 1 : 0;
    const int nb = (int)((k + 255) / 256);
    CU_TRY(cudaMemcpyAsync(c->node, node_xy, sizeof(double) * 2 * k, cudaMemcpyDeviceToDevice, s));
    if (scale) CU_TRY(cudaMemcpyAsync(c->scale, scale, sizeof(double) * k, cudaMemcpyDeviceToDevice, s));
    else if (c->has_scale) TESS_LAUNCH((k_fill), nb, 256, s, c->scale, k, 1.0);
    TRY(reset_iteration_state(c, s));
    TESS_LAUNCH((k_check_generators), nb, 256, s, c->st, c->node, c->has_scale ? c->scale : nullptr, c->inv_s2, (int)k);
    TRY(restart_tail(c, s));
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaMemcpyAsync(c->h_st, c->st, sizeof(TessState), cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaStreamSynchronize(s));
    if (c->h_st->flags & TESS_FLAG_NONFINITE)
        return tess_fail(TESS_ERR_NONFINITE, "tess_set_generators: a generator coordinate or scale is not finite");
    return TESS_OK;
}

// ------------------------------------------------------------------------------------ index
struct Geometry {
    TessGrid g;
    TessTiles T;   // image/grid queries only
    int n_cells;
};

// Grid for an H x W pixel window at (x0, y0): cells of 64 >> shift px over the window (united with
// the generator domain if one was given); generators outside are clamped into the border cells.
static Geometry window_geometry(const tess_ctx* c, double x0, double y0, long H, long W, int shift) {
    Geometry q;
    q.T.x0 = x0; q.T.y0 = y0; q.T.W = W; q.T.H = H;
    q.T.tiles_x = (int)((W + TESS_TS - 1) / TESS_TS);
    q.T.tiles_y = (int)((H + TESS_TS - 1) / TESS_TS);
    q.T.n_tiles = q.T.tiles_x * q.T.tiles_y;
    double gx0 = x0, gy0 = y0, gx1 = x0 + (double)W, gy1 = y0 + (double)H;
    const double cs = (double)TESS_TS / (double)(1 << shift);
    if (c->has_domain) {
        // extend by whole cells so that tiles stay aligned with cells
        if (c->dom[0] < gx0) gx0 -= cs * ceil((gx0 - c->dom[0]) / cs);
        if (c->dom[1] < gy0) gy0 -= cs * ceil((gy0 - c->dom[1]) / cs);
        if (c->dom[2] > gx1) gx1 = c->dom[2];
        if (c->dom[3] > gy1) gy1 = c->dom[3];
    }
    double fw = ceil((gx1 - gx0) / cs), fh = ceil((gy1 - gy0) / cs);
    if (fw * fh > 6.4e7) {   // cap the grid (16k x 4k cells); farther generators are clamped
        gx0 = x0; gy0 = y0; fw = (double)q.T.tiles_x; fh = (double)q.T.tiles_y;
    }
    q.g.ox = gx0; q.g.oy = gy0; q.g.cs = cs; q.g.inv_cs = 1.0 / cs;
    q.g.gw = (int)(fw < 1.0 ? 1.0 : fw); q.g.gh = (int)(fh < 1.0 ? 1.0 : fh);
    const double ext = fmax(fmax(fabs(gx0), fabs(gx1)), fmax(fabs(gy0), fabs(gy1))) + cs;
    q.g.tol = ext * 9.094947017729282e-13 /* 2^-40 */;
    q.n_cells = q.g.gw * q.g.gh;
    return q;
}

// Grid for a point query.  It covers the part of the plane where both the query points and the
// generators are (the intersection of the two bounding boxes; the generators' box if they do not
// overlap), with about 0.75 generators per cell on average (point sets are rarely uniform: the dense
// cells then stay short, empty ones cost two index loads per ring row).  Whatever lies outside --
// query points or generators -- is clamped into the border cells; the ring bound stays exact because
// clamping to a box is non-expansive.  Degenerate boxes (a single point, points on a line) get a
// one-cell-wide grid in the flat direction instead of absurdly thin cells.
static Geometry box_geometry(const tess_ctx* c, const double* pbox) {
    Geometry q;
    memset(&q, 0, sizeof(q));
    double b[4] = {pbox[0], pbox[1], pbox[2], pbox[3]};
    if (c->gb_valid) {
        const double ix0 = fmax(b[0], c->gb[0]), ix1 = fmin(b[1], c->gb[1]);
        const double iy0 = fmax(b[2], c->gb[2]), iy1 = fmin(b[3], c->gb[3]);
        if (ix0 <= ix1 && iy0 <= iy1) { b[0] = ix0; b[1] = ix1; b[2] = iy0; b[3] = iy1; }
        else { b[0] = c->gb[0]; b[1] = c->gb[1]; b[2] = c->gb[2]; b[3] = c->gb[3]; }
    }
    const double w = fmax(b[1] - b[0], 0.0), h = fmax(b[3] - b[2], 0.0);
    const double cells = fmax((double)c->k / 0.75, 1.0);
    double cs;
    if (w > 0.0 && h > 0.0) cs = sqrt(w * h / cells);
    else if (w > 0.0 || h > 0.0) cs = fmax(w, h) / fmin(cells, 4096.0);
    else cs = 1.0;
    if (!(cs > 0.0) || !isfinite(cs)) cs = 1.0;
    cs = fmax(cs, fmax(w, h) / 4096.0);
    const int gw = (int)fmin(floor(w / cs) + 1.0, 4097.0), gh = (int)fmin(floor(h / cs) + 1.0, 4097.0);
    q.g.ox = b[0]; q.g.oy = b[2]; q.g.cs = cs; q.g.inv_cs = 1.0 / cs;
    q.g.gw = gw < 1 ? 1 : gw; q.g.gh = gh < 1 ? 1 : gh;
    const double ext = fmax(fmax(fabs(b[0]), fabs(b[1])), fmax(fabs(b[2]), fabs(b[3]))) + cs;
    q.g.tol = ext * 9.094947017729282e-13;
    q.n_cells = q.g.gw * q.g.gh;
    return q;
}

// bounding box of the generator table (point-set mode only; once per tess_set_generators)
static int generator_bbox(tess_ctx* c, cudaStream_t s) {
    if (c->gb_valid) return TESS_OK;
    TRY(points_bbox(c, c->k, c->node, c->gb, s, false));
    c->gb_valid = 1;
    return TESS_OK;
}

static int ensure_cells(tess_ctx* c, int n_cells) {
    if (n_cells + 1 <= c->cells_cap) return TESS_OK;
    if (c->cell_cnt) cudaFree(c->cell_cnt);
    if (c->cell_start) cudaFree(c->cell_start);
    if (c->block_sum) cudaFree(c->block_sum);
    c->cell_cnt = c->cell_start = c->block_sum = nullptr;
    c->cells_cap = 0;
    CU_TRY(cudaMalloc((void**)&c->cell_cnt, sizeof(int) * (size_t)(n_cells + 1)));
    CU_TRY(cudaMemset(c->cell_cnt, 0, sizeof(int) * (size_t)(n_cells + 1)));   // scatter keeps it zero
    CU_TRY(cudaMalloc((void**)&c->cell_start, sizeof(int) * (size_t)(n_cells + 1)));
    CU_TRY(cudaMalloc((void**)&c->block_sum, sizeof(int) * (size_t)(n_cells / TESS_SCAN_BLOCK + 2)));
    c->cells_cap = n_cells + 1;
    return TESS_OK;
}

// Geometry of an image query, with the cell size decided once per (generator table, window):
// one tile (64 px) per cell unless some tile holds more than ~110 generators (bins below ~40 px).
// Such tiles overflow the list builder and their pixels run the per-pixel ring search, which wants a
// few generators per cell: the cell is halved until the fullest one holds <= 16 and -- for tables
// that are dense everywhere -- the mean occupancy is <= 4.  Cells stay power-of-two fractions of a
// tile, so tiles remain unions of cells.  Sparser tables keep the coarse grid: the list builder walks
// it faster (13-24 % of the whole iteration at 60-100 px per bin).  Costs one small kernel and one
// host read-back after tess_set_generators / tess_set_image / tess_set_domain, never per iteration.
static int image_geometry(tess_ctx* c, double x0, double y0, long H, long W, cudaStream_t s, Geometry* out) {
    const double key[4] = {x0, y0, (double)H, (double)W};
    if (c->cell_shift < 0 || memcmp(key, c->geo_key, sizeof(key)) != 0) {
        int shift = 0;
        if (c->k > 80) {
            const Geometry q0 = window_geometry(c, x0, y0, H, W, 0);
            TRY(ensure_cells(c, q0.n_cells));
            CU_TRY(cudaMemsetAsync(&c->st->tile_counter, 0, sizeof(int), s));
            int nb = (c->k + 255) / 256;
            if (nb > c->sm_count * 16) nb = c->sm_count * 16;
            TESS_LAUNCH((k_cell_occupancy_max), nb, 256, s, c->st, q0.g, c->node, c->k, c->cell_cnt);
            CU_TRY(cudaMemsetAsync(c->cell_cnt, 0, sizeof(int) * (size_t)(q0.n_cells + 1), s));
            CU_TRY(cudaMemcpyAsync(c->h_st, c->st, sizeof(TessState), cudaMemcpyDeviceToHost, s));
            CU_TRY(cudaStreamSynchronize(s));
            double occ_max = (double)c->h_st->tile_counter;
            double occ_mean = (double)c->k / fmax((double)q0.n_cells, 1.0);
            double cells = (double)q0.n_cells;
            if (occ_max > 110.0)
                while ((occ_max > 16.0 || occ_mean > 4.0) && shift < 5 && cells * 4.0 <= 3.2e7) {
                    ++shift; occ_max *= 0.25; occ_mean *= 0.25; cells *= 4.0;
                }
        }
        c->cell_shift = shift;
        memcpy(c->geo_key, key, sizeof(key));
    }
    *out = window_geometry(c, x0, y0, H, W, c->cell_shift);
    return TESS_OK;
}

static int ensure_lists(tess_ctx* c, const TessTiles& T) {
    const int nqx = 2 * T.tiles_x, nqy = 2 * T.tiles_y;
    const long nq = (long)nqx * nqy;
    if (nq > c->quads_cap) {
        if (c->ql.q_off) cudaFree(c->ql.q_off);
        if (c->ql.q_cnt) cudaFree(c->ql.q_cnt);
        if (c->ql.ring_q) cudaFree(c->ql.ring_q);
        if (c->tile_ring) cudaFree(c->tile_ring);
        c->ql.q_off = c->ql.q_cnt = c->ql.ring_q = nullptr;
        c->tile_ring = nullptr;
        c->quads_cap = 0;
        CU_TRY(cudaMalloc((void**)&c->tile_ring, sizeof(int) * (size_t)nq));      // >= n_tiles
        CU_TRY(cudaMemset(c->tile_ring, 0, sizeof(int) * (size_t)nq));
        CU_TRY(cudaMalloc((void**)&c->ql.q_off, sizeof(int) * (size_t)nq));
        CU_TRY(cudaMalloc((void**)&c->ql.q_cnt, sizeof(int) * (size_t)nq));
        CU_TRY(cudaMalloc((void**)&c->ql.ring_q, sizeof(int) * (size_t)nq));
        c->quads_cap = nq;
    }
    // arena: a generator is listed by every quadrant it can win in -- its own bin plus a margin
    double want = 32.0 * (double)nq + 48.0 * (double)c->k + 65536.0;
    if (want > 2.0e9) want = 2.0e9;
    if ((unsigned)want > c->ql.capacity || (c->has_scale && !c->arena_has_inv)) {
        if (c->ql.xy) cudaFree(c->ql.xy);
        if (c->ql.gid) cudaFree(c->ql.gid);
        if (c->ql.inv) cudaFree(c->ql.inv);
        c->ql.xy = nullptr; c->ql.gid = nullptr; c->ql.inv = nullptr;
        c->ql.capacity = 0;
        c->arena_has_inv = 0;
        if ((double)c->ql.capacity > want) want = (double)c->ql.capacity;
        CU_TRY(cudaMalloc((void**)&c->ql.xy, sizeof(double2) * (size_t)want));
        CU_TRY(cudaMalloc((void**)&c->ql.gid, sizeof(int) * (size_t)want));
        if (c->has_scale) {
            CU_TRY(cudaMalloc((void**)&c->ql.inv, sizeof(double) * (size_t)want));
            c->arena_has_inv = 1;
        }
        c->ql.capacity = (unsigned)want;
    }
    c->ql.nqx = nqx;
    c->ql.nqy = nqy;
    return TESS_OK;
}

static int ensure_moments(tess_ctx* c) {
    const long need = (long)c->k * 6 + TESS_TAIL;
    if (need <= c->mom_cap) return TESS_OK;
    if (c->mom_external)
        return tess_fail(TESS_ERR_ARG, "the moment buffer given to tess_set_moment_buffer holds %ld doubles, %ld needed", c->mom_cap, need);
    if (c->mom) cudaFree(c->mom);
    c->mom = nullptr;
    c->mom_cap = 0;
    CU_TRY(cudaMalloc((void**)&c->mom, sizeof(double) * (size_t)need));
    CU_TRY(cudaMemset(c->mom, 0, sizeof(double) * (size_t)need));
    c->mom_cap = need;
    return TESS_OK;
}

// counting sort of the generators into q.g; `mom`/n_mom: moment vector to clear (or null/0)
static int build_index(tess_ctx* c, const Geometry& q, int honor, double* mom, long n_mom, cudaStream_t s) {
    TRY(ensure_cells(c, q.n_cells));
    c->index_valid = 0;      // whatever grid the iteration had built is overwritten
    const int k = c->k;
    int nb = (int)(((long)k + 255) / 256);
    int nbz = (int)((n_mom + 255) / 256);
    int grid = nb > nbz ? nb : nbz;
    if (grid > c->sm_count * 16) grid = c->sm_count * 16;
    TESS_LAUNCH((k_cell_count), grid, 256, s, c->st, honor, q.g, c->node, k, c->cell_cnt, c->cell_of, mom, n_mom);
    const int nsb = (q.n_cells + TESS_SCAN_BLOCK - 1) / TESS_SCAN_BLOCK;
    TESS_LAUNCH((k_cell_blocksum), nsb, TESS_SCAN_THREADS, s, c->st, honor, c->cell_cnt, q.n_cells, c->block_sum);
    TESS_LAUNCH((k_cell_scan), nsb, TESS_SCAN_THREADS, s, c->st, honor, c->cell_cnt, c->cell_start, q.n_cells, c->block_sum);
    TESS_LAUNCH((k_cell_scatter), nb, 256, s, c->st, honor, c->cell_of, k, c->cell_start, c->cell_cnt, c->node,
                                      c->has_scale ? c->inv_s2 : nullptr, c->cell_items, c->cell_xy, c->cell_inv);
    CU_TRY(cudaGetLastError());
    return TESS_OK;
}

static int build_lists(tess_ctx* c, const Geometry& q, int honor, int mode, cudaStream_t s, double* mom_clear = nullptr,
                       long n_clear = 0) {
    TRY(ensure_lists(c, q.T));
    const int qcap = TESS_QCAP(mode);          // longest list the image kernel of this mode stages
    int grid = (q.T.n_tiles + TESS_LIST_WARPS - 1) / TESS_LIST_WARPS;
    if (grid > c->sm_count * 32) grid = c->sm_count * 32;
    if (c->has_scale)
        TESS_LAUNCH((k_build_qlists<true>), grid, TESS_LIST_WARPS * 32, s, c->st, honor, q.T, q.g, c->cell_start,
                    c->cell_items, c->cell_xy, c->cell_inv, c->ql, c->tile_ring, c->cell_shift > 0 ? 1 : 0, qcap, mom_clear, n_clear);
    else
        TESS_LAUNCH((k_build_qlists<false>), grid, TESS_LIST_WARPS * 32, s, c->st, honor, q.T, q.g, c->cell_start,
                    c->cell_items, c->cell_xy, c->cell_inv, c->ql, c->tile_ring, c->cell_shift > 0 ? 1 : 0, qcap, mom_clear, n_clear);
    CU_TRY(cudaGetLastError());
    return TESS_OK;
}

// persistent grid: exactly the CTAs that are resident at once (the kernel strides over quadrants)
template <class K>
static int resident_ctas(tess_ctx* c, K kern) {
#ifndef TESS_EMU
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, TESS_MAIN_THREADS, 0) != cudaSuccess || per_sm < 1)
        per_sm = 4;
    return c->sm_count * per_sm;
#else
    (void)kern;
    return c->sm_count * 2;
#endif
}

// ------------------------------------------------------------------------------------ phased O(k) steps
static bool same_grid(const TessGrid& a, const TessGrid& b) {
    return a.ox == b.ox && a.oy == b.oy && a.cs == b.cs && a.gw == b.gw && a.gh == b.gh;
}

static int n_moments(const tess_ctx* c);
static unsigned long long* range_slots(tess_ctx* c);

static long staged_off(tess_ctx* c);
static void phase_args(tess_ctx* c, const Geometry& q, TessPhaseArgs* a) {
    memset(a, 0, sizeof(*a));
    a->st = c->st;
    a->mom = c->mom; a->M = n_moments(c); a->k = c->k;
    a->prev_count = c->prev_count; a->node = c->node;
    a->scale = c->has_scale ? c->scale : nullptr;
    a->inv_s2 = c->inv_s2;
    a->inv_in = c->has_scale ? c->inv_s2 : nullptr;
    a->wvt_update = c->wvt_update;
    a->delta_part = c->delta_part;
    a->range = range_slots(c);
    a->staged = (c->staged_valid && staged_off(c) >= 0) ? c->mom + staged_off(c) : nullptr;
    a->mask_mode = (c->source == SRC_IMAGE) ? c->img_mode : 0;
    a->lab0 = c->lab[0]; a->lab1 = c->lab[1];
    a->n = (c->source == SRC_IMAGE) ? c->H * c->W : c->n_pts;
    a->dens = c->dens; a->sig = c->sig; a->noise = c->noise;
    a->uniform_weight = c->uniform_weight;
    a->g = q.g; a->n_cells = q.n_cells;
    a->cell_cnt = c->cell_cnt; a->cell_of = c->cell_of; a->cell_start = c->cell_start; a->block_sum = c->block_sum;
    a->cell_items = c->cell_items; a->cell_xy = c->cell_xy; a->cell_inv = c->cell_inv;
}

// phases lo..hi of k_iteration_phases (tess_phases.cuh): one cooperative launch, or one launch per
// phase where grid barriers do not exist (the SIMT emulator; a device without cooperative launch)
static int run_phases(tess_ctx* c, const Geometry& q, int lo, int hi, int with_update, cudaStream_t s) {
    TRY(ensure_cells(c, q.n_cells));
    TessPhaseArgs a;
    phase_args(c, q, &a);
#ifndef TESS_EMU
    if (c->phase_grid == 0) {
        int per_sm = 0, coop = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_iteration_phases, 256, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
        int cap = 4;
        if (const char* e = getenv("TESS_PHASE_CTAS")) cap = atoi(e) > 0 ? atoi(e) : cap;      // tuning aid
        if (per_sm > cap) per_sm = cap;
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, c->device);
        cudaGetLastError();
        c->phase_grid = c->sm_count * per_sm;
        c->coop_ok = coop;
    }
    if (c->coop_ok && hi > lo) {
        void* args[] = {(void*)&a, (void*)&lo, (void*)&hi, (void*)&with_update};
        ++g_tess_launches;
        CU_TRY(cudaLaunchCooperativeKernel((const void*)k_iteration_phases, dim3(c->phase_grid), dim3(256), args, 0, s));
        return TESS_OK;
    }
    const int grid = c->phase_grid;
#else
    const int grid = 2;          // (every emulated thread is an OS thread: keep the grid small)
#endif
    for (int ph = lo; ph <= hi; ++ph) {
        if (ph == TESS_PH_SCAN && q.n_cells <= TESS_SINGLE_SCAN_MAX) continue;
        TESS_LAUNCH((k_iteration_phases), grid, 256, s, a, ph, ph, with_update);
    }
    CU_TRY(cudaGetLastError());
    return TESS_OK;
}

// ---- TMA descriptor of an [H][W] fp64 map with a 32x32-px box (the quadrant tile).  The encoder is a
// driver entry point, fetched through the runtime so that the library does not link libcuda.
#ifndef TESS_EMU
#include <cuda.h>
#include <cudaTypedefs.h>
static bool make_tile_map(const double* base, long H, long W, TessTileMap* out) {
    static PFN_cuTensorMapEncodeTiled_v12000 enc = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) == cudaSuccess &&
            qr == cudaDriverEntryPointSuccess)
            enc = (PFN_cuTensorMapEncodeTiled_v12000)fn;
        cudaGetLastError();
    }
    if (!enc || !base || ((uintptr_t)base & 15u) || (W & 1) || W < 32 || H < 32 || W > (1L << 31) || H > (1L << 31)) return false;
    static_assert(sizeof(CUtensorMap) == sizeof(TessTileMap), "tensor map size");
    CUtensorMap m;
    const cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)H};
    const cuuint64_t strides[1] = {(cuuint64_t)W * 8u};
    const cuuint32_t box[2] = {32u, 32u};
    const cuuint32_t estr[2] = {1u, 1u};
    const CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)base, dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return false;
    memcpy(out, &m, sizeof(m));
    return true;
}
#else
static bool make_tile_map(const double*, long, long, TessTileMap*) { return false; }
#endif

// descriptors of the current pixel maps (re-encoded only when a pointer or the shape changed)
static void ensure_tile_maps(tess_ctx* c, int mode, long H, long W) {
    const double* src[2] = {mode == 2 ? c->sig : c->dens, mode == 2 ? c->noise : nullptr};
    if (c->tmap_ok && c->tmap_src[0] == src[0] && c->tmap_src[1] == src[1] && c->tmap_hw[0] == H && c->tmap_hw[1] == W) return;
    c->tmap_ok = 0;
    c->tmap_src[0] = src[0]; c->tmap_src[1] = src[1]; c->tmap_hw[0] = H; c->tmap_hw[1] = W;
    if (mode == 0 || !make_tile_map(src[0], H, W, &c->tmap[0])) return;
    if (mode == 2 && !make_tile_map(src[1], H, W, &c->tmap[1])) return;
    if (mode != 2) c->tmap[1] = c->tmap[0];
    c->tmap_ok = 1;
}

template <int MODE>
static void launch_image(tess_ctx* c, const TessImgArgs& a, bool vec, long want, cudaStream_t s) {
#define TESS_GO(KERN)                                                   \
    do {                                                                \
        long g_ = resident_ctas(c, KERN);                               \
        /* one warp per quadrant, but the tail phase has 32 / TESS_RING_ROWS strips per queued quadrant */ \
        if (g_ > want * (32 / TESS_RING_ROWS)) g_ = want * (32 / TESS_RING_ROWS); \
        if (g_ < 1) g_ = 1;                                             \
        TESS_LAUNCH((KERN), (int)g_, TESS_MAIN_THREADS, s, a, vec ? 1 : 0, use_tma, c->tmap[0], c->tmap[1]); \
    } while (0)
    // interior quadrants (aligned 128-bit accesses), then the ragged edge / unaligned remainder
    const bool has_edge = !vec || (a.W % 32 != 0) || (a.H % 32 != 0);
    int use_tma = 0;
    static const bool no_tma = getenv("TESS_NO_TMA") != nullptr;     // measurement aid: A/B the prefetch
    if (vec && MODE != 0 && !no_tma) {
        ensure_tile_maps(c, MODE, a.H, a.W);
        use_tma = c->tmap_ok;
    }
    if (c->has_scale) {
        if (vec) TESS_GO((k_image_main<MODE, true, false>));
        if (has_edge) TESS_GO((k_image_main<MODE, true, true>));
    } else {
        if (vec) TESS_GO((k_image_main<MODE, false, false>));
        if (has_edge) TESS_GO((k_image_main<MODE, false, true>));
    }
#undef TESS_GO
}

static bool aligned16(const void* p) { return ((uintptr_t)p & 15u) == 0; }

static int run_image(tess_ctx* c, const Geometry& q, int mode, int honor, int32_t* labels_ext, cudaStream_t s) {
    TessImgArgs a;
    memset(&a, 0, sizeof(a));
    a.dens = c->dens; a.sig = c->sig; a.noise = c->noise;
    a.labels_ext = labels_ext; a.lab0 = c->lab[0]; a.lab1 = c->lab[1];
    a.mom = c->mom;
    a.L = c->ql;
    a.g = q.g; a.cell_start = c->cell_start; a.cell_items = c->cell_items; a.cell_xy = c->cell_xy; a.cell_inv = c->cell_inv;
    a.st = c->st;
    a.H = q.T.H; a.W = q.T.W; a.x0 = q.T.x0; a.y0 = q.T.y0;
    a.uniform_weight = c->uniform_weight;
    a.honor = honor;
    bool vec = (q.T.W % 2 == 0);
    if (labels_ext) vec = vec && (((uintptr_t)labels_ext & 7u) == 0);
    if (mode == 1) vec = vec && aligned16(c->dens);
    if (mode == 2) vec = vec && aligned16(c->sig) && aligned16(c->noise);
    // one warp per quadrant, strided: enough CTAs to fill every SM at the compiled occupancy
    const long nq = (long)c->ql.nqx * c->ql.nqy;
    const int nw = TESS_MAIN_THREADS / 32;
    const long want = (nq + nw - 1) / nw;
    if (mode == 0) launch_image<0>(c, a, vec, want, s);
    else if (mode == 1) launch_image<1>(c, a, vec, want, s);
    else launch_image<2>(c, a, vec, want, s);
    CU_TRY(cudaGetLastError());
    return TESS_OK;
}

static int run_points(tess_ctx* c, const Geometry& q, int mode, int honor, long n, const double* xy, const double* w,
                      const int* perm, int32_t* labels_ext, cudaStream_t s) {
    const int nb = (int)((n + 255) / 256);
    if (mode == 1) {
        if (c->has_scale)
            TESS_LAUNCH((k_points_main<1, true>), nb, 256, s, c->st, honor, q.g, c->cell_start, c->cell_items, c->cell_xy,
                                                     c->cell_inv, n, xy, w, perm, labels_ext, c->lab[0], c->lab[1], c->mom);
        else
            TESS_LAUNCH((k_points_main<1, false>), nb, 256, s, c->st, honor, q.g, c->cell_start, c->cell_items, c->cell_xy,
                                                      c->cell_inv, n, xy, w, perm, labels_ext, c->lab[0], c->lab[1], c->mom);
    } else {
        if (c->has_scale)
            TESS_LAUNCH((k_points_main<0, true>), nb, 256, s, c->st, honor, q.g, c->cell_start, c->cell_items, c->cell_xy,
                                                     c->cell_inv, n, xy, w, perm, labels_ext, c->lab[0], c->lab[1], c->mom);
        else
            TESS_LAUNCH((k_points_main<0, false>), nb, 256, s, c->st, honor, q.g, c->cell_start, c->cell_items, c->cell_xy,
                                                      c->cell_inv, n, xy, w, perm, labels_ext, c->lab[0], c->lab[1], c->mom);
    }
    CU_TRY(cudaGetLastError());
    return TESS_OK;
}

static int fetch_state(tess_ctx* c, cudaStream_t s);

// event bracket around the dominant kernel (TESS_OPT_KERNEL_TIMING)
static void timing_mark(tess_ctx* c, cudaStream_t s, int end) {
#ifndef TESS_EMU
    if (!c->timing) return;
    if (!end && c->n_ev_used + 2 > 2 * 512) return;
    if (end && (c->n_ev_used & 1) == 0) return;
    while (c->n_ev_alloc <= c->n_ev_used) {
        if (cudaEventCreate(&c->ev[c->n_ev_alloc]) != cudaSuccess) return;
        ++c->n_ev_alloc;
    }
    cudaEventRecord(c->ev[c->n_ev_used++], s);
#else
    (void)c; (void)s; (void)end;
#endif
}

extern "C" int tess_get_kernel_timing(tess_ctx* c, double* ms_total, long* launches) {
    TRY(bind(c));
    double tot = 0.0;
    long n = 0;
#ifndef TESS_EMU
    CU_TRY(cudaDeviceSynchronize());
    for (int i = 0; i + 1 < c->n_ev_used; i += 2) {
        float ms = 0.f;
        CU_TRY(cudaEventElapsedTime(&ms, c->ev[i], c->ev[i + 1]));
        tot += ms;
        ++n;
    }
    c->n_ev_used = 0;
#endif
    if (ms_total) *ms_total = tot;
    if (launches) *launches = n;
    return TESS_OK;
}

extern "C" long tess_launch_count(void) { return g_tess_launches; }

// list diagnostics of the most recent list build, all from one snapshot of the state block:
// [0] quadrants of tiles that overflowed the builder (per-pixel ring search), [1] arena records used,
// [2] arena capacity, [3] quadrants whose list is longer than the fast path stages (scanned from L2)
extern "C" int tess_get_list_stats(tess_ctx* c, long* out4) {
    TRY(bind(c));
    if (!out4) return tess_fail(TESS_ERR_ARG, "tess_get_list_stats: out is null");
    TRY(fetch_state(c, nullptr));
    out4[0] = c->h_st->n_overflow;
    out4[1] = c->h_st->list_cursor;
    out4[2] = c->ql.capacity;
    out4[3] = c->h_st->n_brute - c->h_st->n_overflow;
    return TESS_OK;
}

static int check_ready(tess_ctx* c, const char* who) {
    TRY(bind(c));
    if (c->source == SRC_NONE) return tess_fail(TESS_ERR_STATE, "%s: call tess_set_image or tess_set_points first", who);
    if (c->k <= 0) return tess_fail(TESS_ERR_STATE, "%s: call tess_set_generators first", who);
    return TESS_OK;
}

// ------------------------------------------------------------------------------------ iteration
// geometry of the generator grid for the current source (image window or point set)
static int source_geometry(tess_ctx* c, cudaStream_t s, Geometry* q) {
    if (c->source == SRC_IMAGE) return image_geometry(c, c->x0, c->y0, c->H, c->W, s, q);
    TRY(generator_bbox(c, s));
    *q = box_geometry(c, c->pb);
    return TESS_OK;
}

// accumulate without its closing "did anything change" phases (the caller appends them)
static int accumulate_core(tess_ctx* c, const Geometry& q, cudaStream_t s) {
    const long n_rows = (long)c->k * n_moments(c);
    c->staged_valid = 0;
    // the generator grid: left behind by the previous update (phases 3-6 of its launch), else built now
    if (!(c->index_valid && same_grid(q.g, c->index_geo))) TRY(run_phases(c, q, TESS_PH_UPDATE, TESS_PH_SCATTER, 0, s));
    c->index_valid = 0;
    if (c->source == SRC_IMAGE) {
        TRY(build_lists(c, q, 1, c->img_mode, s, c->mom, n_rows));
        timing_mark(c, s, 0);
        TRY(run_image(c, q, c->img_mode, 1, nullptr, s));
        timing_mark(c, s, 1);
    } else {
        int grid = (int)((n_rows + 255) / 256);
        if (grid > c->sm_count * 8) grid = c->sm_count * 8;
        TESS_LAUNCH((k_clear_moments), grid, 256, s, c->st, 1, c->mom, n_rows);
        timing_mark(c, s, 0);
        TRY(run_points(c, q, 1, 1, c->n_pts, reinterpret_cast<const double*>(c->pts_xy_s), c->pts_w_s, c->pts_perm, nullptr, s));
        timing_mark(c, s, 1);
    }
    CU_TRY(cudaGetLastError());
    return TESS_OK;
}

extern "C" int tess_lloyd_accumulate(tess_ctx* c, void* stream) {
    TRY(check_ready(c, "tess_lloyd_accumulate"));
    cudaStream_t s = S(stream);
    TRY(ensure_moments(c));
    Geometry q;
    TRY(source_geometry(c, s, &q));
    TRY(accumulate_core(c, q, s));
    TRY(run_phases(c, q, TESS_PH_PREFILTER, TESS_PH_COMPARE, 1, s));
    return TESS_OK;
}

extern "C" int tess_moments_ptr(tess_ctx* c, double** dev_ptr, long* n_doubles, int* n_mom) {
    TRY(check_ready(c, "tess_moments_ptr"));
    TRY(ensure_moments(c));
    const int M = n_moments(c);
    if (dev_ptr) *dev_ptr = c->mom;
    if (n_doubles) *n_doubles = (long)c->k * M + TESS_TAIL;
    if (n_mom) *n_mom = M;
    return TESS_OK;
}

extern "C" int tess_set_moment_buffer(tess_ctx* c, double* buf, long n_doubles) {
    TRY(bind(c));
    if (!buf) {                       // back to a library-owned vector
        if (c->mom_external) { c->mom = nullptr; c->mom_cap = 0; c->mom_external = 0; }
        c->sparse_mc = nullptr;
        return TESS_OK;
    }
    if (n_doubles < TESS_TAIL + 6 || ((uintptr_t)buf & 15u))
        return tess_fail(TESS_ERR_ARG, "tess_set_moment_buffer: need a 16-byte aligned buffer of at least 6 k + %d doubles", TESS_TAIL);
    if (c->mom && !c->mom_external) cudaFree(c->mom);
    c->mom = buf;
    c->mom_cap = n_doubles;
    c->mom_external = 1;
    // the slots behind the vector that the sparse exchange uses (barrier flags, published range)
    if (n_doubles >= TESS_SPARSE_EXTRA + TESS_TAIL + 6)
        TESS_LAUNCH((k_reset_sparse_slots), 1, 64, nullptr, reinterpret_cast<unsigned long long*>(buf + n_doubles - TESS_SPARSE_EXTRA));
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaStreamSynchronize(nullptr));
    return TESS_OK;
}

// offset (doubles) of the node area of the caller's buffer -- 2 k doubles right behind the vector -- or -1
// when owner-computed updates do not apply (option off, 6-moment rows or WVT scale updates: those need
// the totals everywhere; no room in the buffer)
static long staged_off(tess_ctx* c) {
    if (!c->owner_update || !c->mom_external || !c->mom || n_moments(c) != 4 || c->wvt_update) return -1;
    const long off = (long)c->k * 4 + TESS_TAIL;          // even
    if (c->mom_cap < off + 2L * c->k + TESS_SPARSE_EXTRA) return -1;
    return off;
}

// slots of the published row range, or null when the caller's buffer has no room for them
static unsigned long long* range_slots(tess_ctx* c) {
    if (!c->mom_external || !c->mom) return nullptr;
    const long used = (long)c->k * n_moments(c) + TESS_TAIL;
    if (c->mom_cap < used + TESS_SPARSE_EXTRA || (c->mom_cap & 1)) return nullptr;
    return reinterpret_cast<unsigned long long*>(c->mom + c->mom_cap - TESS_RANGE_SLOTS);
}

extern "C" int tess_peer_publish_range(tess_ctx* c, void* stream) {
    TRY(check_ready(c, "tess_peer_publish_range"));
    unsigned long long* r = range_slots(c);
    if (!r) return tess_fail(TESS_ERR_STATE, "tess_peer_publish_range: needs a moment buffer from tess_set_moment_buffer with "
                                             "%d spare doubles behind the vector", TESS_RANGE_SLOTS);
    int nb = (c->k + 255) / 256;
    if (nb > c->sm_count * 8) nb = c->sm_count * 8;
    TESS_LAUNCH((k_touched_range), nb, 256, S(stream), c->st, c->mom, n_moments(c), c->k, r);
    CU_TRY(cudaGetLastError());
    return TESS_OK;
}

static int peer_allreduce(tess_ctx* c, int world, int rank, double* const* peer_ptrs, double* multicast_ptr,
                          long n_doubles, int sparse, void* stream);

extern "C" int tess_peer_allreduce(tess_ctx* c, int world, int rank, double* const* peer_ptrs, double* multicast_ptr,
                                   long n_doubles, void* stream) {
    return peer_allreduce(c, world, rank, peer_ptrs, multicast_ptr, n_doubles, 0, stream);
}

extern "C" int tess_peer_set_multicast(tess_ctx* c, double* multicast_ptr) {
    TRY(bind(c));
    c->sparse_mc = multicast_ptr;
    return TESS_OK;
}

extern "C" int tess_peer_allreduce_sparse(tess_ctx* c, int world, int rank, double* const* peer_ptrs, long n_doubles,
                                          void* stream) {
    return peer_allreduce(c, world, rank, peer_ptrs, nullptr, n_doubles, 1, stream);
}

static int peer_allreduce(tess_ctx* c, int world, int rank, double* const* peer_ptrs, double* multicast_ptr,
                          long n_doubles, int sparse, void* stream) {
    TRY(bind(c));
#ifdef TESS_EMU
    (void)world; (void)rank; (void)peer_ptrs; (void)multicast_ptr; (void)n_doubles; (void)stream; (void)sparse;
    return tess_fail(TESS_ERR_STATE, "tess_peer_allreduce: no peer memory in the emulator build");
#else
    if (world < 1 || world > TESS_PEER_MAX || rank < 0 || rank >= world || n_doubles <= 0 || (!peer_ptrs && !multicast_ptr))
        return tess_fail(TESS_ERR_ARG, "tess_peer_allreduce: bad arguments (world <= %d)", TESS_PEER_MAX);
    // rank r owns [lo, hi); slices start on even elements (16-byte accesses), the vector is even
    if (n_doubles & 1) return tess_fail(TESS_ERR_ARG, "tess_peer_allreduce: n_doubles must be even");
    long per = (n_doubles + world - 1) / world;
    per += per & 1;
    long lo = per * rank, hi = (lo + per < n_doubles) ? lo + per : n_doubles;
    const long node_off = sparse ? staged_off(c) : -1;
    if (node_off >= 0) {
        // owner-computed update: slices start on rows (4 doubles), the last rank also owns the tail
        const long rows_per = ((long)c->k + world - 1) / world;
        const long r0 = (rows_per * rank < c->k) ? rows_per * rank : c->k;
        const long r1 = (rows_per * (rank + 1) < c->k) ? rows_per * (rank + 1) : c->k;
        lo = r0 * 4;
        hi = (rank == world - 1) ? n_doubles : r1 * 4;
    }
    if (hi <= lo && !sparse) return TESS_OK;     // (the sparse kernel carries the cross-rank barriers: always launched)
    long nb = (hi - lo + 255) / 256;
    if (nb > c->sm_count * 8) nb = c->sm_count * 8;
    if (multicast_ptr) {
        nb = (hi - lo + 4 * 256 - 1) / (4 * 256);           // four elements per thread and trip
        if (nb > c->sm_count * 8) nb = c->sm_count * 8;
        TESS_LAUNCH((k_peer_sum_multimem), (int)nb, 256, S(stream), multicast_ptr, lo, hi);
    } else {
        TessPeers P;
        memset(&P, 0, sizeof(P));
        P.world = world; P.rank = rank;
        for (int r = 0; r < world; ++r) {
            if (!peer_ptrs[r]) return tess_fail(TESS_ERR_ARG, "tess_peer_allreduce: null peer pointer for rank %d", r);
            P.ptr[r] = peer_ptrs[r];
        }
        nb = ((hi - lo) / 2 + 255) / 256;
        if (nb > c->sm_count * 8) nb = c->sm_count * 8;
        if (sparse) {
            unsigned long long* rs = range_slots(c);
            if (!rs) return tess_fail(TESS_ERR_STATE, "tess_peer_allreduce_sparse: no published range (tess_peer_publish_range)");
            const int M = n_moments(c);
            const long range_off = (long)(reinterpret_cast<double*>(rs) - c->mom);
            static const int peer_ctas = getenv("TESS_PEER_CTAS") ? atoi(getenv("TESS_PEER_CTAS")) : 2;   // tuning aid
            nb = (hi - lo + TESS_PEER_CHUNK - 1) / TESS_PEER_CHUNK;          // one 16 KB chunk per CTA and trip
            if (nb > c->sm_count * peer_ctas) nb = c->sm_count * peer_ctas;
            if (nb < 1) nb = 1;
            c->staged_valid = node_off >= 0 ? 1 : 0;
            if (c->sparse_mc && node_off < 0) {              // register path: no 16 KB chunks, a pair of totals per thread and trip
                nb = ((hi - lo) / 2 + 255) / 256;
                if (nb > c->sm_count * 4) nb = c->sm_count * 4;
                if (nb < 1) nb = 1;
            }
            TESS_LAUNCH((k_peer_sum_p2p_sparse), (int)nb, 256, S(stream), P, node_off >= 0 ? nullptr : c->sparse_mc, lo, hi, M,
                        (long)c->k * M, range_off - TESS_FLAG_SLOTS, range_off, node_off, (unsigned long long)(++c->exchange_epoch),
                        reinterpret_cast<unsigned*>(c->scratch + 6));
        } else {
            TESS_LAUNCH((k_peer_sum_p2p), (int)nb, 256, S(stream), P, lo, hi);
        }
    }
    CU_TRY(cudaGetLastError());
    return TESS_OK;
#endif
}

// Node update, delta and latch, and -- in the same launch -- the generator grid of the new nodes for the
// next accumulate (wasted only after the very last iteration: ~10 us).
extern "C" int tess_lloyd_update(tess_ctx* c, void* stream) {
    TRY(check_ready(c, "tess_lloyd_update"));
    if (!c->mom) return tess_fail(TESS_ERR_STATE, "tess_lloyd_update: no accumulate has run");
    cudaStream_t s = S(stream);
    Geometry q;
    TRY(source_geometry(c, s, &q));
    TRY(run_phases(c, q, TESS_PH_UPDATE, TESS_PH_SCATTER, 1, s));
    c->staged_valid = 0;
    c->index_valid = 1;
    c->index_geo = q.g;
    return TESS_OK;
}

// One iteration = list builder + image (or point) kernel + ONE phased launch (change tests, update,
// latch, next grid).
extern "C" int tess_lloyd_step(tess_ctx* c, long n_steps, void* stream) {
    TRY(check_ready(c, "tess_lloyd_step"));
    cudaStream_t s = S(stream);
    for (long i = 0; i < n_steps; ++i) {
        TRY(ensure_moments(c));
        Geometry q;
        TRY(source_geometry(c, s, &q));
        TRY(accumulate_core(c, q, s));
        TRY(run_phases(c, q, TESS_PH_PREFILTER, TESS_PH_SCATTER, 1, s));
        c->index_valid = 1;
        c->index_geo = q.g;
    }
    return TESS_OK;
}

static int fetch_state(tess_ctx* c, cudaStream_t s) {
    CU_TRY(cudaMemcpyAsync(c->h_st, c->st, sizeof(TessState), cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaStreamSynchronize(s));
    // generators that drift into a dense clump overflow tiles the coarse grid was not refined for:
    // the host is synchronised here anyway, so let the next accumulate decide the cell size again
    // (true overflows only -- a merely long list does not want a finer grid -- and once per table)
    if (c->source == SRC_IMAGE && c->cell_shift == 0 && c->h_st->n_overflow > 0 && !c->regrid_tried) {
        c->cell_shift = -1;
        c->regrid_tried = 1;
    }
    return TESS_OK;
}

extern "C" int tess_lloyd_run(tess_ctx* c, long max_iters, int* converged, long* n_iters, void* stream) {
    TRY(check_ready(c, "tess_lloyd_run"));
    cudaStream_t s = S(stream);
    // lloyd.pyx:85-90: at most max_iters + 2 loop bodies (and at least one)
    long budget = max_iters + 2;
    if (budget < 1) budget = 1;
    TRY(fetch_state(c, s));
    const long long start = c->h_st->iters_done;
    long done = 0;
    int conv = c->h_st->converged;
    while (!conv && done < budget) {
        long chunk = budget - done;
        if (chunk > 4) chunk = 4;
        TRY(tess_lloyd_step(c, chunk, stream));
        TRY(fetch_state(c, s));
        done = (long)(c->h_st->iters_done - start);
        conv = c->h_st->converged;
        if (c->h_st->flags & TESS_FLAG_NONFINITE) {
            // the reference's next cKDTree build raises ValueError on NaN nodes (lloyd.pyx:54);
            // if the budget ended on this very iteration it returns them with converged = False
            if (converged) *converged = 0;
            if (n_iters) *n_iters = done;
            if (done < budget)
                return tess_fail(TESS_ERR_NONFINITE,
                                 "tess_lloyd_run: a bin with pixels has zero total weight: generator became NaN/Inf "
                                 "after iteration %ld", done);
            return TESS_OK;
        }
    }
    if (converged) *converged = conv;
    if (n_iters) *n_iters = done;
    return TESS_OK;
}

// ------------------------------------------------------------------------------------ results
extern "C" int tess_get_labels(tess_ctx* c, int32_t* out, void* stream) {
    TRY(check_ready(c, "tess_get_labels"));
    if (!out) return tess_fail(TESS_ERR_ARG, "tess_get_labels: out is null");
    const long n = c->source == SRC_IMAGE ? c->H * c->W : c->n_pts;
    long want = (n + 256L * 8 - 1) / (256L * 8);
    int grid = (int)(want < 1 ? 1 : (want > c->sm_count * 16 ? c->sm_count * 16 : want));
    TESS_LAUNCH((k_export_labels), grid, 256, S(stream), c->st, c->lab[0], c->lab[1], n, out);
    CU_TRY(cudaGetLastError());
    return TESS_OK;
}

extern "C" int tess_get_moments(tess_ctx* c, double* out, void* stream) {
    TRY(check_ready(c, "tess_get_moments"));
    if (!out) return tess_fail(TESS_ERR_ARG, "tess_get_moments: out is null");
    if (!c->mom) return tess_fail(TESS_ERR_STATE, "tess_get_moments: no accumulate has run");
    TESS_LAUNCH((k_export_moments), (c->k + 255) / 256, 256, S(stream), c->mom, n_moments(c), c->k, out);
    CU_TRY(cudaGetLastError());
    return TESS_OK;
}

extern "C" int tess_get_generators(tess_ctx* c, double* out, void* stream) {
    TRY(bind(c));
    if (c->k <= 0) return tess_fail(TESS_ERR_STATE, "tess_get_generators: no generators set");
    if (!out) return tess_fail(TESS_ERR_ARG, "tess_get_generators: out is null");
    CU_TRY(cudaMemcpyAsync(out, c->node, sizeof(double) * 2 * (size_t)c->k, cudaMemcpyDeviceToDevice, S(stream)));
    return TESS_OK;
}

extern "C" int tess_get_scales(tess_ctx* c, double* out, void* stream) {
    TRY(bind(c));
    if (c->k <= 0) return tess_fail(TESS_ERR_STATE, "tess_get_scales: no generators set");
    if (!out) return tess_fail(TESS_ERR_ARG, "tess_get_scales: out is null");
    if (c->has_scale)
        CU_TRY(cudaMemcpyAsync(out, c->scale, sizeof(double) * (size_t)c->k, cudaMemcpyDeviceToDevice, S(stream)));
    else
        TESS_LAUNCH((k_fill), (c->k + 255) / 256, 256, S(stream), out, c->k, 1.0);
    CU_TRY(cudaGetLastError());
    return TESS_OK;
}

extern "C" int tess_get_iter_stats(tess_ctx* c, double* delta, long* n_changed, long* n_iters_done, void* stream) {
    TRY(bind(c));
    TRY(fetch_state(c, S(stream)));
    if (delta) *delta = c->h_st->delta;
    if (n_changed) *n_changed = c->h_st->counted ? (long)c->h_st->n_changed : -1;
    if (n_iters_done) *n_iters_done = (long)c->h_st->iters_done;
    return TESS_OK;
}

extern "C" int tess_get_status(tess_ctx* c, int* converged, int* nonfinite, void* stream) {
    TRY(bind(c));
    TRY(fetch_state(c, S(stream)));
    if (converged) *converged = c->h_st->converged;
    if (nonfinite) *nonfinite = (c->h_st->flags & TESS_FLAG_NONFINITE) ? 1 : 0;
    return TESS_OK;
}

// ------------------------------------------------------------------------------------ assignment only
extern "C" int tess_assign_grid(tess_ctx* c, double x0, double y0, long H, long W, int32_t* labels_out, void* stream) {
    TRY(bind(c));
    if (c->k <= 0) return tess_fail(TESS_ERR_STATE, "tess_assign_grid: call tess_set_generators first");
    if (H <= 0 || W <= 0 || !labels_out) return tess_fail(TESS_ERR_ARG, "tess_assign_grid: bad arguments");
    cudaStream_t s = S(stream);
    Geometry q;
    TRY(image_geometry(c, x0, y0, H, W, s, &q));
    TRY(build_index(c, q, 0, nullptr, 0, s));
    TRY(build_lists(c, q, 0, 0, s));
    TRY(run_image(c, q, 0, 0, labels_out, s));
    return TESS_OK;
}

extern "C" int tess_assign_points(tess_ctx* c, long m, const double* xy, int32_t* out, void* stream) {
    TRY(bind(c));
    if (c->k <= 0) return tess_fail(TESS_ERR_STATE, "tess_assign_points: call tess_set_generators first");
    if (m <= 0 || !xy || !out) return tess_fail(TESS_ERR_ARG, "tess_assign_points: bad arguments");
    cudaStream_t s = S(stream);
    double box[4];
    TRY(points_bbox(c, m, xy, box, s));
    TRY(generator_bbox(c, s));
    Geometry q = box_geometry(c, box);
    TRY(build_index(c, q, 0, nullptr, 0, s));
    TRY(run_points(c, q, 0, 0, m, xy, nullptr, nullptr, out, s));
    return TESS_OK;
}

extern "C" int tess_label_sums(tess_ctx* c, long n, const int32_t* labels, const double* weights, long k, double* out,
                               void* stream) {
    TRY(bind(c));
    if (n < 0 || k <= 0 || !out || (n > 0 && !labels)) return tess_fail(TESS_ERR_ARG, "tess_label_sums: bad arguments");
    cudaStream_t s = S(stream);
    CU_TRY(cudaMemsetAsync(out, 0, sizeof(double) * (size_t)k, s));
    if (n > 0) {
        const long threads = (n + TESS_RUN - 1) / TESS_RUN;
        TESS_LAUNCH((k_label_sums), (int)((threads + 255) / 256), 256, s, labels, weights, n, k, out);
        CU_TRY(cudaGetLastError());
    }
    return TESS_OK;
}

extern "C" int tess_label_moments(tess_ctx* c, long H, long W, const int32_t* labels, const double* weights, long k,
                                  double* out, void* stream) {
    TRY(bind(c));
    if (H < 0 || W < 0 || k <= 0 || !out || (H * W > 0 && !labels)) return tess_fail(TESS_ERR_ARG, "tess_label_moments: bad arguments");
    cudaStream_t s = S(stream);
    CU_TRY(cudaMemsetAsync(out, 0, sizeof(double) * 4 * (size_t)k, s));
    if (H * W > 0) {
        const long threads = H * ((W + TESS_RUN - 1) / TESS_RUN);
        TESS_LAUNCH((k_label_moments), (int)((threads + 255) / 256), 256, s, labels, weights, H, W, k, out);
        CU_TRY(cudaGetLastError());
    }
    return TESS_OK;
}

extern "C" int tess_assign_grid_exhaustive(tess_ctx* c, double x0, double y0, long H, long W, int32_t* labels_out,
                                           void* stream) {
    TRY(bind(c));
    if (c->k <= 0) return tess_fail(TESS_ERR_STATE, "tess_assign_grid_exhaustive: call tess_set_generators first");
    if (H <= 0 || W <= 0 || !labels_out) return tess_fail(TESS_ERR_ARG, "tess_assign_grid_exhaustive: bad arguments");
    const long n = H * W;
    const int nb = (int)((n + 255) / 256);
    if (c->has_scale)
        TESS_LAUNCH((k_assign_exhaustive<true, true>), nb, 256, S(stream), c->node, c->inv_s2, c->k, n, W, x0, y0, nullptr, labels_out);
    else
        TESS_LAUNCH((k_assign_exhaustive<false, true>), nb, 256, S(stream), c->node, c->inv_s2, c->k, n, W, x0, y0, nullptr, labels_out);
    CU_TRY(cudaGetLastError());
    return TESS_OK;
}

extern "C" int tess_assign_points_exhaustive(tess_ctx* c, long m, const double* xy, int32_t* out, void* stream) {
    TRY(bind(c));
    if (c->k <= 0) return tess_fail(TESS_ERR_STATE, "tess_assign_points_exhaustive: call tess_set_generators first");
    if (m <= 0 || !xy || !out) return tess_fail(TESS_ERR_ARG, "tess_assign_points_exhaustive: bad arguments");
    const int nb = (int)((m + 255) / 256);
    if (c->has_scale)
        TESS_LAUNCH((k_assign_exhaustive<true, false>), nb, 256, S(stream), c->node, c->inv_s2, c->k, m, 1, 0.0, 0.0, xy, out);
    else
        TESS_LAUNCH((k_assign_exhaustive<false, false>), nb, 256, S(stream), c->node, c->inv_s2, c->k, m, 1, 0.0, 0.0, xy, out);
    CU_TRY(cudaGetLastError());
    return TESS_OK;
}
