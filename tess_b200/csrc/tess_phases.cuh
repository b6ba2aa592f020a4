// tess_phases.cuh -- the O(k) steps of a Lloyd iteration in ONE cooperative launch.
//
// Between two launches of the image (or point) kernel an iteration does eight small things: "did a
// population change", "did a label change", node update (reference tess/lloyd.pyx:66-74), delta and
// latch (lloyd.pyx:76-90), and the counting sort of the moved generators into the grid (the
// reference's per-iteration cKDTree build, lloyd.pyx:54): histogram, block sums, scan, scatter.  As
// eight kernels they cost ~5 us each at k = 1e5 -- launch latency and ramp, not work.  Here they are
// phases of one kernel, separated by grid-wide barriers (cooperative launch, every CTA resident):
//
//   0 PREFILTER   population of any bin changed?                           tess_p_count_prefilter
//   1 COMPARE     only if not: any label changed?                          tess_p_compare_labels
//     -- multi-GPU: the launch ends here, the moment vector is summed over the ranks, the next starts at 2
//   2 UPDATE      nodes, delta partials, WVT scales, and each new node straight into the cell
//                 histogram of the next grid                               tess_p_update_nodes
//                 (a launch without an update -- the first grid of a table -- only histograms)
//   3 FINISH      one thread: delta, latch, iteration counter              tess_p_update_finish
//                 meanwhile: small grids are scanned right here            tess_p_cell_scan (no block sums)
//                            large ones get their block totals             tess_p_cell_blocksum
//   4 SCAN        large grids only                                         tess_p_cell_scan
//   5 SCATTER     generators into cell order                               tess_p_cell_scatter
//
// [lo, hi] selects the phases of a launch; `with_update` says whether phases 2 and 3 include the
// node update and its bookkeeping.  The SIMT emulator (tests/emu) has no grid barrier: the host
// launches one phase at a time there, and it does the same when a device cannot launch
// cooperatively.
#pragma once
#include "tess_common.cuh"
#include "tess_index.cuh"
#include "tess_update.cuh"

#ifndef TESS_EMU
#include <cooperative_groups.h>
#endif

enum { TESS_PH_PREFILTER = 0, TESS_PH_COMPARE = 1, TESS_PH_UPDATE = 2, TESS_PH_FINISH = 3, TESS_PH_SCAN = 4, TESS_PH_SCATTER = 5 };

struct TessPhaseArgs {
    TessState* st;
    // moments, nodes
    double* mom;            // [k][M] + tail
    int M, k;
    double* prev_count;
    double* node;
    double* scale;          // WVT scale lengths or null
    double* inv_s2;
    const double* inv_in;   // inv_s2 when the metric is scaled, else null (copied into cell order)
    int wvt_update;
    double* delta_part;
    unsigned long long* range;   // published row range of the sparse exchange, or null
    const double* staged;        // new nodes computed and broadcast by the slices' owners (multi-GPU), or null
    // label comparison
    int mask_mode;          // 0: all; 1: finite density; 2: finite signal, noise, weight
    int32_t *lab0, *lab1;
    long n;
    const double *dens, *sig, *noise;
    int uniform_weight;
    // generator grid
    TessGrid g;
    int n_cells;
    int *cell_cnt, *cell_of, *cell_start, *block_sum, *cell_items;
    double2* cell_xy;
    double* cell_inv;
};

TESS_DEVFN void tess_grid_barrier() {
#ifndef TESS_EMU
    cooperative_groups::this_grid().sync();
#endif
}

__global__ void __launch_bounds__(256)
k_iteration_phases(const TESS_GRID_CONSTANT TessPhaseArgs a, int lo, int hi, int with_update) {
    TessState* const st = a.st;
    double* const tail = a.mom + (long)a.k * a.M;
    const bool single = a.n_cells <= TESS_SINGLE_SCAN_MAX;
    // the latch as the launch finds it: phases up to the node update are no-ops once it closed
    const bool stopped0 = (st->converged | st->flags) != 0;
    bool dirty = false;       // something ran since the last grid barrier
    for (int ph = lo; ph <= hi; ++ph) {
        if (ph == TESS_PH_SCAN && single) continue;
        if (dirty) tess_grid_barrier();
        dirty = true;
        if (ph == TESS_PH_PREFILTER) {
            if (!stopped0) tess_p_count_prefilter(st, a.mom, a.M, a.k, a.prev_count, tail, a.range);
        } else if (ph == TESS_PH_COMPARE) {
            // Labels are compared only when no population changed, i.e. a handful of times near
            // convergence.  Every CTA reads the same answer here (the barrier above published it), so
            // the common case skips the phase AND the barrier after it.
            if (stopped0 || st->count_changed || !st->have_prev) { dirty = false; continue; }
            tess_p_compare_labels(st, a.mask_mode, a.lab0, a.lab1, a.n, a.dens, a.sig, a.noise, a.uniform_weight, tail);
        } else if (ph == TESS_PH_UPDATE) {
            if (stopped0) continue;
            if (!with_update) tess_p_cell_count(a.g, a.node, a.k, a.cell_cnt, a.cell_of, nullptr, 0);
            else if (a.M == 6) tess_p_update_nodes<6>(st, a.mom, tail, a.k, a.node, a.scale, a.inv_s2, a.wvt_update, a.delta_part,
                                                      &a.g, a.cell_cnt, a.cell_of);
            else tess_p_update_nodes<4>(st, a.mom, tail, a.k, a.node, nullptr, a.inv_s2, 0, a.delta_part, &a.g, a.cell_cnt, a.cell_of,
                                        a.staged);
        } else if (ph == TESS_PH_FINISH) {
            // The latch AFTER this phase, from values that no longer change during it: the scan belongs
            // to the NEXT iteration and must know now, while one thread of CTA 0 is still writing the
            // state.  (The moments are NOT cleared here: callers read them after an update; the list
            // builder / k_clear_moments of the next accumulate does it.  The tail slots, read right
            // here by every CTA, are cleared by the scatter phase.)
            bool stopped1 = stopped0;
            if (with_update && !stopped0) stopped1 = (tail[TESS_TAIL_CHANGED] == 0.0) || (st->flags_next != 0);
            __syncthreads();
            if (with_update && blockIdx.x == 0)
                tess_p_update_finish(st, tail, a.delta_part, (int)gridDim.x, (a.M == 6 && a.wvt_update && a.scale) ? 1 : 0);
            if (!stopped1) {
                if (!single) tess_p_cell_blocksum(a.cell_cnt, a.n_cells, a.block_sum);
                else tess_p_cell_scan(st, 1, a.cell_cnt, a.cell_start, a.n_cells, nullptr);
            }
        } else if (ph == TESS_PH_SCAN) {
            if (!((st->converged | st->flags) != 0)) tess_p_cell_scan(st, 1, a.cell_cnt, a.cell_start, a.n_cells, a.block_sum);
        } else if (ph == TESS_PH_SCATTER) {
            if (!((st->converged | st->flags) != 0)) {
                tess_p_cell_scatter(a.cell_of, a.k, a.cell_start, a.cell_cnt, a.node, a.inv_in, a.cell_items, a.cell_xy, a.cell_inv);
                if (blockIdx.x == 0 && threadIdx.x == 0) {
                    tail[TESS_TAIL_CHANGED] = 0.0;
                    tail[TESS_TAIL_SPARE] = 0.0;
                    if (a.range) { a.range[0] = ~0ull; a.range[1] = 0ull; }     // the published row range is spent
                }
            }
        }
    }
}
