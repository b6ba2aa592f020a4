// tess_update.cuh -- the O(k) tail of a Lloyd iteration and the convergence test.
//
//   k_count_prefilter : did any bin's population change since the previous iteration?
//   k_compare_labels  : only if not -- count labels that differ from the previous iteration;
//                       its last block writes the local "changed" flag into the tail slot of the
//                       moment vector (so a multi-GPU host sums it in the same allreduce as the
//                       moments)
//   k_update_nodes    : node update (reference tess/lloyd.pyx:66-74), delta (lloyd.pyx:76-81),
//                       convergence latch (lloyd.pyx:85-86), optional WVT scale update
//
// Convergence.  The reference stops when the node table is an exact fixed point (delta == 0,
// lloyd.pyx:85).  With its sequential sums that happens exactly when the labels of two
// consecutive iterations are identical.  The GPU sums are order-nondeterministic in the last
// bits, so the latch uses the order-independent form of the same test: no label changed since
// the previous iteration.  When the latch closes the node table is NOT rewritten (and delta is
// reported as 0), which reproduces the reference's bitwise node_{t+1} == node_t.
#pragma once
#include "tess_common.cuh"
#include "tess_index.cuh"

// moment-vector tail layout (after k*M doubles)
#define TESS_TAIL 2
#define TESS_TAIL_CHANGED 0   // sum over ranks of (labels changed ? 1 : 0)
#define TESS_TAIL_SPARE 1

// `range` (nullable): first row / last row + 1 with a non-zero count go there (atomic min / max), for the
// sparse multi-GPU exchange (tess_peer.cuh).
TESS_DEVFN void tess_p_count_prefilter(TessState* st, const double* mom, int M, int k, double* prev_count, double* mom_tail,
                                       unsigned long long* range = nullptr) {
    unsigned long long lo = ~0ull, hi = 0ull;
    // whole warps walk the table together (the vote below), one store per warp at most (every writer
    // stores the same value; a store per thread would serialise in L2).  The tail slot was zeroed
    // together with the moments.
    for (long base = (long)blockIdx.x * blockDim.x; base < k; base += (long)gridDim.x * blockDim.x) {
        const long j = base + threadIdx.x;
        bool diff = false;
        if (j < k) {
            const double c = mom[j * M];
            diff = (c != prev_count[j]);
            prev_count[j] = c;
            if (c != 0.0) {
                const unsigned long long u = (unsigned long long)j;
                lo = lo < u ? lo : u;
                hi = hi > u + 1 ? hi : u + 1;
            }
        }
        if (__any_sync(TESS_FULL, diff) && (threadIdx.x & 31) == 0) {
            st->count_changed = 1;
            mom_tail[TESS_TAIL_CHANGED] = 1.0;
        }
    }
    if (range) {
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) {
            const unsigned long long l2 = __shfl_xor_sync(TESS_FULL, lo, m), h2 = __shfl_xor_sync(TESS_FULL, hi, m);
            lo = lo < l2 ? lo : l2;
            hi = hi > h2 ? hi : h2;
        }
        if ((threadIdx.x & 31) == 0 && hi > 0ull) {
            atomicMin(range, lo);
            atomicMax(range + 1, hi);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        st->last_parity = (int)(st->iters_done & 1);
        if (!st->have_prev) mom_tail[TESS_TAIL_CHANGED] = 1.0;   // nothing to compare with yet
    }
}
__global__ void __launch_bounds__(256)
k_count_prefilter(TessState* st, int honor, const double* mom, int M, int k, double* prev_count, double* mom_tail) {
    if (TESS_STOPPED(st, honor)) return;
    tess_p_count_prefilter(st, mom, M, k, prev_count, mom_tail);
}

// Compares the two label buffers on the pixels/points that take part in the sums.
// MASK_MODE 0: all; 1: finite density; 2: finite signal, noise and weight
// No fences, no "last block" tail: the flag is an idempotent store, ordering comes from the
// kernel boundaries / grid barriers (a device-wide fence in a streaming kernel drains the SM's
// memory pipeline and costs far more than an extra launch).
TESS_DEVFN void tess_p_compare_labels(TessState* st, int mask_mode, int32_t* lab0, int32_t* lab1, long n, const double* dens,
                                      const double* sig, const double* noise, int uniform_weight, double* mom_tail) {
    const bool need = (!st->count_changed) && st->have_prev;
    if (blockIdx.x == 0 && threadIdx.x == 0) st->counted = need ? 1 : 0;
    if (!need) return;
    const int32_t* cur = (st->iters_done & 1) ? lab1 : lab0;
    const int32_t* prev = (st->iters_done & 1) ? lab0 : lab1;
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long stride = (long)gridDim.x * blockDim.x;
    unsigned long long diff = 0;
    for (; i < n; i += stride) {
        bool ok = true;
        if (mask_mode == 1) ok = isfinite(dens[i]);
        if (mask_mode == 2) {
            const double S = sig[i], N = noise[i];
            double wv = 1.0;
            if (!uniform_weight) wv = tess_sn_weight(S, N);
            ok = isfinite(S) && isfinite(N) && isfinite(wv);
        }
        if (ok && cur[i] != prev[i]) ++diff;
    }
    for (int m = 16; m >= 1; m >>= 1) diff += __shfl_xor_sync(TESS_FULL, diff, m);
    if ((threadIdx.x & 31) == 0 && diff) {
        atomicAdd(&st->n_changed, diff);
        mom_tail[TESS_TAIL_CHANGED] = 1.0;
    }
}
template <int MASK_MODE>
__global__ void __launch_bounds__(256)
k_compare_labels(TessState* st, int honor, int32_t* lab0, int32_t* lab1, long n, const double* dens,
                 const double* sig, const double* noise, int uniform_weight, double* mom_tail) {
    if (TESS_STOPPED(st, honor)) return;
    tess_p_compare_labels(st, MASK_MODE, lab0, lab1, n, dens, sig, noise, uniform_weight, mom_tail);
}

// 128-bit global load the compiler may not split or sink below a dependent branch
TESS_DEVFN double2 tess_ld2(const double* p) {
#ifndef TESS_EMU
    double2 v;
    asm volatile("ld.global.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
#else
    return *reinterpret_cast<const double2*>(p);
#endif
}

// One thread per generator (grid-stride).  `mom` holds the (globally reduced) moments, `mom_tail` the
// summed "changed" flags.  Nodes are rewritten in place unless the latch closes.
// With `cell_cnt` the thread also drops its NEW node into the histogram of the generator grid `g` (the
// first step of the next iteration's counting sort: it needs nothing but the node just computed).
template <int M>
TESS_DEVFN void tess_p_update_nodes(TessState* st, const double* mom, const double* mom_tail, int k, double* node,
                                    double* scale, double* inv_s2, int wvt_update, double* delta_part,
                                    const TessGrid* g = nullptr, int* cell_cnt = nullptr, int* cell_of = nullptr,
                                    const double* staged = nullptr) {
    __shared__ double s_red[8];
    const bool changed = mom_tail[TESS_TAIL_CHANGED] != 0.0;
    if (changed) {
        double d = 0.0;
        const long stride = (long)gridDim.x * blockDim.x;
        long j = (long)blockIdx.x * blockDim.x + threadIdx.x;
        // the three 16-byte operands of a generator (moment row as 128-bit loads -- rows are 32 or 48
        // bytes, 16-byte aligned -- and the old node) are loaded together, and those of the thread's
        // NEXT generator are requested before this one is worked on: issued after the divide (one more
        // dependent DRAM round trip, interleaved with the node stores) the kernel measured 5x slower
        // `staged` (multi-GPU, tess_peer.cuh): the new nodes were computed by the rank that owns the bin's
        // slice of the moment vector -- by the very rule below, from the totals -- and broadcast; they
        // are taken as they are (16 bytes per bin instead of the 32-byte row)
        double2 n01 = make_double2(0.0, 0.0), n23 = n01, nold = n01;
        if (j < k) {
            if (staged) n23 = tess_ld2(staged + 2 * j);
            else { n01 = tess_ld2(mom + j * M); n23 = tess_ld2(mom + j * M + 2); }
            nold = tess_ld2(node + 2 * j);
        }
        for (; j < k; j += stride) {
            const double* m = mom + j * M;
            const double2 m01 = n01, m23 = n23, old = nold;
            if (j + stride < k) {
                if (staged) n23 = tess_ld2(staged + 2 * (j + stride));
                else { n01 = tess_ld2(mom + (j + stride) * M); n23 = tess_ld2(mom + (j + stride) * M + 2); }
                nold = tess_ld2(node + 2 * (j + stride));
            }
            double2* np = reinterpret_cast<double2*>(node) + j;
            double nx = 0.0, ny = 0.0;
            if (staged) {
                nx = m23.x; ny = m23.y;
            } else if (m01.x > 0.0) {   // population test, not weight (lloyd.pyx:68)
                nx = __ddiv_rn(m23.x, m01.y);
                ny = __ddiv_rn(m23.y, m01.y);
            }
            // raised into flags_next, not flags: CTAs that start later still test the entry condition
            // against the flags of the previous kernel, so every node is rewritten (lloyd.pyx:66-74)
            if (!(isfinite(nx) && isfinite(ny))) atomicOr(&st->flags_next, TESS_FLAG_NONFINITE);
            const double dx = __dsub_rn(old.x, nx), dy = __dsub_rn(old.y, ny);
            d = __dadd_rn(d, __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
            *np = make_double2(nx, ny);
            if (cell_cnt) {
                const int c = tess_cell_coord(ny, g->oy, g->inv_cs, g->gh) * g->gw + tess_cell_coord(nx, g->ox, g->inv_cs, g->gw);
                cell_of[j] = c;
                atomicAdd(&cell_cnt[c], 1);
            }
            if (M == 6 && wvt_update && scale) {
                // WVT scale length (Diehl & Statler 2006 eq. 4, as in vorbin): sqrt(area / (S/N));
                // bins with no pixels or no signal keep their scale.
                const double sn = __ddiv_rn(m[4], __dsqrt_rn(m[5]));
                if (m01.x > 0.0 && sn > 0.0 && isfinite(sn)) {
                    const double sc = __dsqrt_rn(__ddiv_rn(m01.x, sn));
                    scale[j] = sc;
                    inv_s2[j] = __ddiv_rn(1.0, __dmul_rn(sc, sc));
                }
                atomicMin(&st->min_inv_next, (unsigned long long)__double_as_longlong(inv_s2[j]));
            }
        }
        // block reduction of delta (reporting only: the reference's sequential order is not reproduced)
        double s = tess_warp_sum(d);
        if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x < 32) {
            double v = (threadIdx.x < (blockDim.x >> 5)) ? s_red[threadIdx.x] : 0.0;
            v = tess_warp_sum(v);
            // one partial per CTA, summed by the finish step
            if (threadIdx.x == 0) delta_part[blockIdx.x] = v;
        }
    }
    // "this update ran" marker for the finish step (a flag raised in the loop above must not make
    // it skip the bookkeeping of this iteration)
    if (blockIdx.x == 0 && threadIdx.x == 0) st->update_ran = 1u;
}
template <int M>
__global__ void __launch_bounds__(256)
k_update_nodes(TessState* st, int honor, const double* mom, const double* mom_tail, int k, double* node,
               double* scale, double* inv_s2, int wvt_update, double* delta_part) {
    if (TESS_STOPPED(st, honor)) return;
    tess_p_update_nodes<M>(st, mom, mom_tail, k, node, scale, inv_s2, wvt_update, delta_part);
}

// One CTA, right after the node update: sums the per-CTA deltas and advances the iteration state
// (latch lloyd.pyx:85-86).  A separate step instead of a "last block done" tail: the device-wide
// fences that pattern needs cost ~90 us per update at k = 1e6 on cold lines.
TESS_DEVFN void tess_p_update_finish(TessState* st, const double* mom_tail, const double* delta_part, int n_part, int wvt_scales) {
    if (st->update_ran != 1u) return;   // the node update was a no-op (latch closed or error flag up)
    __shared__ double s_red[8];
    const bool changed = mom_tail[TESS_TAIL_CHANGED] != 0.0;
    double tot = 0.0;
    if (changed) {
        for (int i = threadIdx.x; i < n_part; i += blockDim.x) tot += delta_part[i];
        tot = tess_warp_sum(tot);
        if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = tot;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (!changed) {
            st->converged = 1;
            st->conv_iter = st->iters_done;
            st->delta = 0.0;
        } else {
            tot = 0.0;
            for (int i = 0; i < (int)(blockDim.x >> 5); ++i) tot += s_red[i];
            st->delta = tot;
            if (wvt_scales) st->min_inv_s2 = __longlong_as_double((long long)st->min_inv_next);
        }
        st->iters_done += 1;
        st->have_prev = 1;
        st->update_ran = 0u;
        st->flags |= st->flags_next;      // (flags_next itself is cleared with the per-iteration scalars: other
                                          //  CTAs of a phased launch may still be reading it)
    }
}
__global__ void __launch_bounds__(256)
k_update_finish(TessState* st, const double* mom_tail, const double* delta_part, int n_part, int wvt_scales) {
    tess_p_update_finish(st, mom_tail, delta_part, n_part, wvt_scales);
}

// moment rows of the previous iteration -> 0 (point mode; image mode clears them in the list builder)
__global__ void __launch_bounds__(256)
k_clear_moments(TessState* st, int honor, double* mom, long n) {
    if (TESS_STOPPED(st, honor)) return;
    for (long j = (long)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (long)gridDim.x * blockDim.x) mom[j] = 0.0;
    if (honor && blockIdx.x == 0 && threadIdx.x == 0) tess_reset_accumulate_scalars(st);
}

// finiteness of the generator table / scales, inv_s2 (run at set_generators)
__global__ void __launch_bounds__(256)
k_check_generators(TessState* st, const double* node, const double* scale, double* inv_s2, int k) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= k) return;
    bool bad = !(isfinite(node[2 * (long)j]) && isfinite(node[2 * (long)j + 1]));
    if (scale) {
        const double s = scale[j];
        const double inv = __ddiv_rn(1.0, __dmul_rn(s, s));
        inv_s2[j] = inv;
        if (!(isfinite(inv) && inv > 0.0)) bad = true;
    }
    if (bad) atomicOr(&st->flags, TESS_FLAG_NONFINITE);
}

// min over inv_s2 (the largest scale length bounds the search radius); single CTA
__global__ void __launch_bounds__(1024)
k_min_inv_s2(TessState* st, const double* inv_s2, int k) {
    __shared__ double s_red[32];
    double v = INFINITY;
    for (int j = threadIdx.x; j < k; j += blockDim.x) v = fmin(v, inv_s2[j]);
    v = tess_warp_min(v);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        double u = (threadIdx.x < (blockDim.x >> 5)) ? s_red[threadIdx.x] : INFINITY;
        u = tess_warp_min(u);
        if (threadIdx.x == 0) st->min_inv_s2 = u;
    }
}

// moments [k][M] -> [k][6] with zero padding (tess_get_moments)
__global__ void __launch_bounds__(256)
k_export_moments(const double* mom, int M, int k, double* out) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= k) return;
    for (int c = 0; c < 6; ++c) out[(long)j * 6 + c] = (c < M) ? mom[(long)j * M + c] : 0.0;
}

// widen / copy labels of the iteration that ran last (buffer parity lives on the device)
__global__ void __launch_bounds__(256)
k_export_labels(const TessState* st, const int32_t* lab0, const int32_t* lab1, long n, int32_t* out) {
    const int32_t* cur = st->last_parity ? lab1 : lab0;
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long stride = (long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = cur[i];
}

// Per-bin sums over a label array (np.bincount of tess/voronoi.py:150,195).  Each thread walks
// TESS_RUN consecutive elements and flushes one reduction per same-label run: segmentation maps
// are spatially coherent, so most threads issue a single atomic.
#define TESS_RUN 16
__global__ void __launch_bounds__(256)
k_label_sums(const int32_t* labels, const double* w, long n, long k, double* out) {
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long b = t * TESS_RUN;
    if (b >= n) return;
    const long e = (b + TESS_RUN < n) ? b + TESS_RUN : n;
    int cur = -1;
    double acc = 0.0;
    for (long i = b; i < e; ++i) {
        const int l = labels[i];
        if (l != cur) {
            if (cur >= 0 && cur < k) atomicAdd(out + cur, acc);
            cur = l;
            acc = 0.0;
        }
        acc = __dadd_rn(acc, w ? w[i] : 1.0);
    }
    if (cur >= 0 && cur < k) atomicAdd(out + cur, acc);
}

// Per-label count, sum w, sum w*row, sum w*col over a row-major [H][W] label image in ONE pass, pixel
// coordinates implicit (PixelAccretor.centroids, tess/pixel_accretion.py:212-240: a pixel always counts,
// its weight only if finite; w == nullptr means w = 1).  Same run aggregation as k_label_sums; a run
// ends with the row.  out is [k][4], zeroed by the caller.
__global__ void __launch_bounds__(256)
k_label_moments(const int32_t* labels, const double* w, long H, long W, long k, double* out) {
    const long per_row = (W + TESS_RUN - 1) / TESS_RUN;
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= H * per_row) return;
    const long row = t / per_row, c0 = (t - row * per_row) * TESS_RUN;
    const long c1 = (c0 + TESS_RUN < W) ? c0 + TESS_RUN : W;
    int cur = -1;
    double n = 0.0, sw = 0.0, swc = 0.0;
    auto flush = [&]() {
        if (cur >= 0 && cur < k) {
            double* m = out + 4 * (long)cur;
            atomicAdd(m, n);
            atomicAdd(m + 1, sw);
            atomicAdd(m + 2, __dmul_rn(sw, (double)row));
            atomicAdd(m + 3, swc);
        }
    };
    for (long c = c0; c < c1; ++c) {
        const int l = labels[row * W + c];
        if (l != cur) {
            flush();
            cur = l; n = 0.0; sw = 0.0; swc = 0.0;
        }
        const double wv = w ? w[row * W + c] : 1.0;
        n += 1.0;
        if (isfinite(wv)) { sw = __dadd_rn(sw, wv); swc = __dadd_rn(swc, __dmul_rn(wv, (double)c)); }
    }
    flush();
}
