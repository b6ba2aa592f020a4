// tess_peer.cuh -- the one exchange step of the sharded Lloyd iteration: the sum
// of the packed moment vector [k*M | tail] over the ranks of one NVSwitch node, written by this
// library instead of calling ncclAllReduce.
//
// Every rank holds the vector in a symmetric allocation (same size on every rank, peer-mapped; the
// host side gets it from torch.distributed._symmetric_memory -- plumbing).  Rank r owns slice r:
//
//   k_peer_sum_multimem  with a multicast mapping (NVLS): one `multimem.ld_reduce.add.f64` per element
//       makes the SWITCH read the element from all ranks and return the sum; one `multimem.st`
//       broadcasts it back into every rank's copy.  Per rank: n/G elements in, n/G out.
//   k_peer_sum_p2p       without multicast: the owner loads its slice from every peer over NVLink
//       (rank order 0..G-1, so every rank ends up with bit-identical sums), and stores the total into
//       every peer's copy (two-shot all-reduce in one kernel).
//
// Ordering across GPUs is the caller's: a device-side barrier over all ranks before the launch (all
// local sums final) and after it (all broadcasts landed).  Peer data is read with system-scope
// relaxed loads so that no stale L1 line of an earlier iteration can be returned.
#pragma once
#include "tess_common.cuh"

#define TESS_PEER_MAX 16

struct TessPeers {
    double* ptr[TESS_PEER_MAX];   // the vector on rank 0..world-1 (peer-mapped device addresses)
    int world, rank;
};

// ---- sparse variant.  With generators numbered in spatial (row-major cell) order -- the sharded
// driver renumbers them, tess_b200/sharded.py -- the bins a rank's pixel rows touch are one short
// range of ids: of the k rows of its moment vector about k / world are non-zero.  Every rank publishes
// that range (first, last row with a non-zero count) in two 64-bit slots behind the vector;
// k_peer_sum_p2p_sparse then reads a row only from the peers whose range holds it -- adding the exact
// zeros of the others would not change a bit -- so the reduce half of the exchange moves ~1/world of the
// dense traffic.  The broadcast half stays dense: every rank updates the whole (replicated) node table.
#define TESS_RANGE_SLOTS 2
// Behind the vector, in this order: 2 x TESS_PEER_MAX 64-bit flags (the opening and the closing
// cross-rank barrier of k_peer_sum_p2p_sparse: flag[b][r] = last exchange number rank r signalled),
// then the two range slots.
#define TESS_FLAG_SLOTS (2 * TESS_PEER_MAX)
#define TESS_SPARSE_EXTRA (TESS_FLAG_SLOTS + TESS_RANGE_SLOTS)
#define TESS_PEER_CHUNK 2048      // doubles per bulk copy (16 KB)

// range[0] = first row with a non-zero count, range[1] = last such row + 1 (empty: ~0, 0)
__global__ void __launch_bounds__(256)
k_touched_range(const TessState* st, const double* mom, int M, int k, unsigned long long* range) {
    if ((st->converged | st->flags) != 0) return;
    unsigned long long lo = ~0ull, hi = 0ull;
    for (long j = (long)blockIdx.x * blockDim.x + threadIdx.x; j < k; j += (long)gridDim.x * blockDim.x)
        if (mom[j * M] != 0.0) {
            const unsigned long long u = (unsigned long long)j;
            lo = lo < u ? lo : u;
            hi = hi > u + 1 ? hi : u + 1;
        }
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

__global__ void k_reset_range(unsigned long long* range) {
    range[0] = ~0ull;
    range[1] = 0ull;
}
// flags (zero: no exchange signalled yet) and range of a fresh caller buffer
__global__ void k_reset_sparse_slots(unsigned long long* extra) {
    for (int i = threadIdx.x; i < TESS_FLAG_SLOTS; i += blockDim.x) extra[i] = 0ull;
    if (threadIdx.x == 0) { extra[TESS_FLAG_SLOTS] = ~0ull; extra[TESS_FLAG_SLOTS + 1] = 0ull; }
}

#ifndef TESS_EMU
TESS_DEVFN void tess_flag_store(double* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
TESS_DEVFN unsigned long long tess_flag_load(const double* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// The whole exchange in ONE kernel, its two cross-rank barriers included (one process and one GPU
// per rank; every rank launches it once per iteration with the same `epoch`):
//   open   every rank tells every peer "my local sums (and my range) are final" and waits for the
//          same word from all of them -- every CTA waits by itself, no CTA waits for another CTA;
//   sum    rank r reads slice r from the peers whose published range holds the row, adds in rank
//          order, stores the total into every rank's copy;
//   close  the last CTA to finish tells every peer "my slice is everywhere" and waits for all peers
//          to say so: when the kernel ends the whole vector of this rank is final.
__global__ void __launch_bounds__(256)
k_peer_sum_p2p_sparse(TessPeers P, double* mc, long lo, long hi, int M, long n_rows_doubles, long flag_off, long range_off,
                      long node_off, unsigned long long epoch, unsigned* done) {
    __shared__ unsigned long long s_lo[TESS_PEER_MAX], s_hi[TESS_PEER_MAX];
    __shared__ int s_last;
    if (threadIdx.x < P.world) {
        if (blockIdx.x == 0) tess_flag_store(P.ptr[threadIdx.x] + flag_off + P.rank, epoch);
        while (tess_flag_load(P.ptr[P.rank] + flag_off + threadIdx.x) < epoch) { }
        unsigned long long a, b;
        asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(P.ptr[threadIdx.x] + range_off) : "memory");
        s_lo[threadIdx.x] = a;
        s_hi[threadIdx.x] = b;
    }
    __syncthreads();
    // Chunks of TESS_PEER_CHUNK doubles: the CTA sums a chunk into shared memory (predicated 16-byte
    // peer loads), then ONE bulk copy of the TMA unit per destination rank sends it out
    // (cp.async.bulk shared -> global: UBLKCP; 16 KB bursts over NVLink instead of 16-byte stores).
    // Two buffers: the copies of a chunk drain while the next one is summed.
    __shared__ __align__(128) double s_out[2][TESS_PEER_CHUNK];
    int buf = 0;
    if (mc) {
        // With a multicast mapping of the buffer (NVLS) the broadcast half is ONE 16-byte store per pair
        // of totals: the switch replicates it into every rank's copy, so a rank sends its slice once
        // instead of once per peer (multimem.st moves bits: the two doubles travel as four .f32 words).
        for (long i = lo + 2L * ((long)blockIdx.x * blockDim.x + threadIdx.x); i < hi; i += 2L * gridDim.x * blockDim.x) {
            const bool tail = i >= n_rows_doubles;
            const unsigned long long row = (unsigned long long)(i / M);
            double a[TESS_PEER_MAX], b[TESS_PEER_MAX];
#pragma unroll
            for (int r = 0; r < TESS_PEER_MAX; ++r) {
                a[r] = 0.0; b[r] = 0.0;
                if (r < P.world && (tail || (row >= s_lo[r] && row < s_hi[r])))
                    asm volatile("ld.relaxed.sys.global.v2.f64 {%0, %1}, [%2];" : "=d"(a[r]), "=d"(b[r]) : "l"(P.ptr[r] + i) : "memory");
            }
            double sa = 0.0, sb = 0.0;
#pragma unroll
            for (int r = 0; r < TESS_PEER_MAX; ++r)
                if (r < P.world) { sa = __dadd_rn(sa, a[r]); sb = __dadd_rn(sb, b[r]); }
            const long long ua = __double_as_longlong(sa), ub = __double_as_longlong(sb);
            asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc + i),
                         "r"((unsigned)ua), "r"((unsigned)(ua >> 32)), "r"((unsigned)ub), "r"((unsigned)(ub >> 32)) : "memory");
        }
    } else if (node_off >= 0) {
        // OWNER-COMPUTED UPDATE (4-moment rows; slices start on rows).  What every rank needs from a bin's
        // totals is its new node -- population > 0 ? (sum wx / sum w, sum wy / sum w) : (0, 0), reference
        // tess/lloyd.pyx:66-74 -- so the owner of the slice forms it right here and broadcasts 16 bytes per
        // bin instead of the 32-byte row: half the traffic in and out of every rank, and the replicated
        // update reads half as much.  The totals themselves stay with the owner (its own copy of the
        // slice); the tail (the summed "labels changed" flags) still goes to everyone.
        for (long c0 = lo + (long)blockIdx.x * TESS_PEER_CHUNK; c0 < hi; c0 += (long)gridDim.x * TESS_PEER_CHUNK) {
            const int len = (int)((hi - c0 < TESS_PEER_CHUNK) ? (hi - c0) : TESS_PEER_CHUNK);      // even
            const long row_end = (c0 + len < n_rows_doubles) ? c0 + len : n_rows_doubles;
            const int n_row = (row_end > c0) ? (int)((row_end - c0) / 4) : 0;
            if (threadIdx.x < P.world) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            __syncthreads();
            for (int o = 2 * threadIdx.x; o < len; o += 2 * blockDim.x) {
                const long i = c0 + o;
                const bool tail = i >= n_rows_doubles;
                const unsigned long long row = (unsigned long long)(i / M);
                double a[TESS_PEER_MAX], b[TESS_PEER_MAX];
#pragma unroll
                for (int r = 0; r < TESS_PEER_MAX; ++r) {
                    a[r] = 0.0; b[r] = 0.0;
                    if (r < P.world && (tail || (row >= s_lo[r] && row < s_hi[r])))
                        asm volatile("ld.relaxed.sys.global.v2.f64 {%0, %1}, [%2];" : "=d"(a[r]), "=d"(b[r]) : "l"(P.ptr[r] + i) : "memory");
                }
                double sa = 0.0, sb = 0.0;
#pragma unroll
                for (int r = 0; r < TESS_PEER_MAX; ++r)
                    if (r < P.world) { sa = __dadd_rn(sa, a[r]); sb = __dadd_rn(sb, b[r]); }
                *reinterpret_cast<double2*>(&s_out[buf][o]) = make_double2(sa, sb);
                if (!tail) {
                    *reinterpret_cast<double2*>(P.ptr[P.rank] + i) = make_double2(sa, sb);      // the owner keeps the totals
                } else {
#pragma unroll
                    for (int r = 0; r < TESS_PEER_MAX; ++r)
                        if (r < P.world)
                            asm volatile("st.relaxed.sys.global.v2.f64 [%0], {%1, %2};" ::"l"(P.ptr[r] + i), "d"(sa), "d"(sb) : "memory");
                }
            }
            __syncthreads();
            double2 nd[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int r = (int)threadIdx.x + u * (int)blockDim.x;
                nd[u] = make_double2(0.0, 0.0);
                if (r < n_row) {
                    const double2 m01 = *reinterpret_cast<const double2*>(&s_out[buf][4 * r]);
                    const double2 m23 = *reinterpret_cast<const double2*>(&s_out[buf][4 * r + 2]);
                    if (m01.x > 0.0) nd[u] = make_double2(__ddiv_rn(m23.x, m01.y), __ddiv_rn(m23.y, m01.y));
                }
            }
            __syncthreads();
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int r = (int)threadIdx.x + u * (int)blockDim.x;
                if (r < n_row) *reinterpret_cast<double2*>(&s_out[buf][2 * r]) = nd[u];
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            if (threadIdx.x < P.world) {
                if (n_row > 0)
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(P.ptr[threadIdx.x] + node_off + (c0 / 4) * 2),
                                 "r"((unsigned)__cvta_generic_to_shared(&s_out[buf][0])), "r"(n_row * 16) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            buf ^= 1;
        }
    } else
    for (long c0 = lo + (long)blockIdx.x * TESS_PEER_CHUNK; c0 < hi; c0 += (long)gridDim.x * TESS_PEER_CHUNK) {
        const int len = (int)((hi - c0 < TESS_PEER_CHUNK) ? (hi - c0) : TESS_PEER_CHUNK);      // even
        // the bulk copies that read this buffer two chunks ago are done reading it
        if (threadIdx.x < P.world) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        __syncthreads();
        for (int o = 2 * threadIdx.x; o < len; o += 2 * blockDim.x) {
            const long i = c0 + o;
            const bool tail = i >= n_rows_doubles;                                 // the tail is summed densely
            const unsigned long long row = (unsigned long long)(i / M);
            double a[TESS_PEER_MAX], b[TESS_PEER_MAX];
#pragma unroll
            for (int r = 0; r < TESS_PEER_MAX; ++r) {
                a[r] = 0.0; b[r] = 0.0;
                if (r < P.world && (tail || (row >= s_lo[r] && row < s_hi[r])))
                    asm volatile("ld.relaxed.sys.global.v2.f64 {%0, %1}, [%2];" : "=d"(a[r]), "=d"(b[r]) : "l"(P.ptr[r] + i) : "memory");
            }
            double sa = 0.0, sb = 0.0;
#pragma unroll
            for (int r = 0; r < TESS_PEER_MAX; ++r)
                if (r < P.world) { sa = __dadd_rn(sa, a[r]); sb = __dadd_rn(sb, b[r]); }
            *reinterpret_cast<double2*>(&s_out[buf][o]) = make_double2(sa, sb);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to the TMA unit
        __syncthreads();
        if (threadIdx.x < P.world) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(P.ptr[threadIdx.x] + c0),
                         "r"((unsigned)__cvta_generic_to_shared(&s_out[buf][0])), "r"(len * 8) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        buf ^= 1;
    }
    if (!mc && threadIdx.x < P.world) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");      // all copies of this CTA have landed
    // ---- close
    __threadfence_system();            // this thread's stores are ordered before the flag below
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(done, 1u);
        s_last = (t == gridDim.x - 1) ? 1 : 0;
        if (s_last) *done = 0u;
        __threadfence();
    }
    __syncthreads();
    if (s_last && threadIdx.x < P.world) {
        tess_flag_store(P.ptr[threadIdx.x] + flag_off + TESS_PEER_MAX + P.rank, epoch);
        while (tess_flag_load(P.ptr[P.rank] + flag_off + TESS_PEER_MAX + threadIdx.x) < epoch) { }
    }
}

__global__ void __launch_bounds__(256)
k_peer_sum_multimem(double* mc, long lo, long hi) {
    // multimem.ld_reduce has no vector form for f64: four independent 8-byte reductions per thread
    // and trip keep enough requests in flight to the switch
    const long stride = (long)gridDim.x * blockDim.x;
    for (long i = lo + (long)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += 4 * stride) {
        double v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (i + u * stride < hi)
                asm volatile("multimem.ld_reduce.relaxed.sys.global.add.f64 %0, [%1];" : "=d"(v[u]) : "l"(mc + i + u * stride) : "memory");
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (i + u * stride < hi)
                asm volatile("multimem.st.relaxed.sys.global.f64 [%0], %1;" ::"l"(mc + i + u * stride), "d"(v[u]) : "memory");
    }
}

__global__ void __launch_bounds__(256)
k_peer_sum_p2p(TessPeers P, long lo, long hi) {
    // pairs of doubles (16-byte accesses; lo is even, the host pads the vector to an even length)
    const long stride = 2L * gridDim.x * blockDim.x;
    for (long i = lo + 2L * ((long)blockIdx.x * blockDim.x + threadIdx.x); i < hi; i += stride) {
        double a[TESS_PEER_MAX], b[TESS_PEER_MAX];
#pragma unroll
        for (int r = 0; r < TESS_PEER_MAX; ++r)
            if (r < P.world)
                asm volatile("ld.relaxed.sys.global.v2.f64 {%0, %1}, [%2];" : "=d"(a[r]), "=d"(b[r]) : "l"(P.ptr[r] + i) : "memory");
        double sa = 0.0, sb = 0.0;
#pragma unroll
        for (int r = 0; r < TESS_PEER_MAX; ++r)
            if (r < P.world) { sa = __dadd_rn(sa, a[r]); sb = __dadd_rn(sb, b[r]); }
#pragma unroll
        for (int r = 0; r < TESS_PEER_MAX; ++r)
            if (r < P.world)
                asm volatile("st.relaxed.sys.global.v2.f64 [%0], {%1, %2};" ::"l"(P.ptr[r] + i), "d"(sa), "d"(sb) : "memory");
    }
}
#endif
