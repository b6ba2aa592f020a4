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

#ifndef TESS_EMU
__global__ void __launch_bounds__(256)
k_peer_sum_p2p_sparse(TessPeers P, long lo, long hi, int M, long n_rows_doubles, long range_off) {
    __shared__ unsigned long long s_lo[TESS_PEER_MAX], s_hi[TESS_PEER_MAX];
    if (threadIdx.x < P.world) {
        unsigned long long a, b;
        asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(P.ptr[threadIdx.x] + range_off) : "memory");
        s_lo[threadIdx.x] = a;
        s_hi[threadIdx.x] = b;
    }
    __syncthreads();
    const long stride = 2L * gridDim.x * blockDim.x;
    for (long i = lo + 2L * ((long)blockIdx.x * blockDim.x + threadIdx.x); i < hi; i += stride) {
        const bool tail = i >= n_rows_doubles;                                     // the tail is summed densely
        const unsigned long long row = (unsigned long long)(i / M);
        double sa = 0.0, sb = 0.0;
        double a[TESS_PEER_MAX], b[TESS_PEER_MAX];
#pragma unroll
        for (int r = 0; r < TESS_PEER_MAX; ++r) {
            a[r] = 0.0; b[r] = 0.0;
            if (r < P.world && (tail || (row >= s_lo[r] && row < s_hi[r])))
                asm volatile("ld.relaxed.sys.global.v2.f64 {%0, %1}, [%2];" : "=d"(a[r]), "=d"(b[r]) : "l"(P.ptr[r] + i) : "memory");
        }
#pragma unroll
        for (int r = 0; r < TESS_PEER_MAX; ++r)
            if (r < P.world) { sa = __dadd_rn(sa, a[r]); sb = __dadd_rn(sb, b[r]); }
#pragma unroll
        for (int r = 0; r < TESS_PEER_MAX; ++r)
            if (r < P.world)
                asm volatile("st.relaxed.sys.global.v2.f64 [%0], {%1, %2};" ::"l"(P.ptr[r] + i), "d"(sa), "d"(sb) : "memory");
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
