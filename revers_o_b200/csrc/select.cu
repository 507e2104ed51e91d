// Exact selection kernels around the scan: chunked top-K of ordering keys (bitonic, smem),
// threshold derivation, fp32 re-score + final ordering, and the cross-shard merge (K3).
// Together they replace the `np.argsort(scores)[::-1]` + threshold walk + `limit` cut of
// qdrant-local search behind core_system.py:659-664.
#include "common.cuh"
#include "select.cuh"
#include "scan_tc.cuh"
#include "ptx.cuh"

namespace rvo {

using ptx::grid_dependency_wait;
using ptx::grid_launch_dependents;

// Bitonic sort, descending, n a power of two, keys in shared memory, whole block cooperates.
__device__ __forceinline__ void bitonic_desc_u64(unsigned long long* s, int n) {
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < (n >> 1); t += blockDim.x) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int ixj = i | j;
                const bool desc = (i & k) == 0;
                const unsigned long long a = s[i], b = s[ixj];
                if ((a < b) == desc) {
                    s[i] = b;
                    s[ixj] = a;
                }
            }
            __syncthreads();
        }
    }
}

// Same result, for n <= blockDim.x: one key per thread, held in a register; the compare-exchange steps whose partner is
// in the same warp (j < 32) go through shuffles, only the others through shared memory.  512 keys: 10 shared-memory
// steps instead of 45 (each costs two block barriers), which is most of the latency of the selection kernels.
__device__ __forceinline__ void bitonic_desc_u64_reg(unsigned long long* s, int n) {
    const int t = threadIdx.x;
    unsigned long long x = t < n ? s[t] : 0ull;
    for (int k = 2; k <= n; k <<= 1) {
        const bool desc = (t & k) == 0;
        for (int j = k >> 1; j > 0; j >>= 1) {
            unsigned long long y;
            if (j >= 32) {
                __syncthreads();   // readers of the previous exchange are done
                if (t < n) s[t] = x;
                __syncthreads();
                y = t < n ? s[t ^ j] : 0ull;
            } else {
                y = __shfl_xor_sync(0xFFFFFFFFu, x, j);
            }
            const bool take_max = ((t & j) == 0) == desc;
            x = take_max ? (x > y ? x : y) : (x < y ? x : y);
        }
    }
    __syncthreads();
    if (t < n) s[t] = x;
    __syncthreads();
}

// entry point: keys in shared memory, descending, n a power of two, whole block, synchronised on return
__device__ __forceinline__ void sort_desc_u64(unsigned long long* s, int n) {
    if (n <= (int)blockDim.x) bitonic_desc_u64_reg(s, n);
    else bitonic_desc_u64(s, n);
}

__device__ __forceinline__ int next_pow2(int x) {
    int p = 2;
    while (p < x) p <<= 1;
    return p;
}

// fp32 re-score of the candidates s[0..have) whose tensor score can still reach the top-k (score >= cut): 8 lanes
// per candidate (four candidates per warp in flight, 16 independent 16-byte loads per lane for d = 1024), fp32
// query (normalised) x bf16 DB row, fp32 FMA, then a 3-step shuffle reduction inside the 8-lane group.  Candidates
// below the cut become empty keys.  Whole block; the caller synchronises afterwards.
__device__ __forceinline__ void rescore_candidates(unsigned long long* s, int have, float cut, const uint16_t* db, int d_pad,
                                                   const float* qv) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int g = lane >> 3, l8 = lane & 7;
    const int nchunk = d_pad >> 3;
    const int ntk = d_pad / kTileCols;
    const uint4* base = (const uint4*)db;  // tiled DB storage: 16-byte chunk c of a row lives in tile (row/128, c/8)
    for (int e0 = 0; e0 < have; e0 += nwarps * 4) {
        const int e = e0 + warp * 4 + g;
        const unsigned long long key = e < have ? s[e] : 0ull;
        const bool active = key != 0ull && key_score(key) >= cut;
        const uint32_t row = key_row(key);
        float acc = 0.f;
        if (active) {
            const uint4* r = base + ((size_t)(row >> 7) * ntk * kTileRows + (row & 127)) * 8;
            // the candidate rows are random 2 KB gathers from DRAM: issue eight 16-byte loads per lane before consuming any
            for (int c0 = l8; c0 < nchunk; c0 += 64) {
                uint4 v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int c = c0 + 8 * u;
                    v[u] = c < nchunk ? __ldg(r + (size_t)(c >> 3) * kTileRows * 8 + (c & 7)) : make_uint4(0, 0, 0, 0);
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int c = c0 + 8 * u;
                    if (c < nchunk) {
                        const float4 q0 = __ldg((const float4*)(qv + c * 8));
                        const float4 q1 = __ldg((const float4*)(qv + c * 8 + 4));
                        acc = fmaf(__uint_as_float(v[u].x << 16), q0.x, acc);
                        acc = fmaf(__uint_as_float(v[u].x & 0xFFFF0000u), q0.y, acc);
                        acc = fmaf(__uint_as_float(v[u].y << 16), q0.z, acc);
                        acc = fmaf(__uint_as_float(v[u].y & 0xFFFF0000u), q0.w, acc);
                        acc = fmaf(__uint_as_float(v[u].z << 16), q1.x, acc);
                        acc = fmaf(__uint_as_float(v[u].z & 0xFFFF0000u), q1.y, acc);
                        acc = fmaf(__uint_as_float(v[u].w << 16), q1.z, acc);
                        acc = fmaf(__uint_as_float(v[u].w & 0xFFFF0000u), q1.w, acc);
                    }
                }
            }
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, o);
        __syncwarp();
        if (l8 == 0 && e < have) s[e] = active ? make_key(acc, row) : 0ull;
    }
}

// ---- peer-memory exchange (PushArgs) ----------------------------------------------------------------------------
__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// element i of query q's list (and, from thread 0, its count) into this rank's slot of every peer
__device__ __forceinline__ void push_element(const PushArgs& pa, int q, int i, long long id, float score) {
    const size_t e = (size_t)q * pa.k + i, nk = (size_t)pa.nq * pa.k;
    for (int g = 0; g < pa.world; ++g) {
        if (g == pa.rank) continue;
        unsigned char* slot = pa.peer_slot[g];
        ((long long*)slot)[e] = id;
        ((float*)(slot + nk * 8))[e] = score;
    }
}
__device__ __forceinline__ void push_count(const PushArgs& pa, int q, int count) {
    const size_t nk = (size_t)pa.nq * pa.k;
    for (int g = 0; g < pa.world; ++g)
        if (g != pa.rank) ((int*)(pa.peer_slot[g] + nk * 12))[q] = count;
}
// Whole block, after the CTA's last store: the last CTA of the grid publishes the epoch in every peer's flag.  Each
// thread's stores are fenced at system scope before the block barrier; thread 0's counter increment orders after them.
__device__ __forceinline__ void push_complete(const PushArgs& pa) {
    if (pa.world <= 1) return;
    // the block barrier orders every thread's peer stores before thread 0; thread 0's system-scope fence is cumulative, so
    // they are visible to the peers before the counter / flag updates that follow (the cooperative-groups grid-sync idiom).
    // One fence per CTA instead of one per thread: a system-scope fence costs microseconds.
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        const unsigned int prev = atomicAdd(pa.done, 1u);
        if (prev == gridDim.x - 1) {
            __threadfence_system();
            *pa.done = 0u;                     // ready for the next use of this parity (two searches later)
            for (int g = 0; g < pa.world; ++g)      // the fence above orders them; a release per store would fence world times
                asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(pa.peer_flag[g]), "l"(pa.epoch) : "memory");
        }
    }
}

constexpr int kFusedMax = 2048;      // world * k the in-kernel merge holds in shared memory
struct FusedScratch {                // shared memory the in-kernel merge may overwrite once the CTA's own list is in registers
    long long* sid;                  // [kFusedMax]
    float* ssc;                      // [kFusedMax]
    int* scnt;                       // [kMaxPeers + 2]
};

__device__ __forceinline__ bool list_before(float sa, long long ia, float sb, long long ib) {
    return sa > sb || (sa == sb && ia < ib);
}

// Fused exchange + merge (PushArgs::fused).  Called by the whole CTA with element `threadIdx.x` of its list in registers
// (k <= blockDim.x) and the list's count (-1: overflowed).  Record of (source rank r, query q) in rank g's region:
// lists[g] + (r * nq_max + q) * rec_bytes = [k ids | k scores | count].
__device__ __forceinline__ void fused_exchange_merge(const FinalArgs& a, int q, long long my_id, float my_sc, int my_count,
                                                     const FusedScratch& fs) {
    const PushArgs& pa = a.push;
    const int tid = threadIdx.x, G = pa.world, k = pa.k;
    const size_t rec_off = ((size_t)pa.rank * pa.nq_max + q) * pa.rec_bytes;
    for (int g = 0; g < G; ++g) {
        unsigned char* rec = pa.lists[g] + rec_off;
        if (tid < k) {
            ((long long*)rec)[tid] = my_id;
            ((float*)(rec + (size_t)k * 8))[tid] = my_sc;
        }
        if (tid == 0) *(int*)(rec + (size_t)k * 12) = my_count;
    }
    // block barrier, then ONE system-scope fence (cumulative over the CTA's stores), then the per-query flags
    __syncthreads();
    if (tid < G) {
        // lanes 0..G-1 of warp 0: ONE fence instruction for the warp (cumulative over the barrier), then the G flag stores in
        // parallel as relaxed system-scope stores — `st.release.sys` per flag would pay a system-scope fence for each of them
        // (measured: +27 us per search at 8 GPUs)
        __threadfence_system();
        asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(pa.qflags[tid] + (size_t)pa.rank * pa.nq_max + q), "l"(pa.epoch)
                     : "memory");
    }
    if (tid == 0) {
        fs.scnt[kMaxPeers] = 0;      // timeout flag
        fs.scnt[kMaxPeers + 1] = 0;  // bad (overflowed) list seen
    }
    __syncthreads();
    const unsigned long long* myflags = pa.qflags[pa.rank];
    if (tid < G && tid != pa.rank) {
        unsigned long long t0, t1;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
        while (ld_acquire_sys_u64(myflags + (size_t)tid * pa.nq_max + q) < pa.epoch) {
            __nanosleep(100);
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
            if (t1 - t0 > pa.timeout_ns) {
                fs.scnt[kMaxPeers] = 1;
                break;
            }
        }
    }
    __syncthreads();
    int64_t* oid = a.out_ids + (size_t)q * (size_t)k;
    float* osc = a.out_scores + (size_t)q * (size_t)k;
    if (fs.scnt[kMaxPeers]) {        // a peer never published: report, do not hang or trap (the host raises / falls back)
        for (int i = tid; i < k; i += blockDim.x) {
            oid[i] = -1;
            osc[i] = -__int_as_float(0x7f800000);
        }
        if (tid == 0) a.out_counts[q] = -2;
        return;
    }
    const unsigned char* base = pa.lists[pa.rank];
    if (tid < G) {
        const int c = __ldcg((const int*)(base + ((size_t)tid * pa.nq_max + q) * pa.rec_bytes + (size_t)k * 12));
        if (c < 0) fs.scnt[kMaxPeers + 1] = 1;
        fs.scnt[tid] = c < 0 ? 0 : (c > k ? k : c);
    }
    __syncthreads();
    const int n = G * k;
    for (int e = tid; e < n; e += blockDim.x) {
        const int g = e / k, i = e - g * k;
        if (i < fs.scnt[g]) {
            const unsigned char* rec = base + ((size_t)g * pa.nq_max + q) * pa.rec_bytes;
            fs.sid[e] = __ldcg((const long long*)rec + i);
            fs.ssc[e] = __ldcg((const float*)(rec + (size_t)k * 8) + i);
        }
    }
    __syncthreads();
    const bool bad = fs.scnt[kMaxPeers + 1] != 0;
    int total = 0;
    for (int g = 0; g < G; ++g) total += fs.scnt[g];
    const int n_out = bad ? 0 : (total < k ? total : k);
    // every list is sorted by (score desc, id asc) and ids are unique: the merged position of an element is its own index plus,
    // for every other list, the number of that list's elements ranked before it (one binary search each)
    for (int e = tid; e < n; e += blockDim.x) {
        const int g = e / k, i = e - g * k;
        if (i >= fs.scnt[g] || bad) continue;
        const float sc = fs.ssc[e];
        const long long id = fs.sid[e];
        int pos = i;
        for (int h = 0; h < G && pos < k; ++h) {
            if (h == g) continue;
            int lo = 0, hi = fs.scnt[h];
            const float* hs = fs.ssc + h * k;
            const long long* hid = fs.sid + h * k;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (list_before(hs[mid], hid[mid], sc, id)) lo = mid + 1;
                else hi = mid;
            }
            pos += lo;
        }
        if (pos < k) {
            oid[pos] = id;
            osc[pos] = sc;
        }
    }
    for (int i = n_out + tid; i < k; i += blockDim.x) {
        oid[i] = -1;
        osc[i] = -__int_as_float(0x7f800000);
    }
    if (tid == 0) a.out_counts[q] = bad ? -1 : n_out;
}

// emit: first k keys of the sorted list with score >= score_threshold (the reference's walk stops at the first score
// below it; `score == threshold` is kept).  s_n must be 0 on entry; whole block.
__device__ __forceinline__ void emit_topk(const unsigned long long* s, int n_sorted, const FinalArgs& a, int q, int* s_n,
                                          const FusedScratch* fs = nullptr) {
    if (a.push.fused && fs) {
        // the shard's list goes straight into the exchange records; the caller's outputs receive the MERGED list
        const int i = threadIdx.x;
        const unsigned long long key = (i < a.k && i < n_sorted) ? s[i] : 0ull;
        const bool ok = key != 0ull && key_score(key) >= a.score_threshold;
        const long long id = ok ? (long long)key_row(key) + a.id_offset : -1;
        const float sc = ok ? key_score(key) : -__int_as_float(0x7f800000);
        const int cnt = __syncthreads_count(ok);
        fused_exchange_merge(a, q, id, sc, cnt, *fs);
        return;
    }
    int64_t* oid = a.out_ids + (size_t)q * (size_t)a.k;
    float* osc = a.out_scores + (size_t)q * (size_t)a.k;
    int n_out_local = 0;
    for (int i = threadIdx.x; i < a.k; i += blockDim.x) {
        const unsigned long long key = i < n_sorted ? s[i] : 0ull;
        const bool ok = key != 0ull && key_score(key) >= a.score_threshold;
        const long long id = ok ? (long long)key_row(key) + a.id_offset : -1;
        const float sc = ok ? key_score(key) : -__int_as_float(0x7f800000);
        oid[i] = id;
        osc[i] = sc;
        if (a.push.world > 1) push_element(a.push, q, i, id, sc);
        n_out_local += ok;
    }
    if (n_out_local) atomicAdd(s_n, n_out_local);
    __syncthreads();
    if (threadIdx.x == 0) {
        a.out_counts[q] = *s_n;
        if (a.push.world > 1) push_count(a.push, q, *s_n);
    }
    push_complete(a.push);
}

__device__ __forceinline__ void emit_overflow(const FinalArgs& a, int q, const FusedScratch* fs = nullptr) {
    if (a.push.fused && fs) {
        __syncthreads();
        fused_exchange_merge(a, q, -1, -__int_as_float(0x7f800000), -1, *fs);
        return;
    }
    int64_t* oid = a.out_ids + (size_t)q * (size_t)a.k;
    float* osc = a.out_scores + (size_t)q * (size_t)a.k;
    for (int i = threadIdx.x; i < a.k; i += blockDim.x) {
        oid[i] = -1;
        osc[i] = -__int_as_float(0x7f800000);
        if (a.push.world > 1) push_element(a.push, q, i, -1, -__int_as_float(0x7f800000));
    }
    if (threadIdx.x == 0) {
        a.out_counts[q] = -1;
        if (a.push.world > 1) push_count(a.push, q, -1);
    }
    push_complete(a.push);
}

// ---- exact top-K of one query's candidates by histogram refinement ---------------------------------
// One CTA per query.  The 64-bit ordering keys are unique (score, row), so the K-th largest key is
// found by narrowing a key range with 2048-bin histograms until at most kSelSort keys lie at or above
// the range's lower bound; those are compacted into shared memory and sorted.  A histogram over a small
// SAMPLE first proposes the lower bound, so that the full-data histogram only sees the few keys near
// the top (no shared-memory atomic contention, fine bins); if the proposal was too high the range is
// widened to the true minimum, so the result is exact in every case.  Work is O(n) per round instead of
// the O(n log^2 n) of sorting every candidate.
// Visit every key of query q (or every `sample_every`-th).  Loads are issued four at a time before any of them is
// consumed, so that a thread has four L2 round trips in flight instead of one (the loop body may contain atomics,
// which keeps the compiler from hoisting the loads itself).
template <class F>
__device__ __forceinline__ void for_each_key(const SelectArgs& a, int q, const int* s_cnt, int sample_every, F f) {
    if (a.dense) {
        const float* src = a.dense + (size_t)q * (size_t)a.dense_ld;
        const float NINF = -__int_as_float(0x7f800000);
        const int n = (int)a.n_dense, step = (int)blockDim.x * sample_every;
        for (int i0 = (int)threadIdx.x * sample_every; i0 < n; i0 += 4 * step) {
            float v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = (i0 + u * step) < n ? __ldg(src + i0 + u * step) : NINF;
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (v[u] != NINF) f(make_key(v[u], (uint32_t)(i0 + u * step)));  // -inf: rows past the end of the DB
        }
    } else {
        // all sub-lists of the query as one index space: thread t takes elements t, t + blockDim, ... of each.  A SAMPLE is the
        // first 1/sample_every of every sub-list (arrival order is unrelated to the score, and a contiguous prefix touches
        // 1/sample_every of the cache lines — a strided sample touches them all and costs as much as the full pass).
        const int step = (int)blockDim.x;
        for (int sg0 = 0; sg0 < a.nseg; sg0 += 8) {
            // eight sub-lists side by side: eight independent loads in flight per thread
            int lim[8], cmax = 0;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                lim[u] = sg0 + u < a.nseg ? (s_cnt[sg0 + u] + sample_every - 1) / sample_every : 0;
                cmax = lim[u] > cmax ? lim[u] : cmax;
            }
            for (int i = (int)threadIdx.x; i < cmax; i += step) {
                unsigned long long k[8];
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    k[u] = i < lim[u] ? a.keys[((size_t)q * a.nseg + sg0 + u) * (size_t)a.cap + (size_t)i] : 0ull;
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    if (k[u]) f(k[u]);
            }
        }
    }
}

__device__ __forceinline__ int hist_shift(unsigned long long lo, unsigned long long hi) {
    int shift = 0;
    while (((hi - lo) >> shift) >= (unsigned long long)kSelBins) ++shift;
    return shift;
}

// Warp 0: find the bin that holds the `need`-th largest key of the histogrammed range.
// s_res = {bin, keys at/above the bin's lower bound (within the range), keys in the bin, found}.
__device__ __forceinline__ void hist_find(const int* hist, int need, int lane, int* s_res) {
    constexpr int per = kSelBins / 32;  // lane l owns bins [per*l, per*l + per)
    int mine = 0;
    for (int b = 0; b < per; ++b) mine += hist[lane * per + b];
    int suffix = mine;  // inclusive suffix over lanes >= lane
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_down_sync(0xFFFFFFFFu, suffix, o);
        if (lane + o < 32) suffix += t;
    }
    const int next = suffix - mine;
    if (lane == 0 && suffix < need) s_res[3] = 0;
    if (suffix >= need && next < need) {  // the crossing lane
        int run = next, b = per - 1;
        for (; b >= 0; --b) {
            run += hist[lane * per + b];
            if (run >= need) break;
        }
        s_res[0] = lane * per + b;
        s_res[1] = run;
        s_res[2] = hist[lane * per + b];
        s_res[3] = 1;
    }
}

__device__ __forceinline__ void sel_stamp(const SelectArgs& a, int q, int slot) {
    if (a.trace && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        a.trace[(size_t)q * 16 + slot] = t;
    }
}

// fp32 re-score, one WARP per candidate row (hot path of the last level): lane l owns the 16-byte chunks l, l + 32, ... of
// the row, i.e. a warp reads whole 128-byte lines of the tiled DB; four candidates and two chunk columns are in flight per lane
// (8 independent 16-byte loads), and every warp takes a contiguous share of the candidates.  act[e] := key(fp32 score, row).
__device__ __forceinline__ void rescore_rows_warp(unsigned long long* act, int n, const uint16_t* db, int d_pad, const float* qv) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int nchunk = d_pad >> 3, ntk = d_pad / kTileCols;
    const uint4* base = (const uint4*)db;
    const int per = (n + nwarps - 1) / nwarps;
    const int e_lo = warp * per, e_hi = (e_lo + per) < n ? (e_lo + per) : n;
    for (int e0 = e_lo; e0 < e_hi; e0 += 4) {
        const uint4* r[4];
        uint32_t rows[4];
        float acc[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int e = (e0 + u) < e_hi ? (e0 + u) : e0;
            rows[u] = key_row(act[e]);
            r[u] = base + ((size_t)(rows[u] >> 7) * ntk * kTileRows + (rows[u] & 127)) * 8;
            acc[u] = 0.f;
        }
        for (int c0 = lane; c0 < nchunk; c0 += 64) {
            uint4 v[4][2];
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int w = 0; w < 2; ++w) {
                    const int c = c0 + 32 * w;
                    v[u][w] = c < nchunk ? __ldg(r[u] + (size_t)(c >> 3) * kTileRows * 8 + (c & 7)) : make_uint4(0, 0, 0, 0);
                }
#pragma unroll
            for (int w = 0; w < 2; ++w) {
                const int c = c0 + 32 * w;
                if (c < nchunk) {
                    const float4 q0 = *(const float4*)(qv + c * 8);
                    const float4 q1 = *(const float4*)(qv + c * 8 + 4);
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        float x = acc[u];
                        x = fmaf(__uint_as_float(v[u][w].x << 16), q0.x, x);
                        x = fmaf(__uint_as_float(v[u][w].x & 0xFFFF0000u), q0.y, x);
                        x = fmaf(__uint_as_float(v[u][w].y << 16), q0.z, x);
                        x = fmaf(__uint_as_float(v[u][w].y & 0xFFFF0000u), q0.w, x);
                        x = fmaf(__uint_as_float(v[u][w].z << 16), q1.x, x);
                        x = fmaf(__uint_as_float(v[u][w].z & 0xFFFF0000u), q1.y, x);
                        x = fmaf(__uint_as_float(v[u][w].w << 16), q1.z, x);
                        x = fmaf(__uint_as_float(v[u][w].w & 0xFFFF0000u), q1.w, x);
                        acc[u] = x;
                    }
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            float x = acc[u];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xFFFFFFFFu, x, o);
            if (lane == 0 && e0 + u < e_hi) act[e0 + u] = make_key(x, rows[u]);
        }
    }
}

constexpr int kHotActMax = kSelBins / 2;   // candidates re-scored by the hot path (they live in the histogram's shared memory)
constexpr int kHotQMax = 2048;             // query dimensions kept in shared memory by the hot path

// Last level, fast path: the HOT sub-lists alone (scan_tc.cuh).  They hold every candidate whose tensor score reached
// tau_hot[q] — a few hundred keys instead of the thousands admitted by the safe threshold.  One 2048-bin histogram of the
// scores gives a lower bound `edge` of the k-th best tensor score; every row that can still be in the fp32 top-k has tensor
// score >= edge - margin =: cut (common.cuh), and all of those are in the hot lists iff cut >= tau_hot.  They are re-scored
// in fp32, ranked by counting (keys are unique) and emitted.  Returns false (whole CTA, uniformly) when the hot lists do not
// cover the query — a hot sub-list overflowed, fewer than k hot keys, more than kHotActMax candidates inside the margin — and
// the caller takes the general path over ALL sub-lists.
__device__ __forceinline__ bool select_hot(const SelectArgs& a, const FinalArgs& f, int q, int* hist, unsigned long long* sbuf,
                                           float* s_q, int* s_res, int* s_count, const FusedScratch* fs) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int base[kHotSplit + 1];
    base[0] = 0;
    if (a.hot_nseg > kHotSplit) return false;
#pragma unroll
    for (int u = 0; u < kHotSplit; ++u) {
        int c = 0;
        if (u < a.hot_nseg) {
            c = a.cnt[(size_t)q * a.nseg + a.hot_seg0 + u];
            if (c > a.cap) return false;           // keys were dropped
        }
        base[u + 1] = base[u] + c;
    }
    const int n_hot = base[kHotSplit];
    if (n_hot < f.k || n_hot > kSelSort) return false;
    uint32_t omax = 0u;                             // largest (orderable) score: upper end of the histogram's range
#pragma unroll
    for (int u = 0; u < kHotSplit; ++u) {
        const unsigned long long* src = a.keys + ((size_t)q * a.nseg + a.hot_seg0 + u) * (size_t)a.cap;
        for (int i = tid; i < base[u + 1] - base[u]; i += blockDim.x) {
            const unsigned long long key = src[i];
            sbuf[base[u] + i] = key;
            omax = max(omax, (uint32_t)(key >> 32));
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) omax = max(omax, __shfl_xor_sync(0xFFFFFFFFu, omax, o));
    if (lane == 0 && omax) atomicMax((unsigned int*)s_count, omax);   // *s_count == 0 on entry (kernel prologue)
    const float* qg = f.qn + (size_t)q * (size_t)f.qn_ld;
    const bool q_smem = f.d_pad <= kHotQMax;
    if (q_smem)
        for (int i = tid; i < (f.d_pad >> 2); i += blockDim.x) ((float4*)s_q)[i] = __ldg((const float4*)qg + i);
    for (int i = tid; i < kSelBins; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    const float th = a.tau_hot[q];
    const uint32_t lo = f32_orderable(fmaxf(th, -2.0f)), hi = *(const unsigned int*)s_count;
    if (hi < lo) return false;                      // cannot happen: every hot key scored >= tau_hot
    int shift = 0;
    while (((hi - lo) >> shift) >= (uint32_t)kSelBins) ++shift;
    for (int i = tid; i < n_hot; i += blockDim.x) {
        const uint32_t o = (uint32_t)(sbuf[i] >> 32);
        const uint32_t b = o >= lo ? (o - lo) >> shift : 0u;
        atomicAdd(&hist[b < (uint32_t)kSelBins ? (int)b : kSelBins - 1], 1);
    }
    __syncthreads();
    if (warp == 0) hist_find(hist, f.k, lane, s_res);
    __syncthreads();
    if (!s_res[3]) return false;
    const float m = f.margin ? f.margin[q] : 0.f;
    const float cut = orderable_f32(lo + ((uint32_t)s_res[0] << shift)) - m;
    if (!(cut >= th)) return false;                // candidates between cut and tau_hot live in the safe sub-lists
    unsigned long long* act = (unsigned long long*)hist;   // the histogram is no longer needed
    if (tid == 0) *s_count = 0;
    __syncthreads();
    for (int i = tid; i < n_hot; i += blockDim.x) {
        const unsigned long long key = sbuf[i];
        if (key_score(key) >= cut) {
            const int at = atomicAdd(s_count, 1);
            if (at < kHotActMax) act[at] = key;
        }
    }
    __syncthreads();
    const int n_act = *s_count;
    if (n_act > kHotActMax) return false;
    sel_stamp(a, q, 5);
    rescore_rows_warp(act, n_act, f.db, f.d_pad, q_smem ? s_q : qg);
    __syncthreads();
    sel_stamp(a, q, 6);
    // final order by counting: keys are unique (score, row), so the rank of a key is the number of larger keys
    for (int i = tid; i < n_act; i += blockDim.x) {
        const unsigned long long my = act[i];
        int r = 0;
        for (int j = 0; j < n_act; ++j) r += act[j] > my;
        if (r < f.k) sbuf[r] = my;
    }
    if (tid == 0) *s_count = 0;
    __syncthreads();
    sel_stamp(a, q, 7);
    emit_topk(sbuf, n_act < f.k ? n_act : f.k, f, q, s_count, fs);
    sel_stamp(a, q, 8);
    if (a.trace && tid == 0) {
        a.trace[(size_t)q * 16 + 9] = (unsigned long long)n_hot;
        a.trace[(size_t)q * 16 + 10] = (unsigned long long)n_act;
        a.trace[(size_t)q * 16 + 11] = 0ull;
    }
    return true;
}

template <bool FINAL>
__global__ void __launch_bounds__(kSelThreads, 2) select_kernel(const SelectArgs a, const FinalArgs f) {
    __shared__ int hist[kSelBins];
    __shared__ unsigned long long sbuf[kSelSort];
    __shared__ unsigned long long s_red[2 * (kSelThreads / 32)];
    __shared__ unsigned long long s_lo, s_hi, s_min;
    __shared__ int s_above, s_count, s_done, s_cnt[32], s_res[4], s_over;
    __shared__ __align__(16) float s_q[FINAL ? kHotQMax : 4];
    const int q = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float INF = __int_as_float(0x7f800000);
    grid_dependency_wait();      // candidates / scores come from the scan that precedes this kernel in the stream
    grid_launch_dependents();
    if (q >= a.nq) {  // padded query rows of the tensor path never admit anything
        if (a.tau_out && tid == 0) a.tau_out[q] = INF;
        if (a.tau_hot_out && tid == 0) a.tau_hot_out[q] = INF;
        return;
    }
    if (tid == 0) {
        s_over = 0;
        s_count = 0;
    }
    __syncthreads();
    sel_stamp(a, q, 0);
    if (!a.dense && tid < a.nseg) {
        const int c = a.cnt[q * a.nseg + tid];
        s_cnt[tid] = c < a.cap ? c : a.cap;
        if (c > a.cap) s_over = 1;   // a sub-list dropped candidates
    }
    __syncthreads();
    FusedScratch fs_;
    fs_.sid = (long long*)sbuf;          // the in-kernel merge of the fused exchange reuses the selection's shared memory
    fs_.ssc = (float*)hist;
    fs_.scnt = s_cnt;
    const FusedScratch* fs = (FINAL && f.push.fused) ? &fs_ : nullptr;
    if constexpr (FINAL) {
        if (!a.dense && a.hot_nseg > 0 && select_hot(a, f, q, hist, sbuf, s_q, s_res, &s_count, fs)) return;
        __syncthreads();
        if (s_over) {
            emit_overflow(f, q, fs);
            return;
        }
    }

    // number of keys (dense: sample columns, the few -inf pads of the last tile included) and the key range
    if (tid == 0) {
        int nnz = 0;
        if (a.dense) nnz = (int)a.n_dense;
        else for (int sgm = 0; sgm < a.nseg; ++sgm) nnz += s_cnt[sgm];
        float lo_f = -2.0f;
        if (a.range_lo) lo_f = fmaxf(a.range_lo[q], -2.0f);   // every list key scored >= its admission threshold
        s_min = (unsigned long long)f32_orderable(lo_f) << 32;
        s_lo = s_min;
        s_hi = ((unsigned long long)f32_orderable(2.0f) << 32) | 0xFFFFFFFFull;
        s_above = 0;
        s_count = nnz;
        s_done = nnz <= kSelSort ? 1 : 0;  // few enough: keep every key
    }
    __syncthreads();
    const int nnz = s_count;
    const int want = a.K < nnz ? a.K : nnz;
    auto compact = [&](unsigned long long T) {   // keys >= T -> sbuf (first kSelSort of them), count -> s_count
        __syncthreads();
        if (tid == 0) s_count = 0;
        __syncthreads();
        for_each_key(a, q, s_cnt, 1, [&](unsigned long long k) {
            if (k >= T) {
                const int at = atomicAdd(&s_count, 1);
                if (at < kSelSort) sbuf[at] = k;
            }
        });
        __syncthreads();
    };

    bool fast = false;
    unsigned long long T_used = 0ull;   // sbuf holds every key >= T_used (unless s_count > kSelSort)
    if (s_done) {
        compact(0ull);
        fast = true;
    } else {
        // round S: a histogram over a sample (every `every`-th element of each sub-list / row) proposes the bound
        const int every = nnz > 8 * kSelSort ? 8 : (nnz > 4 * kSelSort ? 4 : 2);
        const unsigned long long lo0 = s_lo, hi0 = s_hi;
        const int sh0 = hist_shift(lo0, hi0);
        for (int i = tid; i < kSelBins; i += blockDim.x) hist[i] = 0;
        __syncthreads();
        for_each_key(a, q, s_cnt, every, [&](unsigned long long k) {
            if (k >= lo0) {
                const unsigned long long b = (k - lo0) >> sh0;   // NaN scores order above +2.0: clamp into the top bin
                atomicAdd(&hist[b < (unsigned long long)kSelBins ? (int)b : kSelBins - 1], 1);
            }
        });
        __syncthreads();
        if (warp == 0) {
            // aim at ~1.5x the wanted count above the bound (plus slack for sampling noise)
            sel_stamp(a, q, 1);   // sample histogram built
            hist_find(hist, (3 * want) / (2 * every) + 24, lane, s_res);
            __syncwarp();
            if (lane == 0) s_lo = s_res[3] ? lo0 + ((unsigned long long)s_res[0] << sh0) : 0ull;
        }
        __syncthreads();
        // one compaction pass with the proposed bound; exact whenever it kept between `want` and kSelSort keys
        const unsigned long long T0 = s_lo;
        sel_stamp(a, q, 2);       // bound proposed
        compact(T0);
        T_used = T0;
        sel_stamp(a, q, 3);       // compacted
        fast = s_count >= want && s_count <= kSelSort;
    }

    if (!fast) {
        // rigorous path: true key range, then histogram refinement until <= kSelSort keys lie at/above the bound
        unsigned long long lmin = ~0ull, lmax = 0ull;
        for_each_key(a, q, s_cnt, 1, [&](unsigned long long k) {
            lmin = k < lmin ? k : lmin;
            lmax = k > lmax ? k : lmax;
        });
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long x = __shfl_xor_sync(0xFFFFFFFFu, lmin, o), y = __shfl_xor_sync(0xFFFFFFFFu, lmax, o);
            lmin = x < lmin ? x : lmin;
            lmax = y > lmax ? y : lmax;
        }
        if (lane == 0) { s_red[warp] = lmin; s_red[kSelThreads / 32 + warp] = lmax; }
        __syncthreads();
        if (tid == 0) {
            unsigned long long mn = ~0ull, mx = 0ull;
            for (int w = 0; w < kSelThreads / 32; ++w) {
                mn = s_red[w] < mn ? s_red[w] : mn;
                mx = s_red[kSelThreads / 32 + w] > mx ? s_red[kSelThreads / 32 + w] : mx;
            }
            s_min = mn; s_lo = mn; s_hi = mx; s_above = 0; s_done = 0;
        }
        __syncthreads();
        for (int round = 0; round < 12 && !s_done; ++round) {
            const unsigned long long lo = s_lo, hi = s_hi;
            const int shift = hist_shift(lo, hi);
            __syncthreads();  // everyone has read s_lo/s_hi and finished with hist[] of the previous step
            for (int i = tid; i < kSelBins; i += blockDim.x) hist[i] = 0;
            __syncthreads();
            for_each_key(a, q, s_cnt, 1, [&](unsigned long long k) {
                if (k >= lo && k <= hi) atomicAdd(&hist[(int)((k - lo) >> shift)], 1);
            });
            __syncthreads();
            if (warp == 0) {
                const int need = want - s_above;
                hist_find(hist, need, lane, s_res);
                __syncwarp();
                if (lane == 0) {
                    if (!s_res[3]) {
                        s_lo = s_min;  // fewer than `want` keys exist (padded dense sample): take everything
                        s_done = 1;
                    } else {
                        const int cge = s_above + s_res[1];  // keys >= lower bound of the crossing bin
                        s_lo = lo + ((unsigned long long)s_res[0] << shift);
                        if (cge <= kSelSort || shift == 0) {
                            s_done = 1;
                        } else {
                            s_above = s_above + s_res[1] - s_res[2];  // keys strictly above the bin
                            s_hi = s_lo + ((1ull << shift) - 1ull);
                        }
                    }
                }
            }
            __syncthreads();
        }
        const unsigned long long T1 = s_lo;
        compact(T1);
        T_used = T1;
    }
    int C = s_count < kSelSort ? s_count : kSelSort;
    int ns = next_pow2(C > 2 ? C : 2);
    for (int i = C + tid; i < ns; i += blockDim.x) sbuf[i] = 0ull;
    __syncthreads();
    sort_desc_u64(sbuf, ns);
    sel_stamp(a, q, 4);           // first sort done

    if constexpr (FINAL) {
        if (!f.rescore) {
            // keys carry exact fp32 scores already (small-Q path): the sorted prefix IS the answer
            if (tid == 0) s_count = 0;
            __syncthreads();
            emit_topk(sbuf, C, f, q, &s_count, fs);
            return;
        }
        // ---- fused last level: every candidate that can still reach the top-k after the fp32 re-score is one whose
        // tensor score lies within the query's margin of the k-th best tensor score (common.cuh) ----
        const float NINF = -INF;
        const float m = f.margin ? f.margin[q] : 0.f;
        const float cut = C >= f.k ? key_score(sbuf[f.k - 1]) - m : NINF;
        // sbuf holds every key >= T (none was dropped if s_count <= kSelSort); keys with score >= cut are >= Kc
        const unsigned long long T = T_used;
        const unsigned long long Kc = cut == NINF ? 0ull : (unsigned long long)f32_orderable(cut) << 32;
        const bool covered = s_count <= kSelSort && (T <= Kc || C == nnz);
        if (!covered) {
            compact(Kc);                                  // a superset of the k best: the cut stays valid
            if (s_count > kSelSort) {                     // more than kSelSort candidates inside the margin
                emit_overflow(f, q, fs);
                return;
            }
            C = s_count;
            ns = next_pow2(C > 2 ? C : 2);
            for (int i = C + tid; i < ns; i += blockDim.x) sbuf[i] = 0ull;
            __syncthreads();
            sort_desc_u64(sbuf, ns);
        }
        // the candidates at or above the cut are a prefix of the sorted list
        if (tid == 0) s_above = 0;
        __syncthreads();
        int mine = 0;
        for (int i = tid; i < C; i += blockDim.x) mine += key_score(sbuf[i]) >= cut;
        if (mine) atomicAdd(&s_above, mine);
        __syncthreads();
        const int n_act = s_above;
        sel_stamp(a, q, 5);       // cut / coverage / prefix count
        rescore_candidates(sbuf, n_act, cut, f.db, f.d_pad, f.qn + (size_t)q * (size_t)f.qn_ld);
        const int ns2 = next_pow2(n_act > 2 ? n_act : 2);
        for (int i = n_act + tid; i < ns2; i += blockDim.x) sbuf[i] = 0ull;
        if (tid == 0) s_count = 0;
        __syncthreads();
        sel_stamp(a, q, 6);       // re-scored
        sort_desc_u64(sbuf, ns2);
        sel_stamp(a, q, 7);       // second sort done
        emit_topk(sbuf, ns2, f, q, &s_count, fs);
        sel_stamp(a, q, 8);
        if (a.trace && tid == 0) { a.trace[(size_t)q * 16 + 9] = (unsigned long long)C; a.trace[(size_t)q * 16 + 10] = (unsigned long long)n_act; a.trace[(size_t)q * 16 + 11] = (unsigned long long)nnz; }
        return;
    }

    if (a.out) {
        unsigned long long* out = a.out + (size_t)q * (size_t)a.out_ld;
        for (int i = tid; i < a.K; i += blockDim.x) out[i] = i < C ? sbuf[i] : 0ull;
    }
    // The k-th best of ANY subset of the DB is a lower bound of the k-th best of the whole DB, so
    // admitting `score >= tau` in the next scan level never drops a true top-k item.
    if (a.tau_out && tid == 0) {
        const float m = a.margin ? a.margin[q] : kBf16QueryMargin;
        float t = a.score_floor - m;
        if (a.tau_prev && a.tau_prev[q] > t) t = a.tau_prev[q];
        const unsigned long long key = (a.tau_k - 1) < C ? sbuf[a.tau_k - 1] : 0ull;
        if (key) {
            const float c = key_score(key) - m;
            if (c > t) t = c;
        }
        a.tau_out[q] = t;
        if (a.tau_hot_out) {
            const unsigned long long hk = (a.hot_rank >= 1 && a.hot_rank - 1 < C) ? sbuf[a.hot_rank - 1] : 0ull;
            a.tau_hot_out[q] = hk ? fmaxf(key_score(hk), t) : t;
        }
    }
}

// ---- exact top-k of a dense score row (small-Q path) -------------------------------------------------------------------
// One CTA of 1024 threads per query.  The fp32 scan of the Q <= 4 path leaves exact scores of the rows it visited (every
// row of a small shard, or a strided sample of a large one); their ordering keys are unique, so the k-th largest key is
// found by narrowing a key range with 2048-bin histograms until at most kDenseSort keys lie at or above the bound, which
// are then compacted and sorted.  With at most 16 columns per thread the keys stay in REGISTERS across the rounds (the
// reference's operating point, 10k rows: one read of 40 KB); otherwise every round re-reads the L2-resident row.
//   FINAL  emit ids / scores / counts (score_threshold walk, id_offset, peer push) — the whole answer of a small shard
//   TAU    tau_key_out[q] = k-th best key of the sample (0: fewer than k sampled rows, admit everything)
constexpr int kDenseThreads = 1024;
constexpr int kDenseSort = 2048;
constexpr int kDenseCache = 16;
constexpr int kDenseMaximaK = 128;

template <bool CACHED, bool FINAL>
__global__ void __launch_bounds__(kDenseThreads, 1) dense_topk_kernel(const DenseTopkArgs a, const FinalArgs f) {
    __shared__ int hist[kSelBins];
    __shared__ unsigned long long sbuf[kDenseSort];
    __shared__ unsigned long long s_red[2 * (kDenseThreads / 32)];
    __shared__ unsigned long long s_lo, s_hi;
    __shared__ int s_above, s_count, s_done, s_nnz, s_res[4];
    const int q = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    grid_dependency_wait();
    grid_launch_dependents();
    const float* src = a.dense + (size_t)q * (size_t)a.dense_ld;
    const long long pstride2 = a.pair_stride * 2;
    auto stamp = [&](int slot) {
        if (a.trace && tid == 0) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
            a.trace[(size_t)q * 16 + slot] = t;
        }
    };
    stamp(0);
    auto key_of = [&](long long c) -> unsigned long long {
        const long long row = (c >> 1) * pstride2 + (c & 1);
        return row < a.n_rows ? make_key(__ldg(src + c), (uint32_t)row) : 0ull;
    };
    unsigned long long kc[CACHED ? kDenseCache : 1];
    if (CACHED) {
#pragma unroll
        for (int i = 0; i < kDenseCache; ++i) {
            const long long c = (long long)tid + (long long)i * kDenseThreads;
            kc[i] = c < a.n_cols ? key_of(c) : 0ull;
        }
    }
    // visit every valid key: from registers (CACHED) or from the L2-resident row, twelve independent loads in flight per thread
    // (one CTA reads the whole row: without them a pass is pure L2 latency).  A macro, not a lambda: the register array must be
    // indexed by compile-time constants only or it is demoted to local memory.
#define RVO_FOR_EACH_KEY(KEY, BODY)                                                              \
    do {                                                                                         \
        if (CACHED) {                                                                            \
            _Pragma("unroll") for (int i_ = 0; i_ < kDenseCache; ++i_) {                         \
                const unsigned long long KEY = kc[CACHED ? i_ : 0];                              \
                if (KEY) { BODY; }                                                               \
            }                                                                                    \
        } else {                                                                                 \
            for (long long c0_ = tid; c0_ < a.n_cols; c0_ += 12 * kDenseThreads) {               \
                unsigned long long k12_[12];                                                     \
                _Pragma("unroll") for (int u_ = 0; u_ < 12; ++u_) {                              \
                    const long long c_ = c0_ + (long long)u_ * kDenseThreads;                    \
                    k12_[u_] = c_ < a.n_cols ? key_of(c_) : 0ull;                                \
                }                                                                                \
                _Pragma("unroll") for (int u_ = 0; u_ < 12; ++u_) {                              \
                    const unsigned long long KEY = k12_[u_];                                     \
                    if (KEY) { BODY; }                                                           \
                }                                                                                \
            }                                                                                    \
        }                                                                                        \
    } while (0)
    if (!FINAL && !CACHED && a.k <= kDenseMaximaK) {
        // threshold from a LARGE sample in one pass: the k-th largest of the 1024 per-thread maxima.  The maxima are distinct
        // elements of the sample, so their k-th largest is a lower bound of the sample's k-th largest key (hence of the
        // shard's): a valid filter threshold, ~5 % more survivors than the exact one, without the multi-pass refinement
        unsigned long long best = 0ull;
        RVO_FOR_EACH_KEY(k, best = k > best ? k : best);
        sbuf[tid] = best;
        __syncthreads();
        sort_desc_u64(sbuf, kDenseThreads);
        if (tid == 0) a.tau_key_out[q] = sbuf[a.k - 1];      // 0 (fewer than k non-empty threads): admit everything
        return;
    }
    // ---- fast path: bound from the per-thread maxima ------------------------------------------------------------------
    // The 1024 per-thread maxima are distinct keys, so the k-th largest of them is a lower bound of the k-th largest key
    // overall.  One histogram of the MAXIMA only (1024 shared-memory atomics spread over the upper tail — histogramming every
    // key piles the bulk of the scores onto a few bins and serialises) gives the lower edge of the bin holding it; the keys at
    // or above that edge (the top-k plus a handful) are compacted and sorted.  Exact; two passes over the keys.
    {
        unsigned long long best = 0ull;
        RVO_FOR_EACH_KEY(k, best = k > best ? k : best);
        for (int i = tid; i < kSelBins; i += kDenseThreads) hist[i] = 0;
        if (tid == 0) {
            s_count = 0;
            s_nnz = 0;
        }
        __syncthreads();
        // cosine scores of normalised rows and queries lie in [-1.01, 1.01]: a fixed key range, no min/max reduction
        const unsigned long long lo = (unsigned long long)f32_orderable(-1.01f) << 32;
        const unsigned long long hi = ((unsigned long long)f32_orderable(1.01f) << 32) | 0xFFFFFFFFull;
        const int shift = hist_shift(lo, hi);
        const unsigned have = __ballot_sync(0xFFFFFFFFu, best != 0ull);
        if (lane == 0 && have) atomicAdd(&s_nnz, __popc(have));
        if (best) {
            const unsigned long long b = best <= lo ? 0ull : (best >= hi ? (unsigned long long)(kSelBins - 1) : (best - lo) >> shift);
            atomicAdd(&hist[(int)b], 1);
        }
        __syncthreads();
        const int nthr = s_nnz;                       // threads that hold at least one key
        if (warp == 0) {
            const int need = a.k < nthr ? a.k : nthr;
            if (need > 0) hist_find(hist, need, lane, s_res);
            __syncwarp();
            if (lane == 0) s_lo = (need > 0 && a.k <= nthr && s_res[3] && s_res[0] > 0) ? lo + ((unsigned long long)s_res[0] << shift) : 0ull;
        }
        __syncthreads();
        const unsigned long long T0 = s_lo;           // 0: fewer than k threads hold keys (tiny shard) — keep everything
        stamp(1);
        RVO_FOR_EACH_KEY(k, if (k >= T0) {
            const int at = atomicAdd(&s_count, 1);
            if (at < kDenseSort) sbuf[at] = k;
        });
        __syncthreads();
        stamp(2);
        if (s_count <= kDenseSort) {
            const int C = s_count;
            const int ns = next_pow2(C > 2 ? C : 2);
            for (int i = C + tid; i < ns; i += kDenseThreads) sbuf[i] = 0ull;
            __syncthreads();
            sort_desc_u64(sbuf, ns);
            stamp(4);
            if (FINAL) {
                if (tid == 0) s_count = 0;
                __syncthreads();
                emit_topk(sbuf, C, f, q, &s_count);
                stamp(5);
                if (a.trace && tid == 0) a.trace[(size_t)q * 16 + 9] = (unsigned long long)C;
            } else if (tid == 0) {
                a.tau_key_out[q] = (a.k - 1) < C ? sbuf[a.k - 1] : 0ull;
            }
            return;
        }
        __syncthreads();                              // more than kDenseSort keys at or above the edge (tie mass): rigorous path
    }
    // ---- rigorous path: key range, then histogram refinement over every key ---------------------------------------------------
    unsigned long long lmin = ~0ull, lmax = 0ull;
    int lcnt = 0;
    RVO_FOR_EACH_KEY(k, lmin = k < lmin ? k : lmin; lmax = k > lmax ? k : lmax; ++lcnt);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long x = __shfl_xor_sync(0xFFFFFFFFu, lmin, o), y = __shfl_xor_sync(0xFFFFFFFFu, lmax, o);
        lmin = x < lmin ? x : lmin;
        lmax = y > lmax ? y : lmax;
        lcnt += __shfl_xor_sync(0xFFFFFFFFu, lcnt, o);
    }
    if (tid == 0) s_nnz = 0;
    __syncthreads();
    if (lane == 0) {
        s_red[warp] = lmin;
        s_red[kDenseThreads / 32 + warp] = lmax;
        if (lcnt) atomicAdd(&s_nnz, lcnt);
    }
    __syncthreads();
    if (tid == 0) {
        unsigned long long mn = ~0ull, mx = 0ull;
        for (int w = 0; w < kDenseThreads / 32; ++w) {
            mn = s_red[w] < mn ? s_red[w] : mn;
            mx = s_red[kDenseThreads / 32 + w] > mx ? s_red[kDenseThreads / 32 + w] : mx;
        }
        s_lo = mn;
        s_hi = mx;
        s_above = 0;
        s_done = s_nnz <= kDenseSort / 2 ? 1 : 0;     // few enough: keep every key
        if (s_done) s_lo = 0ull;
    }
    __syncthreads();
    const int nnz = s_nnz;
    const int want = a.k < nnz ? a.k : nnz;
    stamp(1);     // keys loaded, range known
    for (int round = 0; round < 16 && !s_done; ++round) {
        const unsigned long long lo = s_lo, hi = s_hi;
        const int shift = hist_shift(lo, hi);
        __syncthreads();
        for (int i = tid; i < kSelBins; i += kDenseThreads) hist[i] = 0;
        __syncthreads();
        RVO_FOR_EACH_KEY(k, if (k >= lo && k <= hi) atomicAdd(&hist[(int)((k - lo) >> shift)], 1));
        __syncthreads();
        if (warp == 0) {
            hist_find(hist, want - s_above, lane, s_res);
            __syncwarp();
            if (lane == 0) {
                if (!s_res[3]) {          // cannot happen (want <= nnz): take everything rather than loop
                    s_lo = 0ull;
                    s_done = 1;
                } else {
                    const int cge = s_above + s_res[1];      // keys >= lower bound of the crossing bin
                    s_lo = lo + ((unsigned long long)s_res[0] << shift);
                    if (cge <= kDenseSort / 2 || shift == 0) {
                        s_done = 1;
                    } else {
                        s_above = s_above + s_res[1] - s_res[2];
                        s_hi = s_lo + ((1ull << shift) - 1ull);
                    }
                }
            }
        }
        __syncthreads();
    }
    const unsigned long long T = s_lo;
    stamp(2);     // bound found
    if (tid == 0) s_count = 0;
    __syncthreads();
    RVO_FOR_EACH_KEY(k, if (k >= T) {
        const int at = atomicAdd(&s_count, 1);
        if (at < kDenseSort) sbuf[at] = k;
    });
    __syncthreads();
    stamp(3);     // compacted
    const int C = s_count < kDenseSort ? s_count : kDenseSort;
    const int ns = next_pow2(C > 2 ? C : 2);
    for (int i = C + tid; i < ns; i += kDenseThreads) sbuf[i] = 0ull;
    __syncthreads();
    sort_desc_u64(sbuf, ns);
    stamp(4);     // sorted
    if (FINAL) {
        if (tid == 0) s_count = 0;
        __syncthreads();
        emit_topk(sbuf, C, f, q, &s_count);
        stamp(5);
        if (a.trace && tid == 0) a.trace[(size_t)q * 16 + 9] = (unsigned long long)C;
    } else if (tid == 0) {
        a.tau_key_out[q] = (a.k - 1) < C ? sbuf[a.k - 1] : 0ull;
    }
}
#undef RVO_FOR_EACH_KEY

int launch_dense_topk(const DenseTopkArgs& a, const FinalArgs* f, int nq, cudaStream_t stream) {
    FinalArgs ff;
    if (f) ff = *f;
    else memset(&ff, 0, sizeof(ff));
    const bool cached = a.n_cols <= (long long)kDenseCache * kDenseThreads;
    if (f) {
        if (cached) RVO_CUDA(launch_pdl_small(dense_topk_kernel<true, true>, dim3(nq), dim3(kDenseThreads), 0, stream, a, ff));
        else RVO_CUDA(launch_pdl_small(dense_topk_kernel<false, true>, dim3(nq), dim3(kDenseThreads), 0, stream, a, ff));
    } else {
        if (cached) RVO_CUDA(launch_pdl_small(dense_topk_kernel<true, false>, dim3(nq), dim3(kDenseThreads), 0, stream, a, ff));
        else RVO_CUDA(launch_pdl_small(dense_topk_kernel<false, false>, dim3(nq), dim3(kDenseThreads), 0, stream, a, ff));
    }
    RVO_LAUNCHED();
    return RVO_OK;
}

// ---- seed threshold from per-thread maxima (see select.cuh) ----------------------------------------------------------
__global__ void __launch_bounds__(512) seed_tau_kernel(const float* __restrict__ dense, long long dense_ld, long long n_dense,
                                                       int nq, int k, const float* __restrict__ margin, float score_floor,
                                                       float* __restrict__ tau_out, int hot_rank, float* __restrict__ tau_hot_out,
                                                       unsigned long long* tr) {
    __shared__ uint32_t s[512];
    const int q = blockIdx.x, t = threadIdx.x;
    grid_dependency_wait();
    grid_launch_dependents();
    chain_stamp(tr, 2, false);
    if (q >= nq) {  // padded query rows of the tensor path never admit anything
        if (t == 0) {
            tau_out[q] = __int_as_float(0x7f800000);
            if (tau_hot_out) tau_hot_out[q] = __int_as_float(0x7f800000);
        }
        return;
    }
    const float NINF = -__int_as_float(0x7f800000);
    const float* src = dense + (size_t)q * (size_t)dense_ld;
    float mx[8];                                                   // eight independent chains: loads stay in flight
#pragma unroll
    for (int u = 0; u < 8; ++u) mx[u] = NINF;
    long long i = t;
    for (; i + 7 * 512 < n_dense; i += 8 * 512) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = __ldg(src + i + u * 512);
#pragma unroll
        for (int u = 0; u < 8; ++u) mx[u] = fmaxf(mx[u], v[u]);
    }
    for (; i < n_dense; i += 512) mx[0] = fmaxf(mx[0], __ldg(src + i));
    // fmaxf drops NaNs; -inf (no element) orders lowest
    uint32_t x = f32_orderable(fmaxf(fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])), fmaxf(fmaxf(mx[4], mx[5]), fmaxf(mx[6], mx[7]))));
    // bitonic sort of the 512 maxima, descending, one key per thread: shuffles inside a warp, shared memory across warps
    // (a 2048-bin histogram + bin search was measured SLOWER: 12.5 vs 8.4 us)
    for (int kk = 2; kk <= 512; kk <<= 1) {
        const bool desc = (t & kk) == 0;
        for (int j = kk >> 1; j > 0; j >>= 1) {
            uint32_t y;
            if (j >= 32) {
                __syncthreads();
                s[t] = x;
                __syncthreads();
                y = s[t ^ j];
            } else {
                y = __shfl_xor_sync(0xFFFFFFFFu, x, j);
            }
            const bool take_max = ((t & j) == 0) == desc;
            x = take_max ? (x > y ? x : y) : (x < y ? x : y);
        }
    }
    __syncthreads();
    s[t] = x;
    __syncthreads();
    if (t == 0) {
        const float mg = margin ? margin[q] : kBf16QueryMargin;
        float tau = score_floor - mg;                              // -inf stays -inf
        const float kth = orderable_f32(s[k - 1]);
        if (kth > NINF && kth - mg > tau) tau = kth - mg;          // fewer than k maxima exist: keep the floor (admit everything)
        tau_out[q] = tau;
        if (tau_hot_out) {
            // a PROPOSAL for "the few hundred best rows": the hot_rank-th largest maximum, never below the safe threshold
            const float rth = (hot_rank >= 1 && hot_rank <= 512) ? orderable_f32(s[hot_rank - 1]) : NINF;
            tau_hot_out[q] = rth > tau ? rth : tau;
        }
    }
    chain_stamp(tr, 2, true);
}

int launch_seed_tau(const float* dense, long long dense_ld, long long n_dense, int nq, int grid_q, int k, const float* margin,
                    float score_floor, float* tau_out, cudaStream_t stream, int hot_rank, float* tau_hot_out) {
    if (grid_q <= 0) return RVO_OK;
    RVO_CUDA(launch_pdl(seed_tau_kernel, dim3(grid_q), dim3(512), 0, stream, dense, dense_ld, n_dense, nq, k, margin, score_floor,
                        tau_out, hot_rank, tau_hot_out, (unsigned long long*)(uintptr_t)g_chain_trace.load()));
    RVO_LAUNCHED();
    return RVO_OK;
}

int launch_select(const SelectArgs& a, int grid_q, cudaStream_t stream) {
    if (grid_q <= 0) return RVO_OK;
    FinalArgs f;
    memset(&f, 0, sizeof(f));
    RVO_CUDA(launch_pdl(select_kernel<false>, dim3(grid_q), dim3(kSelThreads), 0, stream, a, f));
    RVO_LAUNCHED();
    return RVO_OK;
}

int launch_select_final(const SelectArgs& a, const FinalArgs& f, int nq, cudaStream_t stream) {
    if (nq <= 0) return RVO_OK;
    RVO_CUDA(launch_pdl(select_kernel<true>, dim3(nq), dim3(kSelThreads), 0, stream, a, f));
    RVO_LAUNCHED();
    return RVO_OK;
}

// ---- K3: merge of per-shard lists -------------------------------------------------------------
constexpr int kMaxMergeLists = 64;
std::atomic<long long> g_exchange_timeout_ms{60000};   // option "exchange_timeout_ms"
std::atomic<long long> g_merge_trace{0};               // option "merge_trace": device buffer [3] u64 (start, flags seen, end)
__device__ __forceinline__ bool before(float sa, long long ia, float sb, long long ib) {
    return sa > sb || (sa == sb && ia < ib);
}

// per-shard blocks are `*_gs` elements apart (contiguous [G,nq,k] arrays, the packed all-gather buffer or the exchange region).
// Every shard list arrives sorted by (score desc, id asc) and global ids are unique, so the merged position of an element is
// its own index plus, for every other list, the number of that list's elements ranked before it (one binary search each):
// no sorting network, two block barriers.
__global__ void __launch_bounds__(256) merge_kernel(const int64_t* ids, const float* scores, const int32_t* counts,
                                                    long long ids_gs, long long scores_gs, long long counts_gs,
                                                    int G, int k, int64_t* out_ids,
                                                    float* out_scores, int32_t* out_counts,
                                                    const unsigned long long* wait_flags, unsigned long long wait_epoch,
                                                    unsigned long long wait_timeout_ns, unsigned long long* tr) {
    extern __shared__ unsigned char sm[];
    __shared__ int s_timeout;
    if (tr && blockIdx.x == 0 && threadIdx.x == 0) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tr[0]));
    if (wait_flags) {
        // peer-memory exchange: rank g's push of this search has landed in the local region once its flag shows the epoch
        // A late peer (host-side GC, lazy module load, ingest) only delays this rank, as a collective would.  After
        // `wait_timeout_ns` (option "exchange_timeout_ms", default 60 s) the kernel gives up WITHOUT trapping: every query
        // of the batch reports count -2 and the host raises / falls back to the all-gather path; the context stays usable.
        if (threadIdx.x == 0) s_timeout = 0;
        __syncthreads();
        if (threadIdx.x < G) {
            unsigned long long t0, t1;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
            while (ld_acquire_sys_u64(wait_flags + threadIdx.x) < wait_epoch) {
                __nanosleep(200);
                asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
                if (t1 - t0 > wait_timeout_ns) {
                    s_timeout = 1;
                    break;
                }
            }
        }
        __syncthreads();
        if (tr && blockIdx.x == 0 && threadIdx.x == 0) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tr[1]));
        if (s_timeout) {
            for (int i = threadIdx.x; i < k; i += blockDim.x) {
                out_ids[(size_t)blockIdx.x * k + i] = -1;
                out_scores[(size_t)blockIdx.x * k + i] = -__int_as_float(0x7f800000);
            }
            if (threadIdx.x == 0) out_counts[blockIdx.x] = -2;
            return;
        }
    }
    long long* sid = (long long*)sm;                 // [G][k]
    float* ssc = (float*)(sid + (size_t)G * k);      // [G][k]
    __shared__ int s_cnt[kMaxMergeLists], s_bad;
    const int q = blockIdx.x;
    if (threadIdx.x == 0) s_bad = 0;
    __syncthreads();
    if (threadIdx.x < G) {
        const int c = __ldcg(counts + (size_t)threadIdx.x * counts_gs + q);
        if (c < 0) s_bad = 1;
        s_cnt[threadIdx.x] = c < 0 ? 0 : (c > k ? k : c);
    }
    __syncthreads();
    const int n = G * k;
    for (int e = threadIdx.x; e < n; e += blockDim.x) {
        const int g = e / k, i = e - g * k;
        if (i < s_cnt[g]) {
            ssc[e] = __ldcg(scores + (size_t)g * scores_gs + (size_t)q * k + i);
            sid[e] = __ldcg(ids + (size_t)g * ids_gs + (size_t)q * k + i);
        }
    }
    __syncthreads();
    int total = 0;
    for (int g = 0; g < G; ++g) total += s_cnt[g];
    const int n_out = s_bad ? 0 : (total < k ? total : k);
    for (int e = threadIdx.x; e < n; e += blockDim.x) {
        const int g = e / k, i = e - g * k;
        if (i >= s_cnt[g] || s_bad) continue;
        const float sc = ssc[e];
        const long long id = sid[e];
        int pos = i;
        for (int h = 0; h < G && pos < k; ++h) {
            if (h == g) continue;
            // number of elements of list h ranked before (sc, id): first index whose element is NOT before it
            int lo = 0, hi = s_cnt[h];
            const float* hs = ssc + h * k;
            const long long* hi_ids = sid + h * k;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (before(hs[mid], hi_ids[mid], sc, id)) lo = mid + 1;
                else hi = mid;
            }
            pos += lo;
        }
        if (pos < k) {
            out_ids[(size_t)q * k + pos] = id;
            out_scores[(size_t)q * k + pos] = sc;
        }
    }
    for (int i = n_out + threadIdx.x; i < k; i += blockDim.x) {
        out_ids[(size_t)q * k + i] = -1;
        out_scores[(size_t)q * k + i] = -__int_as_float(0x7f800000);
    }
    if (threadIdx.x == 0) out_counts[q] = s_bad ? -1 : n_out;
    if (tr && blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tr[2]));
}

// ---- host wrappers ------------------------------------------------------------------------------
int launch_merge(const int64_t* ids, const float* scores, const int32_t* counts, long long ids_gs, long long scores_gs,
                 long long counts_gs, int G, int nq, int k, int64_t* out_ids, float* out_scores, int32_t* out_counts,
                 cudaStream_t stream, const unsigned long long* wait_flags, unsigned long long wait_epoch) {
    if (G > kMaxMergeLists) {
        set_error("merge: at most %d lists (G=%d)", kMaxMergeLists, G);
        return RVO_E_INVALID;
    }
    const size_t smem = (size_t)G * k * 12;
    if (smem > 47 * 1024)   // the 48 KB default covers dynamic + static shared memory (the kernel has ~300 B of the latter)
        RVO_CUDA(cudaFuncSetAttribute(merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    merge_kernel<<<nq, 256, smem, stream>>>(ids, scores, counts, ids_gs, scores_gs, counts_gs, G, k, out_ids, out_scores,
                                            out_counts, wait_flags, wait_epoch,
                                            (unsigned long long)g_exchange_timeout_ms.load() * 1000000ull,
                                            (unsigned long long*)(uintptr_t)g_merge_trace.load());
    RVO_LAUNCHED();
    return RVO_OK;
}

}  // namespace rvo
