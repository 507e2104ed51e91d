// Declarations shared between the tcgen05 scan kernel (scan_tc.cu) and the search driver (api.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

namespace rvo {

constexpr int kBlockM = 128;                              // DB rows per sub-tile == TMEM lanes
constexpr int kBlockK = 64;                               // bf16 per k-chunk == one 128-B swizzle span
constexpr int kSubTileBytes = kBlockM * kBlockK * 2;      // 16 KiB
constexpr int kMaxStages = 8;
constexpr int kMaxSlots = 8;
constexpr int kEpiWarps = 8;                              // two epilogue warps per TMEM lane quarter
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kScanThreads = 64 + kEpiThreads;            // TMA warp + MMA warp + epilogue warps
constexpr int kQueueCap = 12;                             // smem survivor queue entries per epilogue thread
constexpr int kSmemLimit = 232448;                        // 227 KiB opt-in dynamic smem per CTA
constexpr int kCandSplit = 16;                            // sub-lists per query: spreads the append atomics over
                                                          // 16x more L2 addresses (same-address atomics serialise)
constexpr int kHotSplit = 4;                              // extra "hot" sub-lists per query (search only): survivors whose score
                                                          // also reaches tau_hot (a proposal for "the few hundred best") go here,
                                                          // so the final select reads ~600 keys instead of every survivor
constexpr int kCandSegs = kCandSplit + kHotSplit;         // sub-lists per query of the search path ([safe 0..15 | hot 16..19])
constexpr int kTauSmemBytes = 4 * 256 * 4;                // [2][256] tau + [2][256] tau_hot, double-buffered by work item
constexpr int kModeDense = 0;
constexpr int kModeFilter = 1;
constexpr int kModeDenseMax = 2;   // launch-time variant of DENSE: only the maximum of every 32 consecutive sample rows is written
                                   // (dense[q][sample / 32]) — all the seed threshold needs (select.cu seed_tau_kernel: the k-th
                                   // largest of maxima over DISJOINT row groups is a lower bound of the k-th best score), 32x less
                                   // output to write and to read back

struct ScanParams {
    long long n_rows;        // DB rows
    long long super_stride;  // tile sampling: only every super_stride-th super-tile is visited (1 = all)
    long long num_super;     // visited super-tiles of 128*m_sub rows
    int d_pad, nq_blk, num_qblk, m_sub, num_stages, resident_q, slot_w, num_slots;
    uint32_t stage_bytes, off_stages, off_queue, off_tau, off_bars;
    // FILTER
    const float* tau;              // [nq_pad] per-query admission thresholds
    const float* tau_hot;          // [nq_pad] or nullptr: survivors with score >= tau_hot[q] go to a hot sub-list
    unsigned long long* cand;      // [nq_pad][nseg][cap] candidate ordering keys
    int* cand_cnt;                 // [nq_pad][nseg]; super-tile st appends to sub-list st % kCandSplit (hot: kCandSplit + st % kHotSplit)
    int nseg;                      // sub-lists per query in cand / cand_cnt (kCandSplit, or kCandSegs with hot lists)
    int cap;                       // capacity of ONE sub-list
    // DENSE
    float* dense;                  // [nq_pad][dense_ld]
    long long dense_ld;
    int dense_max;                 // 1: kModeDenseMax
    unsigned long long* trace;     // option "chain_trace" (common.cuh) or nullptr
};

#ifdef __CUDACC__
// Column maxima over the 32 rows of a warp.  Every lane holds 16 column values of its own row; after the butterfly (16 shuffles)
// lane L holds, for column (L >> 1) & 15, the maximum over all 32 lanes (the two lanes of a pair hold the same value).
__device__ __forceinline__ float warp_colmax16(const float (&v)[16], int lane) {
    float a[8], b[4], c[2];
    const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4, h2 = lane & 2;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float keep = h16 ? v[i + 8] : v[i], send = h16 ? v[i] : v[i + 8];
        a[i] = fmaxf(keep, __shfl_xor_sync(0xFFFFFFFFu, send, 16));
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float keep = h8 ? a[i + 4] : a[i], send = h8 ? a[i] : a[i + 4];
        b[i] = fmaxf(keep, __shfl_xor_sync(0xFFFFFFFFu, send, 8));
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float keep = h4 ? b[i + 2] : b[i], send = h4 ? b[i] : b[i + 2];
        c[i] = fmaxf(keep, __shfl_xor_sync(0xFFFFFFFFu, send, 4));
    }
    const float keep = h2 ? c[1] : c[0], send = h2 ? c[0] : c[1];
    const float d = fmaxf(keep, __shfl_xor_sync(0xFFFFFFFFu, send, 2));
    return fmaxf(d, __shfl_xor_sync(0xFFFFFFFFu, d, 1));
}
#endif

struct TcPlan {
    int nq_blk, num_qblk, nq_pad, m_sub, num_stages, resident, slot_w, num_slots;
    int pair;   // 1: CTA-pair kernel (cta_group::2, scan_tc2.cu)
    size_t stage_bytes, off_stages, off_queue, off_tau, off_bars, smem_bytes;
};

int plan_scan_tc(int nq, int d_pad, int force_m_sub, TcPlan* pl);   // force_m_sub: 0 auto, 1/2 single-CTA kernel
int plan_scan_tc2(const TcPlan& base, int d_pad, TcPlan* pl);
int launch_scan_tc2(int mode, const uint16_t* db, long long n_rows, long long super_stride, int d_pad,
                    const uint16_t* q_bf16, const TcPlan& pl, const float* tau, unsigned long long* cand, int* cand_cnt,
                    int cap, float* dense, long long dense_ld, int sm_count, cudaStream_t stream, const float* tau_hot,
                    int nseg);
// columns of the DENSE output / rows visited when only every super_stride-th super-tile is scanned
long long scan_tc_sample_rows(long long n_rows, const TcPlan& pl, long long super_stride);
// tau_hot / nseg: see ScanParams (defaults: no hot lists, kCandSplit sub-lists per query)
int launch_scan_tc(int mode, const uint16_t* db, long long n_rows, long long super_stride, int d_pad,
                   const uint16_t* q_bf16, const TcPlan& pl, const float* tau, unsigned long long* cand,
                   int* cand_cnt, int cap, float* dense, long long dense_ld, int sm_count, cudaStream_t stream,
                   const float* tau_hot = nullptr, int nseg = kCandSplit);

}  // namespace rvo
