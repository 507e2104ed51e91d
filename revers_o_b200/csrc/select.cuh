// Declarations of the selection kernels (select.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rvo {

constexpr int kSelThreads = 512;

// exact top-k of dense fp32 scores (small-Q path, select.cu dense_topk_kernel)
struct DenseTopkArgs {
    const float* dense;                 // [nq][dense_ld] scores of the visited rows
    long long dense_ld, n_cols;         // n_cols = 2 * visited row pairs
    long long n_rows, pair_stride;      // column c is DB row (c >> 1) * 2 * pair_stride + (c & 1); valid iff < n_rows
    int k;
    unsigned long long* tau_key_out;    // threshold mode: k-th best ordering key of the sample per query
    unsigned long long* trace;          // option "select_trace": [nq][16] globaltimer stamps of the kernel's phases (debug)
};

constexpr int kSelBins = 2048;    // histogram bins per refinement round
constexpr int kSelSort = 2048;    // keys compacted and sorted in shared memory

struct SelectArgs {
    const float* dense;                 // dense mode: scores [nq][dense_ld], n_dense valid per query
    long long dense_ld, n_dense;
    const unsigned long long* keys;     // list mode: keys [nq][nseg][cap]; sub-list s of query q holds
    const int* cnt;                     //   min(cnt[q*nseg + s], cap) valid keys
    int nseg, cap;
    int K;                              // top keys wanted (<= 1024)
    unsigned long long* out;            // [nq][out_ld] top-K keys, sorted descending, zero padded (or nullptr)
    long long out_ld;
    float* tau_out;                     // optional: tau_out[q] = max(tau_prev[q], score(key[tau_k-1]), score_floor) - margin[q]
    const float* tau_prev;              //   (tau_prev is already lowered by the margin)
    int tau_k;
    const float* margin;                // per-query admission margin (common.cuh query_margin), [nq]
    float score_floor;                  // the caller's score_threshold (-inf when there is none)
    int nq;                             // CTAs q >= nq only write tau_out[q] = +inf (padded query rows)
    const float* range_lo;              // optional per-query lower bound of every key's score (finer first bins)
    // hot lists (scan_tc.cuh kHotSplit).  Threshold levels (select_kernel<false>, seed_tau) also emit
    //   tau_hot_out[q] = max(score of the hot_rank-th best key of the sample, tau_out[q])     (no margin: a proposal only)
    // and the final level (select_kernel<true>) first tries the sub-lists [hot_seg0, hot_seg0 + hot_nseg) alone: they hold
    // every candidate with tensor score >= tau_hot[q], which covers the top-k whenever (k-th best hot score) - margin >= tau_hot.
    float* tau_hot_out;
    int hot_rank;
    const float* tau_hot;
    int hot_seg0, hot_nseg;
    unsigned long long* trace;          // option "select_trace": [nq][16] globaltimer stamps of the kernel's phases (debug)
};

// Peer-memory exchange fused into the last kernel of the sharded search: besides its own result arrays the
// CTA of query q stores its k (id, score) pairs and its count into the slot of THIS rank in every peer's exchange
// region (NVLink stores), and the last CTA of the grid to finish publishes `epoch` in every peer's flag for this rank.
constexpr int kMaxPeers = 8;
struct PushArgs {
    int world, rank;                           // world <= 1: no exchange
    int nq, k;                                 // blob layout [ids int64 nq*k | scores f32 nq*k | counts i32 nq]
    unsigned long long epoch;
    unsigned char* peer_slot[kMaxPeers];       // this rank's slot (current parity) in peer g's region; [rank] unused
    unsigned long long* peer_flag[kMaxPeers];  // this rank's flag (current parity) in peer g's region; [rank] = local
    unsigned int* done;                        // local CTA counter of the current parity (zero between uses)
    // fused exchange + merge (rvo_search_topk_fused): the CTA of query q stores its list as ONE record into every rank's region,
    // publishes a per-(rank, query) flag, waits for the peers' flags of the SAME query and merges the `world` lists itself —
    // no grid-wide counter, no separate merge launch.
    int fused;
    int nq_max;
    unsigned int rec_bytes;                    // bytes of one (rank, query) record: [k ids int64 | k scores f32 | count i32], 16-aligned
    unsigned char* lists[kMaxPeers];           // record area of the current slot set in rank g's region ([g] == rank: local)
    unsigned long long* qflags[kMaxPeers];     // per-(source rank, query) epoch flags of the current slot set in rank g's region
    unsigned long long timeout_ns;
};

struct FinalArgs {
    const unsigned long long* top;      // [nq][top_ld], first K2 sorted descending
    long long top_ld;
    const int* cnt;                     // raw candidate counts [nq][nseg] (overflow detection) or nullptr
    int nseg, cap, K2, k;
    float score_threshold;
    const float* margin;                // per-query admission margin (nullptr: 0)
    int rescore;                        // 1: keys carry bf16-query tensor scores -> fp32 re-score
    const uint16_t* db;                 // tiled DB storage (common.cuh)
    int d_pad;
    const float* qn;                    // normalised fp32 queries [nq][qn_ld], zero padded to d_pad
    long long qn_ld;
    long long id_offset;
    int64_t* out_ids;
    float* out_scores;
    int32_t* out_counts;
    PushArgs push;
};

// f != nullptr: emit the final answer (ids / scores / counts, threshold walk); f == nullptr: write a.tau_key_out
struct FinalArgs;
int launch_dense_topk(const DenseTopkArgs& a, const FinalArgs* f, int nq, cudaStream_t stream);
int launch_select(const SelectArgs& a, int grid_q, cudaStream_t stream);
// Seed level only (dense sample, k <= kSeedTauMaxK): tau_out[q] = max(k-th largest of the 512 per-thread maxima of the sample,
// score_floor) - margin[q].  The per-thread maxima are distinct elements of the sample, so their k-th largest is a lower bound of
// the sample's k-th largest (hence of the DB's): a valid threshold from ONE pass and one 512-key sort.
constexpr int kSeedTauMaxK = 128;
int launch_seed_tau(const float* dense, long long dense_ld, long long n_dense, int nq, int grid_q, int k, const float* margin,
                    float score_floor, float* tau_out, cudaStream_t stream, int hot_rank = 0, float* tau_hot_out = nullptr);
// last level fused with the fp32 re-score and the final ordering: `a` selects (a.K = candidates aimed at, a.out unused),
// `f` supplies k, score_threshold, margin, db/qn and the outputs (f.top / f.cnt / f.K2 unused)
int launch_select_final(const SelectArgs& a, const FinalArgs& f, int nq, cudaStream_t stream);
// wait_flags != nullptr: every CTA first waits until wait_flags[g] >= wait_epoch for all g < G (peer pushes landed)
int launch_merge(const int64_t* ids, const float* scores, const int32_t* counts, long long ids_gs, long long scores_gs,
                 long long counts_gs, int G, int nq, int k, int64_t* out_ids, float* out_scores, int32_t* out_counts,
                 cudaStream_t stream, const unsigned long long* wait_flags = nullptr, unsigned long long wait_epoch = 0);

}  // namespace rvo
