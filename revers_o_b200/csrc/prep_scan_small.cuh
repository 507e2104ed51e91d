// Declarations for prep_scan_small.cu and mask_pool.cu launchers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rvo {

int launch_normalize_rows(const float* src, long long n, int d, long long src_ld, uint16_t* dst_bf16, long long dst_ld,
                          long long tiled_row0, float* dst_f32, long long f32_ld, cudaStream_t stream,
                          float* margin_out = nullptr,   // row-major bf16 output only: per-row admission margin (common.cuh)
                          // rows n .. n_pad_rows-1 of the row-major outputs are written as zeros (query operand padding), and
                          // zero_bytes (a multiple of 16) at zero_base (16-byte aligned) are cleared by the same launch
                          long long n_pad_rows = 0, void* zero_base = nullptr, size_t zero_bytes = 0);

// fp32 CUDA-core scan for batches of <= RVO_SMALL_Q queries (prep_scan_small.cu)
struct SmallScanArgs {
    const uint4* db;
    long long n_rows;
    int nk, nchunks, nq, d;
    const float* q;            // NORM: raw queries [nq][q_ld]; else normalised, zero padded [NQ][q_ld]
    long long q_ld;
    float* qn_out;             // NORM: block 0 writes the normalised queries here, [RVO_SMALL_Q][qn_ld] zero padded
    long long qn_ld;
    // DENSE
    float* out;
    long long out_ld, pair_stride;
    uint4* zero_base;          // NORM: zero_u4 16-byte words cleared by this launch (the FILTER pass's counters), or null
    long long zero_u4;
    // FILTER
    const unsigned long long* tau_key;
    unsigned long long* cand;
    int* cnt;
    int nseg, cap;
};

bool scan_small_supports(int d_pad);
// dense: scores of the sampled rows + min/max; !dense: FILTER.  norm (dense only): the kernel normalises the raw queries itself
int launch_scan_small(const SmallScanArgs& a, bool dense, bool norm, int sm_count, cudaStream_t stream);

size_t mask_pool_workspace_bytes(int B, int M, int P, int D);
int launch_mask_pool(const uint16_t* feats, const uint8_t* masks, int B, int M, int P, int D, int max_regions, float* out,
                     int32_t* out_counts, int32_t* out_src, int32_t* out_total, void* workspace, size_t workspace_bytes,
                     int sm_count, cudaStream_t stream, uint16_t* db = nullptr, long long db_row0 = 0, int feat_f16 = 0);

}  // namespace rvo
