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

int launch_scan_small(const uint16_t* db, long long n_rows, int d_pad, const float* qn, long long qn_ld,
                      int nq, float* out, long long out_ld, int sm_count, cudaStream_t stream);

size_t mask_pool_workspace_bytes(int B, int M, int P, int D);
int launch_mask_pool(const uint16_t* feats, const uint8_t* masks, int B, int M, int P, int D, int max_regions, float* out,
                     int32_t* out_counts, int32_t* out_src, int32_t* out_total, void* workspace, size_t workspace_bytes,
                     int sm_count, cudaStream_t stream, uint16_t* db = nullptr, long long db_row0 = 0, int feat_f16 = 0);

}  // namespace rvo
