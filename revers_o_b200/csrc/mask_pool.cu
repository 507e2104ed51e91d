// K1 — segmented mask pooling + L2 normalise.
// Replaces the per-region python loop of core_system.py:363-408 with the mask-pooled definition the
// reference states (main.py:8-9): e_m = mean_{p in mask m} F[p,:], e_m /= ||e_m||; empty masks are
// dropped and later regions shift up (core_system.py:402-404); only the first `max_regions` regions
// are visited (core_system.py:363).
//
// Pipeline (all on one stream, no host round trip):
//   mask_index_kernel    masks u8 [B,M,P] -> per-region compact patch lists (u16) + areas
//   mask_offsets_kernel  exclusive scan of (area > 0) -> output row of every region, per-image counts
//   mask_pool_kernel     CTA = (image, 32-channel slab): the slab of the patch-feature map is staged once
//                        in shared memory as fp32 (128-bit coalesced loads), each warp walks the patch
//                        list of one region, 8 lanes x float4 per patch row, 4 rows in flight per step;
//                        warp-shuffle reduction; writes the un-normalised mean and accumulates ||e||^2
//   mask_scale_kernel    e /= sqrt(||e||^2)   (rows just written are L2-resident)
// Work is proportional to sum of mask areas (segmented), not to M*P.
#include <cuda_fp16.h>

#include "common.cuh"
#include "prep_scan_small.cuh"

namespace rvo {

constexpr int kPoolThreads = 512;
constexpr int kPoolWarps = kPoolThreads / 32;
constexpr int kSlab = 32;  // channels per CTA
void* g_pool_trace = nullptr;          // option "pool_trace": device buffer [B][8] u64 for per-image pipeline time stamps (debug)
bool g_force_cuda_core_pool = false;  // option "pool_path" = 1: always use the CUDA-core kernels below

__global__ void __launch_bounds__(256) mask_index_kernel(const uint8_t* __restrict__ masks, int BM, int M, int P, int lim,
                                                         uint16_t* __restrict__ idx, int* __restrict__ area) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bm = blockIdx.x * 8 + warp;
    if (bm >= BM) return;
    const int m = bm % M;
    const uint8_t* row = masks + (size_t)bm * P;
    uint16_t* dst = idx + (size_t)bm * P;
    int cnt = 0;
    if (m < lim) {
        if ((P & 3) == 0) {
            const uint32_t* row4 = (const uint32_t*)row;
            for (int w0 = 0; w0 < (P >> 2); w0 += 32) {
                const int w = w0 + lane;
                const uint32_t v = w < (P >> 2) ? __ldg(row4 + w) : 0u;
                const int b0 = (v & 0xFFu) != 0, b1 = (v & 0xFF00u) != 0, b2 = (v & 0xFF0000u) != 0, b3 = (v >> 24) != 0;
                const int c = b0 + b1 + b2 + b3;
                int pre = c;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(0xFFFFFFFFu, pre, o);
                    if (lane >= o) pre += t;
                }
                const int total = __shfl_sync(0xFFFFFFFFu, pre, 31);
                int at = cnt + pre - c;
                if (b0) dst[at++] = (uint16_t)(4 * w);
                if (b1) dst[at++] = (uint16_t)(4 * w + 1);
                if (b2) dst[at++] = (uint16_t)(4 * w + 2);
                if (b3) dst[at++] = (uint16_t)(4 * w + 3);
                cnt += total;
            }
        } else {
            for (int p0 = 0; p0 < P; p0 += 32) {
                const int p = p0 + lane;
                const bool in = p < P && row[p] != 0;
                const unsigned bal = __ballot_sync(0xFFFFFFFFu, in);
                if (in) dst[cnt + __popc(bal & ((1u << lane) - 1u))] = (uint16_t)p;
                cnt += __popc(bal);
            }
        }
    }
    if (lane == 0) area[bm] = cnt;
}

__global__ void __launch_bounds__(1024) mask_offsets_kernel(const int* __restrict__ area, int B, int M, int* __restrict__ out_row,
                                                            int* __restrict__ out_counts, int* __restrict__ out_src,
                                                            int* __restrict__ out_total, float* __restrict__ sumsq) {
    __shared__ int s_scan[1024];
    const int n = B * M;
    const int per = (n + 1023) / 1024;
    const int lo = threadIdx.x * per;
    const int hi = lo + per < n ? lo + per : n;
    for (int b = threadIdx.x; b < B; b += 1024) out_counts[b] = 0;
    int c = 0;
    for (int i = lo; i < hi; ++i) {
        c += area[i] > 0;
        sumsq[i] = 0.f;
    }
    s_scan[threadIdx.x] = c;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const int t = threadIdx.x >= o ? s_scan[threadIdx.x - o] : 0;
        __syncthreads();
        s_scan[threadIdx.x] += t;
        __syncthreads();
    }
    int at = s_scan[threadIdx.x] - c;
    for (int i = lo; i < hi; ++i) {
        if (area[i] > 0) {
            out_row[i] = at;
            if (out_src) out_src[at] = i;
            atomicAdd(out_counts + i / M, 1);
            ++at;
        } else {
            out_row[i] = -1;
        }
    }
    if (threadIdx.x == 1023) *out_total = s_scan[1023];
}

template <bool F16>
__global__ void __launch_bounds__(kPoolThreads) mask_pool_kernel(const uint16_t* __restrict__ feats, int M, int P, int D, int lim,
                                                                 int p_pad, const uint16_t* __restrict__ idx,
                                                                 const int* __restrict__ area, const int* __restrict__ out_row,
                                                                 float* __restrict__ out, float* __restrict__ sumsq) {
    extern __shared__ __align__(16) unsigned char smem[];
    float* sf = (float*)smem;                                   // [P][32] fp32 slab
    uint16_t* sl_all = (uint16_t*)(sf + (size_t)P * kSlab);     // [warps][p_pad] patch list staging
    const int slab = blockIdx.x, b = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    {   // stage the slab: 4 threads x 16 B per patch row, converted to fp32 once
        const int seg = threadIdx.x & 3;
        const uint16_t* base = feats + ((size_t)b * P) * D + (size_t)slab * kSlab + seg * 8;
        for (int r = threadIdx.x >> 2; r < P; r += kPoolThreads / 4) {
            const uint4 v = __ldcs((const uint4*)(base + (size_t)r * D));
            float4 lo4, hi4;
            if constexpr (F16) {
                const float2 a = __half22float2(*(const __half2*)&v.x), b2 = __half22float2(*(const __half2*)&v.y);
                const float2 c2 = __half22float2(*(const __half2*)&v.z), d2 = __half22float2(*(const __half2*)&v.w);
                lo4 = make_float4(a.x, a.y, b2.x, b2.y);
                hi4 = make_float4(c2.x, c2.y, d2.x, d2.y);
            } else {
                lo4.x = __uint_as_float(v.x << 16); lo4.y = __uint_as_float(v.x & 0xFFFF0000u);
                lo4.z = __uint_as_float(v.y << 16); lo4.w = __uint_as_float(v.y & 0xFFFF0000u);
                hi4.x = __uint_as_float(v.z << 16); hi4.y = __uint_as_float(v.z & 0xFFFF0000u);
                hi4.z = __uint_as_float(v.w << 16); hi4.w = __uint_as_float(v.w & 0xFFFF0000u);
            }
            float4* d4 = (float4*)(sf + (size_t)r * kSlab + seg * 8);
            d4[0] = lo4;
            d4[1] = hi4;
        }
    }
    __syncthreads();

    uint16_t* sl = sl_all + (size_t)warp * p_pad;
    const int g = lane >> 3, l8 = lane & 7;
    const float4* sf4 = (const float4*)sf;  // row stride = 8 float4
    for (int m = warp; m < lim; m += kPoolWarps) {
        const int bm = b * M + m;
        const int a = area[bm];
        if (a == 0) continue;
        const int orow = out_row[bm];
        const uint16_t* gl = idx + (size_t)bm * P;
        __syncwarp();
        for (int e = lane; e < a; e += 32) sl[e] = gl[e];
        __syncwarp();
        float4 acc0 = make_float4(0.f, 0.f, 0.f, 0.f), acc1 = acc0;
        int e = g;
        for (; e + 12 < a; e += 16) {
            const int p0 = sl[e], p1 = sl[e + 4], p2 = sl[e + 8], p3 = sl[e + 12];
            const float4 v0 = sf4[p0 * 8 + l8], v1 = sf4[p1 * 8 + l8], v2 = sf4[p2 * 8 + l8], v3 = sf4[p3 * 8 + l8];
            acc0.x += v0.x; acc0.y += v0.y; acc0.z += v0.z; acc0.w += v0.w;
            acc1.x += v1.x; acc1.y += v1.y; acc1.z += v1.z; acc1.w += v1.w;
            acc0.x += v2.x; acc0.y += v2.y; acc0.z += v2.z; acc0.w += v2.w;
            acc1.x += v3.x; acc1.y += v3.y; acc1.z += v3.z; acc1.w += v3.w;
        }
        for (; e < a; e += 4) {
            const float4 v0 = sf4[(int)sl[e] * 8 + l8];
            acc0.x += v0.x; acc0.y += v0.y; acc0.z += v0.z; acc0.w += v0.w;
        }
        acc0.x += acc1.x; acc0.y += acc1.y; acc0.z += acc1.z; acc0.w += acc1.w;
#pragma unroll
        for (int o = 8; o <= 16; o <<= 1) {
            acc0.x += __shfl_xor_sync(0xFFFFFFFFu, acc0.x, o);
            acc0.y += __shfl_xor_sync(0xFFFFFFFFu, acc0.y, o);
            acc0.z += __shfl_xor_sync(0xFFFFFFFFu, acc0.z, o);
            acc0.w += __shfl_xor_sync(0xFFFFFFFFu, acc0.w, o);
        }
        const float inv = 1.0f / (float)a;
        acc0.x *= inv; acc0.y *= inv; acc0.z *= inv; acc0.w *= inv;
        float ss = acc0.x * acc0.x + acc0.y * acc0.y + acc0.z * acc0.z + acc0.w * acc0.w;
#pragma unroll
        for (int o = 1; o <= 4; o <<= 1) ss += __shfl_xor_sync(0xFFFFFFFFu, ss, o);
        if (g == 0) {
            *(float4*)(out + (size_t)orow * D + (size_t)slab * kSlab + l8 * 4) = acc0;
            if (l8 == 0) atomicAdd(sumsq + bm, ss);
        }
    }
}

__global__ void __launch_bounds__(256) mask_scale_kernel(float* __restrict__ out, int D, const float* __restrict__ sumsq,
                                                         const int* __restrict__ src_of_row, const int* __restrict__ total) {
    const int row = blockIdx.x;
    if (row >= *total) return;
    // no epsilon (core_system.py:407) — but a region whose mean is the zero vector, or whose features hold NaN / Inf, becomes
    // the zero vector instead of a NaN row (same rule as rvo_normalize_rows: one bad row must not poison every search)
    const float ssq = sumsq[src_of_row[row]];
    const float inv = (ssq != 0.f && isfinite(ssq)) ? 1.0f / sqrtf(ssq) : 0.f;
    float4* o = (float4*)(out + (size_t)row * D);
    for (int i = threadIdx.x; i < (D >> 2); i += blockDim.x) {
        float4 v = o[i];
        if (inv == 0.f) v = make_float4(0.f, 0.f, 0.f, 0.f);   // 0 * NaN would stay NaN
        else { v.x *= inv; v.y *= inv; v.z *= inv; v.w *= inv; }
        o[i] = v;
    }
}

int launch_mask_pool_tc(const uint16_t* feats, const uint8_t* masks, int B, int M, int P, int D, int lim, float* out,
                        int32_t* out_counts, int32_t* out_src, int32_t* out_total, int* img_base, int* area,
                        unsigned int* ticket, int sm_count, cudaStream_t stream, uint16_t* db, long long db_row0, int feat_f16);

size_t mask_pool_workspace_bytes(int B, int M, int P, int D) {
    (void)D;
    const size_t bm = (size_t)B * M;
    size_t n = 0;
    n += align_up((size_t)B * sizeof(int), 256) + 256;  // img_base + ticket (tensor-core path)
    n += align_up(bm * P * sizeof(uint16_t), 256);  // patch lists
    n += align_up(bm * sizeof(int), 256) * 3;       // area, out_row, src_of_row
    n += align_up(bm * sizeof(float), 256);         // sumsq
    return n + 1024;
}

int launch_mask_pool(const uint16_t* feats, const uint8_t* masks, int B, int M, int P, int D, int max_regions, float* out,
                     int32_t* out_counts, int32_t* out_src, int32_t* out_total, void* workspace, size_t workspace_bytes,
                     int sm_count, cudaStream_t stream, uint16_t* db, long long db_row0, int feat_f16) {
    if (D % kSlab != 0) {
        set_error("mask_pool: D=%d must be a multiple of %d", D, kSlab);
        return RVO_E_INVALID;
    }
    if (P > 65535) {
        set_error("mask_pool: P=%d too large", P);
        return RVO_E_INVALID;
    }
    const int lim = (max_regions <= 0 || max_regions > M) ? M : max_regions;
    const int bm = B * M;
    Arena ar(workspace, workspace_bytes);
    int* img_base = ar.take<int>(B);
    unsigned int* ticket = ar.take<unsigned int>(1);
    uint16_t* idx = ar.take<uint16_t>((size_t)bm * P);
    int* area = ar.take<int>(bm);
    int* out_row = ar.take<int>(bm);
    int* src_of_row = ar.take<int>(bm);
    float* sumsq = ar.take<float>(bm);
    if (!ar.ok()) {
        set_error("mask_pool: workspace too small (%zu < %zu)", workspace_bytes, ar.off);
        return RVO_E_WORKSPACE;
    }
    if (!g_force_cuda_core_pool) {
        // tensor-core path (mask_pool_tc.cu) whenever the image's whole output fits TMEM
        const int rc = launch_mask_pool_tc(feats, masks, B, M, P, D, lim, out, out_counts, out_src, out_total, img_base,
                                           area, ticket, sm_count, stream, db, db_row0, feat_f16);
        if (rc <= 0) return rc;
    }
    if (db) return RVO_E_UNSUPPORTED;   // the fused ingest exists on the tensor-core kernel only
    const int p_pad = (P + 7) & ~7;
    const size_t smem = (size_t)P * kSlab * 4 + (size_t)kPoolWarps * p_pad * 2;
    if (smem > 220 * 1024) {
        set_error("mask_pool: P=%d does not fit shared memory", P);
        return RVO_E_INVALID;
    }
    mask_index_kernel<<<(bm + 7) / 8, 256, 0, stream>>>(masks, bm, M, P, lim, idx, area);
    RVO_LAUNCHED();
    mask_offsets_kernel<<<1, 1024, 0, stream>>>(area, B, M, out_row, out_counts, src_of_row, out_total, sumsq);
    RVO_LAUNCHED();
    if (feat_f16) {
        RVO_CUDA(cudaFuncSetAttribute(mask_pool_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        mask_pool_kernel<true><<<dim3(D / kSlab, B), kPoolThreads, smem, stream>>>(feats, M, P, D, lim, p_pad, idx, area, out_row,
                                                                                   out, sumsq);
    } else {
        RVO_CUDA(cudaFuncSetAttribute(mask_pool_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        mask_pool_kernel<false><<<dim3(D / kSlab, B), kPoolThreads, smem, stream>>>(feats, M, P, D, lim, p_pad, idx, area, out_row,
                                                                                    out, sumsq);
    }
    RVO_LAUNCHED();
    mask_scale_kernel<<<bm, 256, 0, stream>>>(out, D, sumsq, src_of_row, out_total);
    RVO_LAUNCHED();
    if (out_src)
        RVO_CUDA(cudaMemcpyAsync(out_src, src_of_row, (size_t)bm * sizeof(int), cudaMemcpyDeviceToDevice, stream));
    return RVO_OK;
}

}  // namespace rvo
