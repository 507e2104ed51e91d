// K1 on tensor cores — mask pooling as the dense bf16 contraction it is:
//      out^T[d, m] = sum_p F[p, d] * W[m, p]          (per image; W = 0/1 mask weights, exact in bf16)
// followed by the mean (1/area), the L2 normalisation (core_system.py:407) and the compaction of
// non-empty regions (core_system.py:402-404) in the epilogue.  One persistent CTA per SM, one IMAGE per
// work item; the patch-feature map is read from HBM exactly once (TMA), the masks once.
//
//   A operand  = features, M = 128 channels of a slab, K = patches: F is [p][d] with d contiguous, i.e. an
//                MN-major A; TMA boxes {64 d, 64 p} (128-byte swizzle) land as 8-row x 128-byte atoms;
//                descriptor: LBO = 8 KiB (next 64 channels), SBO = 1 KiB (next 8 patches).
//   B operand  = masks, N = regions (padded to 16), K = patches: converted u8 -> bf16 0/1 by four warps
//                straight into the K-major 128B-swizzled smem tiles (double-buffered per image).
//   D (TMEM)   = [128 channels][slab * N + m] fp32: all D/128 slabs of an image at once (D/128 * N <= 512),
//                so the epilogue sees the complete un-normalised embedding of every region of the image:
//                pass 1 reduces ||e_m||^2 over channels (shuffle + smem), pass 2 writes e_m / ||e_m||.
// Warp roles: 0 TMA producer, 1 MMA issuer, 2-9 epilogue (two per TMEM lane quarter), 10-13 mask conversion.
#include "common.cuh"
#include "ptx.cuh"
#include "prep_scan_small.cuh"

namespace rvo {

using namespace ptx;

constexpr int kPtThreads = 448;
constexpr int kPtMaxStages = 12;
constexpr int kPtStageBytes = 16384;  // 64 patches x 128 channels bf16

struct PoolTcParams {
    const uint8_t* masks;
    const int* img_base;   // exclusive scan of kept regions per image
    float* out;
    int* out_src;
    int B, M, P, D, lim, n_pad, num_pc, num_slab, num_stages;
    uint32_t off_b, b_buf_bytes, off_misc;
};

__device__ __forceinline__ void tma_load_3d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1, int c2,
                                            uint64_t hint) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4, %5}], [%2], %6;"
        :
        : "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "l"(hint)
        : "memory");
}

// MN-major operand, 128-byte swizzle: 64 MN-elements x 8 K-rows atoms; LBO = byte distance between
// 64-element MN blocks, SBO = byte distance between 8-row K groups.
__device__ __forceinline__ uint64_t make_mnmajor_sw128_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

__device__ __forceinline__ void named_bar(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__global__ void __launch_bounds__(kPtThreads, 1)
mask_pool_tc_kernel(const __grid_constant__ CUtensorMap tmap_f, const PoolTcParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* s_a = smem;                                  // [num_stages][16 KiB] feature tiles
    uint8_t* s_b = smem + p.off_b;                        // [num_pc][n_pad x 128 B] mask tiles (single buffer)
    int* s_area = (int*)(smem + p.off_misc);              // [2][64]
    int* s_outrow = s_area + 128;                         // [2][64]
    float* s_invarea = (float*)(s_outrow + 128);          // [2][64]
    float* s_ss = s_invarea + 128;                        // [64] + [4][64] per-quarter partials
    uint64_t* bars = (uint64_t*)(s_ss + 64 + 256);
    uint64_t* bar_full = bars;                            // [kPtMaxStages]
    uint64_t* bar_empty = bars + kPtMaxStages;            // [kPtMaxStages]
    uint64_t* bar_bfull = bars + 2 * kPtMaxStages;        // masks of the next image converted
    uint64_t* bar_bfree = bars + 2 * kPtMaxStages + 1;    // MMAs of an image retired: mask tiles may be overwritten
    uint64_t* bar_tfull = bars + 2 * kPtMaxStages + 2;    // accumulators of an image complete
    uint64_t* bar_tempty = bars + 2 * kPtMaxStages + 3;   // TMEM drained
    uint64_t* bar_mfree = bars + 2 * kPtMaxStages + 4;    // [2] epilogue done with s_outrow/s_invarea[buf]
    uint32_t* s_tmem = (uint32_t*)(bars + 2 * kPtMaxStages + 6);
    const uint32_t kPtStages = (uint32_t)p.num_stages;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_pad = p.n_pad, num_pc = p.num_pc, num_slab = p.num_slab;
    const uint32_t b_tile_bytes = (uint32_t)n_pad * 128u;

    if (threadIdx.x == 0) {
        if (smem_u32(smem) & 1023u) __trap();
        for (int i = 0; i < kPtMaxStages; ++i) { mbar_init(&bar_full[i], 1); mbar_init(&bar_empty[i], 1); }
        mbar_init(bar_bfull, 1);
        mbar_init(bar_bfree, 1);
        for (int i = 0; i < 2; ++i) mbar_init(&bar_mfree[i], 8);
        mbar_init(bar_tfull, 1);
        mbar_init(bar_tempty, 8);
        fence_mbar_init();
        fence_proxy_async();
    }
    if (warp == 0 && lane == 0) prefetch_tmap(&tmap_f);
    if (warp == 1) { tmem_alloc(s_tmem, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;

    if (warp == 0) {
        // ===================== TMA producer: features, each element exactly once =====================
        if (elect_one()) {
            uint32_t stage = 0, phase = 0;
            for (int b = blockIdx.x; b < p.B; b += gridDim.x)
                for (int s = 0; s < num_slab; ++s)
                    for (int pc = 0; pc < num_pc; ++pc) {
                        mbar_wait(&bar_empty[stage], phase ^ 1);
                        mbar_expect_tx(&bar_full[stage], kPtStageBytes);
                        uint8_t* dst = s_a + (size_t)stage * kPtStageBytes;
                        tma_load_3d(&tmap_f, &bar_full[stage], dst, s * 128, pc * 64, b, kEvictFirst);
                        tma_load_3d(&tmap_f, &bar_full[stage], dst + 8192, s * 128 + 64, pc * 64, b, kEvictFirst);
                        if (++stage == kPtStages) { stage = 0; phase ^= 1; }
                    }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (elect_one()) {
            // A: MN-major (bit 15), B: K-major, bf16 x bf16 -> fp32, M = 128, N = n_pad
            const uint32_t idesc = make_idesc_bf16(128, (uint32_t)n_pad) | (1u << 15);
            uint32_t stage = 0, phase = 0, it = 0;
            for (int b = blockIdx.x; b < p.B; b += gridDim.x, ++it) {
                mbar_wait(bar_bfull, it & 1u);                    // masks of this image are in smem
                mbar_wait(bar_tempty, (it & 1u) ^ 1u);            // previous image drained from TMEM
                tc_fence_after();
                const uint32_t sB0 = smem_u32(s_b);
                for (int s = 0; s < num_slab; ++s) {
                    const uint32_t d_tmem = tmem_base + (uint32_t)(s * n_pad);
                    for (int pc = 0; pc < num_pc; ++pc) {
                        mbar_wait(&bar_full[stage], phase);
                        tc_fence_after();
                        const uint32_t sA = smem_u32(s_a + (size_t)stage * kPtStageBytes);
                        const uint32_t sB = sB0 + (uint32_t)pc * b_tile_bytes;
#pragma unroll
                        for (int k4 = 0; k4 < 4; ++k4)
                            umma_bf16(d_tmem, make_mnmajor_sw128_desc(sA + k4 * 2048, 8192, 1024),
                                      make_kmajor_sw128_desc(sB + k4 * 32), idesc, (uint32_t)((pc | k4) != 0));
                        umma_commit(&bar_empty[stage]);
                        if (++stage == kPtStages) { stage = 0; phase ^= 1; }
                    }
                }
                umma_commit(bar_tfull);
                umma_commit(bar_bfree);
            }
        }
    } else if (warp < 10) {
        // ===================== epilogue (warps 2..9): mean, ||.||, normalise, compacted store =====================
        // two warps per TMEM lane quarter; warp `half` owns the 16-region groups half, half+2, ...
        const int ew = warp - 2, half = ew >> 2;
        const uint32_t lane_base = (uint32_t)(warp & 3) * 32u;
        const int ch = (int)lane_base + lane;  // channel inside a slab
        const int ngroups = n_pad / 16;
        float* s_part = s_ss + 64;             // [4 quarters][64] per-warp partial sums of squares
        uint32_t it = 0;
        for (int b = blockIdx.x; b < p.B; b += gridDim.x, ++it) {
            const uint32_t buf = it & 1u;
            const float* invarea = s_invarea + buf * 64;
            const int* outrow = s_outrow + buf * 64;
            mbar_wait(bar_tfull, it & 1u);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (lane_base << 16);
            // pass 1: ||mean_m||^2 over all channels
            for (int g = half; g < ngroups; g += 2) {
                const int m0 = g * 16;
                float ss[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) ss[i] = 0.f;
                for (int s = 0; s < num_slab; s += 2) {
                    uint32_t v[2][16];
                    tmem_ld_x16(taddr + (uint32_t)(s * n_pad + m0), v[0]);
                    if (s + 1 < num_slab) tmem_ld_x16(taddr + (uint32_t)((s + 1) * n_pad + m0), v[1]);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float e = __uint_as_float(v[0][i]) * invarea[m0 + i];
                        ss[i] = fmaf(e, e, ss[i]);
                    }
                    if (s + 1 < num_slab) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const float e = __uint_as_float(v[1][i]) * invarea[m0 + i];
                            ss[i] = fmaf(e, e, ss[i]);
                        }
                    }
                }
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    float x = ss[i];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xFFFFFFFFu, x, o);
                    if (lane == 0) s_part[(warp & 3) * 64 + m0 + i] = x;
                }
            }
            named_bar(2, 256);
            if (ew * 32 + lane < 64) {
                const int m = ew * 32 + lane;
                s_ss[m] = s_part[m] + s_part[64 + m] + s_part[128 + m] + s_part[192 + m];
            }
            named_bar(2, 256);
            // pass 2: e / ||e|| -> compacted rows (no epsilon, core_system.py:407)
            for (int g = half; g < ngroups; g += 2) {
                const int m0 = g * 16;
                float sc[16];
                float* optr[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    sc[i] = invarea[m0 + i] / sqrtf(s_ss[m0 + i]);
                    const int r = outrow[m0 + i];
                    optr[i] = r >= 0 ? p.out + (size_t)r * p.D + ch : nullptr;
                }
                for (int s = 0; s < num_slab; ++s) {
                    uint32_t v[16];
                    tmem_ld_x16(taddr + (uint32_t)(s * n_pad + m0), v);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (optr[i]) optr[i][s * 128] = __uint_as_float(v[i]) * sc[i];
                }
            }
            if (p.out_src && warp == 2)
                for (int m = lane; m < n_pad; m += 32)
                    if (outrow[m] >= 0) p.out_src[outrow[m]] = b * p.M + m;
            tc_fence_before();
            named_bar(2, 256);
            if (lane == 0) {
                mbar_arrive(bar_tempty);
                mbar_arrive(&bar_mfree[buf]);
            }
        }
    } else {
        // ===================== mask conversion (warps 10..13): u8 [M][P] -> bf16 K-major swizzled tiles =====================
        const int t = threadIdx.x - 320;  // 0..127
        uint32_t it = 0;
        for (int b = blockIdx.x; b < p.B; b += gridDim.x, ++it) {
            const uint32_t buf = it & 1u;
            mbar_wait(bar_bfree, (it & 1u) ^ 1u);                   // previous image's MMAs no longer read the tiles
            mbar_wait(&bar_mfree[buf], ((it >> 1) & 1u) ^ 1u);      // epilogue of image it-2 released the small arrays
            int* area = s_area + buf * 64;
            uint8_t* dstb = s_b;
            const int chunks_per_row = num_pc * 8;  // 16-byte smem chunks (8 patches) per mask row
            const uint8_t* src = p.masks + (size_t)b * p.M * p.P;
            const bool vec = (p.P & 7) == 0 && (((uintptr_t)src) & 7) == 0;   // one 8-byte load per chunk
            for (int m = warp - 10; m < n_pad; m += 4) {                       // warp per mask row, lane per chunk
                const bool live = m < p.M && m < p.lim;
                const uint8_t* row = src + (size_t)m * p.P;
                int cnt = 0;
                for (int ck = lane; ck < chunks_per_row; ck += 32) {
                    const int p0 = ck * 8;
                    unsigned long long bytes = 0ull;
                    if (live) {
                        if (vec) {
                            if (p0 < p.P) bytes = __ldg((const unsigned long long*)(row + p0));
                        } else {
#pragma unroll
                            for (int i = 0; i < 8; ++i)
                                if (p0 + i < p.P) bytes |= (unsigned long long)row[p0 + i] << (8 * i);
                        }
                    }
                    uint32_t w[4];
#pragma unroll
                    for (int h = 0; h < 4; ++h) {
                        const uint32_t lo = (uint32_t)(bytes >> (16 * h)) & 0xFFu, hi = (uint32_t)(bytes >> (16 * h + 8)) & 0xFFu;
                        w[h] = (lo ? 0x00003F80u : 0u) | (hi ? 0x3F800000u : 0u);   // bf16 1.0 per set patch
                        cnt += (lo != 0) + (hi != 0);
                    }
                    const int pc = ck >> 3, cin = ck & 7;
                    uint4* d4 = (uint4*)(dstb + (size_t)pc * b_tile_bytes + (size_t)m * 128 + (size_t)((cin ^ (m & 7)) << 4));
                    *d4 = make_uint4(w[0], w[1], w[2], w[3]);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, o);
                if (lane == 0 && m < 64) area[m] = cnt;
            }
            named_bar(3, 128);
            if (warp == 10) {
                // compacted output rows of this image: img_base[b] + number of non-empty regions before m
                const int base = p.img_base[b];
                int run = 0;
                for (int m0 = 0; m0 < 64; m0 += 32) {
                    const int m = m0 + lane;
                    const int a = m < n_pad ? area[m] : 0;
                    const unsigned bal = __ballot_sync(0xFFFFFFFFu, a > 0);
                    s_outrow[buf * 64 + m] = a > 0 ? base + run + __popc(bal & ((1u << lane) - 1u)) : -1;
                    s_invarea[buf * 64 + m] = a > 0 ? 1.0f / (float)a : 0.f;
                    run += __popc(bal);
                }
            }
            fence_proxy_async();   // generic-proxy smem writes -> visible to the tensor core's async proxy
            named_bar(3, 128);
            if (t == 0) mbar_arrive(bar_bfull);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// per-image count of kept (non-empty, < lim) regions, then an exclusive scan by the last block to finish
__global__ void __launch_bounds__(256) mask_count_kernel(const uint8_t* __restrict__ masks, int B, int M, int P, int lim,
                                                         int* __restrict__ counts, int* __restrict__ img_base,
                                                         int* __restrict__ out_total, unsigned int* __restrict__ ticket) {
    __shared__ int s_cnt;
    __shared__ bool s_last;
    const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    const bool vec = (P & 15) == 0 && (((uintptr_t)masks) & 15) == 0;
    for (int m = warp; m < M && m < lim; m += 8) {
        const uint8_t* row = masks + ((size_t)b * M + m) * P;
        bool any = false;
        if (vec) {
            for (int i = lane; i < (P >> 4); i += 32) {
                const uint4 v = __ldg((const uint4*)row + i);
                any |= (v.x | v.y | v.z | v.w) != 0u;
            }
        } else {
            for (int i = lane; i < P; i += 32) any |= row[i] != 0;
        }
        if (__any_sync(0xFFFFFFFFu, any) && lane == 0) atomicAdd(&s_cnt, 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        counts[b] = s_cnt;
        __threadfence();
        s_last = atomicAdd(ticket, 1u) == (unsigned)(B - 1);
    }
    __syncthreads();
    if (s_last) {  // block-wide exclusive scan over the B per-image counts, 256 at a time
        __shared__ int s_scan[256];
        __shared__ int s_carry;
        if (threadIdx.x == 0) s_carry = 0;
        __threadfence();
        __syncthreads();
        for (int c0 = 0; c0 < B; c0 += 256) {
            const int i = c0 + threadIdx.x;
            const int v = i < B ? ((volatile int*)counts)[i] : 0;
            s_scan[threadIdx.x] = v;
            __syncthreads();
            for (int o = 1; o < 256; o <<= 1) {
                const int t = threadIdx.x >= o ? s_scan[threadIdx.x - o] : 0;
                __syncthreads();
                s_scan[threadIdx.x] += t;
                __syncthreads();
            }
            if (i < B) img_base[i] = s_carry + s_scan[threadIdx.x] - v;
            __syncthreads();
            if (threadIdx.x == 255) s_carry += s_scan[255];
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            *out_total = s_carry;
            *ticket = 0u;
        }
    }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// returns RVO_OK when the tensor-core path ran, 1 when the shape is not supported by it (caller falls back to the
// CUDA-core kernel — same results), <0 on error.
int launch_mask_pool_tc(const uint16_t* feats, const uint8_t* masks, int B, int M, int P, int D, int lim, float* out,
                        int32_t* out_counts, int32_t* out_src, int32_t* out_total, int* img_base, unsigned int* ticket,
                        int sm_count, cudaStream_t stream) {
    const int n_pad = (M + 15) / 16 * 16;
    const int num_pc = (P + 63) / 64, num_slab = D / 128;
    if (D % 128 != 0 || n_pad > 64 || num_slab * n_pad > 512) return 1;
    const size_t b_buf = (size_t)num_pc * n_pad * 128;
    const size_t tail = (128 * 3 + 64 + 256) * 4 + (2 * kPtMaxStages + 8) * 8;
    if (b_buf + tail + 2 * kPtStageBytes > 227 * 1024) return 1;
    int stages = (int)((227 * 1024 - b_buf - tail) / kPtStageBytes);   // as many 16 KiB feature tiles in flight as fit
    if (stages > kPtMaxStages) stages = kPtMaxStages;
    const size_t off_b = (size_t)stages * kPtStageBytes;
    const size_t off_misc = off_b + b_buf;
    const size_t smem = off_misc + tail;

    static PFN_encodeTiled enc = nullptr;
    if (!enc) {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess) {
            set_error("cuTensorMapEncodeTiled entry point unavailable");
            return RVO_E_NO_DEVICE;
        }
        enc = (PFN_encodeTiled)f;
    }
    CUtensorMap tm;
    cuuint64_t gdim[3] = {(cuuint64_t)D, (cuuint64_t)P, (cuuint64_t)B};
    cuuint64_t gstr[2] = {(cuuint64_t)D * 2ull, (cuuint64_t)P * D * 2ull};
    cuuint32_t box[3] = {64, 64, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<uint16_t*>(feats), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("mask_pool: cuTensorMapEncodeTiled failed (%d)", (int)r);
        return RVO_E_CUDA;
    }
    mask_count_kernel<<<B, 256, 0, stream>>>(masks, B, M, P, lim, out_counts, img_base, out_total, ticket);
    RVO_LAUNCHED();

    PoolTcParams p;
    p.masks = masks;
    p.img_base = img_base;
    p.out = out;
    p.out_src = out_src;
    p.B = B; p.M = M; p.P = P; p.D = D; p.lim = lim;
    p.n_pad = n_pad; p.num_pc = num_pc; p.num_slab = num_slab; p.num_stages = stages;
    p.off_b = (uint32_t)off_b;
    p.b_buf_bytes = (uint32_t)b_buf;
    p.off_misc = (uint32_t)off_misc;
    RVO_CUDA(cudaFuncSetAttribute(mask_pool_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = B < sm_count ? B : sm_count;
    mask_pool_tc_kernel<<<grid, kPtThreads, smem, stream>>>(tm, p);
    RVO_LAUNCHED();
    return RVO_OK;
}

}  // namespace rvo
