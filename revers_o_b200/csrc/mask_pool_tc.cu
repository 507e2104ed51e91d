// K1 on tensor cores — mask pooling as the dense bf16 contraction it is:
//      out^T[d, m] = sum_p F[p, d] * W[m, p]          (per image; W = 0/1 mask weights, exact in bf16)
// followed by the mean (1/area), the L2 normalisation (core_system.py:407) and the compaction of
// non-empty regions (core_system.py:402-404) in the epilogue.  One persistent CTA per SM, one IMAGE per
// work item; the patch-feature map is read from HBM exactly once (TMA), the masks once.
//
//   A operand  = features, M = 128 channels of a slab, K = patches: F is [p][d] with d contiguous, i.e. an
//                MN-major A; TMA boxes {64 d, 64 p} (128-byte swizzle) land as 8-row x 128-byte atoms;
//                descriptor: LBO = 8 KiB (next 64 channels), SBO = 1 KiB (next 8 patches).
//   B operand  = masks, N = regions (padded to 16), K = patches: u8 -> bf16 0/1, converted by four warps into
//                K-major 128B-swizzled smem tiles, ONE 64-patch tile at a time: tile pc of the next image is
//                converted as soon as the current image's last MMA on tile pc has retired (per-tile
//                mbarriers), with the tile's mask bytes prefetched into registers before that wait.  The first
//                version converted a whole image between two images' MMAs (8-12 us of bubble per image).
//   D (TMEM)   = [128 channels][slab * N + m] fp32: all D/128 slabs of an image (D/128 * N <= 512), in TWO
//                parts with their own full/empty barriers.  MMA order is part -> patch tile -> slab, so the
//                epilogue's sum-of-squares pass over a part runs under the MMAs of the next ones, and part 0 is
//                handed back (after its normalise + store pass) while the later parts are still being stored.
// Epilogue: pass 1 reduces ||e_m||^2 over channels (shuffle + smem), pass 2 writes e_m / ||e_m|| (TMEM loads
// software-pipelined against the stores).
// Warp roles: 0 TMA producer, 1 MMA issuer, 2-9 epilogue (two per TMEM lane quarter), 10-11 mask conversion.
#include "common.cuh"
#include "ptx.cuh"
#include "prep_scan_small.cuh"

namespace rvo {

using namespace ptx;

constexpr int kPtThreads = 384;       // 12 warps = 3 per SM sub-partition -> up to 168 registers per thread
constexpr int kPtCvtThreads = 64;     // mask-conversion warps 10..11
constexpr int kPtMaxStages = 12;
constexpr int kPtMaxPc = 16;          // patch tiles per image (P <= 1024)
constexpr int kPtStageBytes = 16384;  // 64 patches x 128 channels bf16
constexpr int kPtMaxParts = 4;        // TMEM accumulator parts with their own full/empty barriers
constexpr int kPtBars = 2 * kPtMaxStages + 2 * kPtMaxPc + 2 * kPtMaxParts + 4;

struct PoolTcParams {
    const uint8_t* masks;
    int* counts;           // [B] kept (non-empty, below the cap) regions per image — the caller's out_counts
    int* area;             // [B*M] set patches per region, 0 for regions at or beyond the cap
    int* out_total;        // sum of counts
    unsigned int* arrived; // grid-wide arrival counter (zeroed by the launcher): CTAs that have published their counts
    float* out;            // fp32 embeddings [B*M, D], compacted (may be null when db is given)
    uint16_t* db;          // optional: tiled bf16 DB storage receiving the same rows at db_row0 + output row (fused ingest)
    long long db_row0;
    int* out_src;
    int B, M, P, D, lim, n_pad, num_pc, num_slab, num_stages;
    int G;                 // region groups per image: work item w = image * G + group covers regions [group*n_pad, +n_pad)
                           // (G > 1 when all D/128 slabs of all regions do not fit the 512 TMEM columns, e.g. D = 1280 x 64 regions)
    uint32_t off_b, b_buf_bytes, off_misc;
    uint32_t one_mul;      // 0x80 * one_mul = the 16-bit pattern of 1.0 in the feature dtype: 127 (bf16 0x3F80), 120 (fp16 0x3C00)
    uint32_t f16;          // features (and therefore the converted masks) are fp16 instead of bf16
    unsigned long long* trace;     // option "pool_trace": [B][8] globaltimer stamps of the per-image pipeline events (or null)
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

__device__ __forceinline__ uint32_t ld_acquire_u32(const unsigned int* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void trace_stamp(const PoolTcParams& p, int b, int slot) {
    if (p.trace) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        p.trace[(size_t)b * 8 + slot] = t;
    }
}

__device__ __forceinline__ void named_bar(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// four mask bytes (any non-zero value = set) -> four bf16 0/1 in two words.  Non-zero bytes are flagged 0x80
// (SWAR), spread into 16-bit lanes and multiplied by 127: 0x80 * 127 = 0x3F80 = bf16(1.0).
__device__ __forceinline__ void mask4_to_bf16(uint32_t x, uint32_t& w0, uint32_t& w1, int& cnt, uint32_t one_mul) {
    const uint32_t f = (((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & 0x80808080u;
    cnt += __popc(f);
    w0 = __byte_perm(f, 0u, 0x4140) * one_mul;   // bytes {f0, 0, f1, 0}; one_mul = 127: bf16 1.0, 120: fp16 1.0
    w1 = __byte_perm(f, 0u, 0x4342) * one_mul;   // bytes {f2, 0, f3, 0}
}

// TO_DB: the fused-ingest variant (bf16 rows into the tiled DB, fp32 output optional); the plain variant keeps the store pass
// free of the extra address arithmetic (it is issue-bound)
template <bool TO_DB>
__global__ void __launch_bounds__(kPtThreads, 1)
mask_pool_tc_kernel(const __grid_constant__ CUtensorMap tmap_f, const PoolTcParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* s_a = smem;                                  // [num_stages][16 KiB] feature tiles
    uint8_t* s_b = smem + p.off_b;                        // [num_pc][n_pad x 128 B] mask tiles (single buffer)
    int* s_outrow = (int*)(smem + p.off_misc);            // [2][64]
    float* s_invarea = (float*)(s_outrow + 128);          // [2][64]
    float* s_ss = s_invarea + 128;                        // [64] + [4][64] per-quarter partials
    uint64_t* bars = (uint64_t*)(s_ss + 64 + 256);
    uint64_t* bar_full = bars;                            // [kPtMaxStages]
    uint64_t* bar_empty = bars + kPtMaxStages;            // [kPtMaxStages]
    uint64_t* bar_bfull = bars + 2 * kPtMaxStages;        // [kPtMaxPc] mask tile pc of the next image converted
    uint64_t* bar_bfree = bar_bfull + kPtMaxPc;           // [kPtMaxPc] this image's MMAs on tile pc retired
    uint64_t* bar_tfull = bar_bfree + kPtMaxPc;           // [kPtMaxParts] accumulators of a part complete
    uint64_t* bar_tempty = bar_tfull + kPtMaxParts;       // [kPtMaxParts] part drained
    uint64_t* bar_mfree = bar_tempty + kPtMaxParts;       // [2] epilogue done with s_outrow/s_invarea[buf]
    uint64_t* bar_mready = bar_mfree + 2;                 // [2] s_outrow/s_invarea[buf] of an image written
    uint32_t* s_tmem = (uint32_t*)(bars + kPtBars);
    const uint32_t kPtStages = (uint32_t)p.num_stages;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_pad = p.n_pad, num_pc = p.num_pc, num_slab = p.num_slab;
    // the slabs form `parts` equal groups, each with its own TMEM full/empty barriers: between two images the MMA issuer only
    // waits for (sum of squares of the LAST part) + (store of the FIRST part).  Two parts; four were measured SLOWER (0.086 vs
    // 0.074 ms per configs[2] batch: the store pass restarts its TMEM-load pipeline and re-derives its scales per part).
    const int parts = (num_slab % 2 == 0) ? 2 : 1;
    const int spp = num_slab / parts;                     // slabs per part
    const int last_half = parts - 1;
    const uint32_t b_tile_bytes = (uint32_t)n_pad * 128u;

    if (threadIdx.x == 0) {
        if (smem_u32(smem) & 1023u) __trap();
        for (int i = 0; i < kPtMaxStages; ++i) { mbar_init(&bar_full[i], 1); mbar_init(&bar_empty[i], 1); }
        for (int i = 0; i < kPtMaxPc; ++i) { mbar_init(&bar_bfull[i], 1); mbar_init(&bar_bfree[i], 1); }
        for (int i = 0; i < kPtMaxParts; ++i) {
            mbar_init(&bar_tfull[i], 1);
            mbar_init(&bar_tempty[i], 8);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bar_mfree[i], 8);
            mbar_init(&bar_mready[i], 1);
        }
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
            for (int w = blockIdx.x; w < p.B * p.G; w += gridDim.x) {
                const int b = w / p.G;
                for (int h = 0; h <= last_half; ++h) {
                    const int s0 = h * spp, s1 = s0 + spp;
                    for (int pc = 0; pc < num_pc; ++pc)
                        for (int s = s0; s < s1; ++s) {
                            mbar_wait(&bar_empty[stage], phase ^ 1);
                            mbar_expect_tx(&bar_full[stage], kPtStageBytes);
                            uint8_t* dst = s_a + (size_t)stage * kPtStageBytes;
                            tma_load_3d(&tmap_f, &bar_full[stage], dst, s * 128, pc * 64, b, kEvictFirst);
                            tma_load_3d(&tmap_f, &bar_full[stage], dst + 8192, s * 128 + 64, pc * 64, b, kEvictFirst);
                            if (++stage == kPtStages) { stage = 0; phase ^= 1; }
                        }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (elect_one()) {
            // A: MN-major (bit 15), B: K-major, bf16 x bf16 -> fp32, M = 128, N = n_pad
            // fp16 features: a_format = b_format = F16 (0) instead of BF16 (1); the masks are converted to the same type
            const uint32_t idesc = (make_idesc_bf16(128, (uint32_t)n_pad) & (p.f16 ? ~((1u << 7) | (1u << 10)) : ~0u)) | (1u << 15);
            uint32_t stage = 0, phase = 0, it = 0;
            const uint32_t sB0 = smem_u32(s_b);
            for (int w = blockIdx.x; w < p.B * p.G; w += gridDim.x, ++it) {
                const int b = w / p.G;
                for (int h = 0; h <= last_half; ++h) {
                    const int s0 = h * spp, s1 = s0 + spp;
                    mbar_wait(&bar_tempty[h], (it & 1u) ^ 1u);        // this half of the previous image drained from TMEM
                    if (h == 0) trace_stamp(p, b, 1);
                    tc_fence_after();
                    for (int pc = 0; pc < num_pc; ++pc) {
                        if (h == 0) {
                            mbar_wait(&bar_bfull[pc], it & 1u);       // mask tile pc of this image is in smem
                            if (pc == 0) trace_stamp(p, b, 0);
                            tc_fence_after();
                        }
                        const uint32_t sB = sB0 + (uint32_t)pc * b_tile_bytes;
                        for (int s = s0; s < s1; ++s) {
                            mbar_wait(&bar_full[stage], phase);
                            tc_fence_after();
                            const uint32_t sA = smem_u32(s_a + (size_t)stage * kPtStageBytes);
                            const uint32_t d_tmem = tmem_base + (uint32_t)(s * n_pad);
#pragma unroll
                            for (int k4 = 0; k4 < 4; ++k4)
                                umma_bf16(d_tmem, make_mnmajor_sw128_desc(sA + k4 * 2048, 8192, 1024),
                                          make_kmajor_sw128_desc(sB + k4 * 32), idesc, (uint32_t)((pc | k4) != 0));
                            umma_commit(&bar_empty[stage]);
                            if (++stage == kPtStages) { stage = 0; phase ^= 1; }
                        }
                        if (h == last_half) umma_commit(&bar_bfree[pc]);   // tile pc may be overwritten with the next image's
                    }
                    umma_commit(&bar_tfull[h]);
                }
                trace_stamp(p, b, 2);
            }
        }
    } else if (warp < 10) {
        // ===================== epilogue (warps 2..9): mean, ||.||, normalise, compacted store =====================
        // two warps per TMEM lane quarter; warp `half_w` owns the 16-region groups half_w, half_w + 2
        const int ew = warp - 2, half_w = ew >> 2;
        const uint32_t lane_base = (uint32_t)(warp & 3) * 32u;
        const int ch = (int)lane_base + lane;  // channel inside a slab
        const int ngroups = n_pad / 16;
        float* s_part = s_ss + 64;             // [4 quarters][64] per-warp partial sums of squares
        // ---- prologue (these warps idle until the first accumulators are complete): areas and kept-region counts of
        // ALL images of this CTA, published grid-wide; the compaction offsets are a prefix sum over every image's count
        {
            int* s_area = (int*)s_part;        // [64] set patches per region of the image being counted
            const int tid = threadIdx.x - 64;  // 0..255
            const bool vec16 = (p.P & 15) == 0 && (((uintptr_t)p.masks) & 15) == 0;
            const int cpr = p.P >> 4;          // 16-byte chunks per mask row
            for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
                if (tid < 64) s_area[tid] = 0;
                named_bar(2, 256);
                const uint8_t* img = p.masks + (size_t)b * p.M * p.P;
                const int lim_m = p.M < p.lim ? p.M : p.lim;
                if (vec16) {
                    // flat (row, chunk) slots, four independent 16-byte loads in flight per thread
                    const int slots = lim_m * cpr;
                    for (int f0 = tid; f0 < slots; f0 += 4 * 256) {
                        uint4 v[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int f = f0 + u * 256;
                            v[u] = make_uint4(0, 0, 0, 0);
                            if (f < slots) v[u] = __ldg((const uint4*)(img + (size_t)(f / cpr) * p.P) + (f % cpr));
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const uint32_t w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
                            int n = 0;
#pragma unroll
                            for (int hh = 0; hh < 4; ++hh)
                                n += __popc((((w[hh] & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | w[hh]) & 0x80808080u);
                            if (n) atomicAdd(&s_area[(f0 + u * 256) / cpr], n);
                        }
                    }
                } else {
                    for (int m = ew; m < lim_m; m += 8) {
                        const uint8_t* row = img + (size_t)m * p.P;
                        int n = 0;
                        for (int i = lane; i < p.P; i += 32) n += __ldg(row + i) != 0;
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xFFFFFFFFu, n, o);
                        if (lane == 0) s_area[m] = n;
                    }
                }
                named_bar(2, 256);
                if (tid < 64) {   // warps 2 and 3
                    const int a = tid < p.M ? s_area[tid] : 0;
                    if (tid < p.M) p.area[(size_t)b * p.M + tid] = a;
                    const unsigned bal = __ballot_sync(0xFFFFFFFFu, a > 0);
                    if (lane == 0) s_area[64 + (tid >> 5)] = __popc(bal);
                }
                named_bar(2, 256);
                if (tid == 0) p.counts[b] = s_area[64] + s_area[65];
            }
            named_bar(2, 256);
            if (tid == 0) {
                __threadfence();               // areas and counts (ordered by the barrier) before the arrival
                atomicAdd(p.arrived, 1u);
            }
        }
        uint32_t it = 0;
        for (int w = blockIdx.x; w < p.B * p.G; w += gridDim.x, ++it) {
            const int b = w / p.G, m_base = (w - b * p.G) * n_pad;
            const uint32_t buf = it & 1u;
            const float* invarea = s_invarea + buf * 64;
            const int* outrow = s_outrow + buf * 64;
            const uint32_t taddr = tmem_base + (lane_base << 16);
            // pass 1: ||mean_m||^2 over all channels; half 0 is reduced while the MMAs of half 1 run
            float ss[2][16];
#pragma unroll
            for (int gi = 0; gi < 2; ++gi)
#pragma unroll
                for (int i = 0; i < 16; ++i) ss[gi][i] = 0.f;
            for (int h = 0; h <= last_half; ++h) {
                const int s0 = h * spp, s1 = s0 + spp;
                mbar_wait(&bar_tfull[h], it & 1u);
                tc_fence_after();
                if (threadIdx.x == 64 && (h == 0 || h == last_half)) trace_stamp(p, b, h == 0 ? 3 : 4);
#pragma unroll
                for (int gi = 0; gi < 2; ++gi) {
                    const int g = half_w + 2 * gi;
                    if (g >= ngroups) break;
                    const int m0 = g * 16;
                    for (int s = s0; s < s1; s += 2) {
                        uint32_t v[2][16];
                        tmem_ld_x16(taddr + (uint32_t)(s * n_pad + m0), v[0]);
                        if (s + 1 < s1) tmem_ld_x16(taddr + (uint32_t)((s + 1) * n_pad + m0), v[1]);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const float e = __uint_as_float(v[0][i]);
                            ss[gi][i] = fmaf(e, e, ss[gi][i]);
                        }
                        if (s + 1 < s1) {
#pragma unroll
                            for (int i = 0; i < 16; ++i) {
                                const float e = __uint_as_float(v[1][i]);
                                ss[gi][i] = fmaf(e, e, ss[gi][i]);
                            }
                        }
                    }
                }
            }
#pragma unroll
            for (int gi = 0; gi < 2; ++gi) {
                const int g = half_w + 2 * gi;
                if (g >= ngroups) break;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    float x = ss[gi][i];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xFFFFFFFFu, x, o);
                    if (lane == 0) s_part[(warp & 3) * 64 + g * 16 + i] = x;
                }
            }
            // the image's output rows / 1/area (written by the conversion warp after the grid-wide count rendezvous) are
            // first needed here: sum_d (acc_d / area)^2 = (sum_d acc_d^2) / area^2
            mbar_wait(&bar_mready[buf], (it >> 1) & 1u);
            named_bar(2, 256);
            if (ew * 32 + lane < 64) {
                const int m = ew * 32 + lane;
                s_ss[m] = (s_part[m] + s_part[64 + m] + s_part[128 + m] + s_part[192 + m]) * invarea[m] * invarea[m];
            }
            named_bar(2, 256);
            if (threadIdx.x == 64) trace_stamp(p, b, 5);
            // pass 2: e / ||e|| -> compacted rows (no epsilon, core_system.py:407); one half at a time so that half 0
            // goes back to the MMA issuer early.  The TMEM load of step j+1 is in flight while step j is stored; the
            // slab loop is unrolled over its maximum (4 slabs per half at D = 1024) so that the store addresses are
            // immediates off one 64-bit base per region.
            float* __restrict__ gout = p.out + ch;
            for (int h = 0; h <= last_half; ++h) {
                const int s0 = h * spp, s1 = s0 + spp;
                const int ns = s1 - s0;
#pragma unroll
                for (int gi = 0; gi < 2; ++gi) {
                    const int g = half_w + 2 * gi;
                    if (g >= ngroups) break;
                    const int m0 = g * 16;
                    float sc[16];
                    int rr[16];                     // output row of the region, < 0: empty region (dropped)
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        // a region whose mean is the zero vector, or whose features hold NaN / Inf, is stored as the zero
                        // vector (sc = 0), never as a NaN row: same rule as rvo_normalize_rows
                        const float ssq = s_ss[m0 + i];
                        sc[i] = (ssq != 0.f && isfinite(ssq)) ? invarea[m0 + i] * rsqrtf(ssq) : 0.f;
                        rr[i] = outrow[m0 + i];
                    }
                    for (int j0 = 0; j0 < ns; j0 += 4) {
                        uint32_t v[2][16];
                        tmem_ld_x16(taddr + (uint32_t)((s0 + j0) * n_pad + m0), v[0]);
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) {
                            const int j = j0 + jj;
                            if (j < ns) {
                                tmem_ld_wait();
                                if (jj < 3 && j + 1 < ns) tmem_ld_x16(taddr + (uint32_t)((s0 + j + 1) * n_pad + m0), v[(jj + 1) & 1]);
                                if (!TO_DB || p.out) {
                                    float* __restrict__ o = gout + (size_t)(s0 + j0) * 128 + jj * 128;
#pragma unroll
                                    for (int i = 0; i < 16; ++i)
                                        if (rr[i] >= 0)
                                            __stcs(o + (size_t)rr[i] * (size_t)p.D,
                                                   sc[i] != 0.f ? __uint_as_float(v[jj & 1][i]) * sc[i] : 0.f);
                                }
                                if constexpr (TO_DB) {
                                    // tiled DB storage [row/128][col/64][row%128][col%64]: the warp's 32 channels of one
                                    // region are 64 contiguous bytes
                                    const int col = (s0 + j) * 128 + ch;
                                    const size_t nk = (size_t)(p.D >> 6);
                                    uint16_t* __restrict__ t0 = p.db + (size_t)(col >> 6) * (kTileRows * kTileCols) + (col & 63);
#pragma unroll
                                    for (int i = 0; i < 16; ++i)
                                        if (rr[i] >= 0) {
                                            const long long row = p.db_row0 + rr[i];
                                            t0[((size_t)(row >> 7) * nk) * (kTileRows * kTileCols) + (size_t)(row & 127) * kTileCols] =
                                                __bfloat16_as_ushort(__float2bfloat16_rn(
                                                    sc[i] != 0.f ? __uint_as_float(v[jj & 1][i]) * sc[i] : 0.f));
                                        }
                                }
                            }
                        }
                    }
                }
                tc_fence_before();
                named_bar(2, 256);
                if (lane == 0) mbar_arrive(&bar_tempty[h]);
            }
            if (p.out_src && warp == 2)
                for (int m = lane; m < n_pad; m += 32)
                    if (outrow[m] >= 0) p.out_src[outrow[m]] = b * p.M + m_base + m;
            if (threadIdx.x == 64) trace_stamp(p, b, 6);
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_mfree[buf]);
        }
    } else {
        // ===================== mask conversion (warps 10..11): u8 [M][P] -> bf16 K-major swizzled tiles =====================
        const int t = threadIdx.x - 320;  // 0..63
        const size_t img_bytes = (size_t)p.M * p.P;
        const bool vec = (p.P & 7) == 0 && (((uintptr_t)p.masks) & 7) == 0;   // one 8-byte load per 8-patch chunk
        // the 8-patch chunks of one 64-patch tile: n_pad rows x 8 chunks, thread t takes chunks t, t + 64, ...
        auto load_tile = [&](int w, int pc, unsigned long long (&r)[8]) {
            const int b = w / p.G, m_base = (w - b * p.G) * n_pad;
            const uint8_t* src = p.masks + (size_t)b * img_bytes;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int idx = t + kPtCvtThreads * j, m = m_base + (idx >> 3), p0 = pc * 64 + (idx & 7) * 8;
                unsigned long long bytes = 0ull;
                if ((idx >> 3) < n_pad && m < p.M && m < p.lim && p0 < p.P) {
                    const uint8_t* row = src + (size_t)m * p.P;
                    if (vec) {
                        bytes = __ldg((const unsigned long long*)(row + p0));
                    } else {
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            if (p0 + i < p.P) bytes |= (unsigned long long)__ldg(row + p0 + i) << (8 * i);
                    }
                }
                r[j] = bytes;
            }
        };
        unsigned long long cur[8], nxt[8];
        if ((int)blockIdx.x < p.B * p.G) load_tile(blockIdx.x, 0, cur);
        int img_base = 0, prefix_from = 0;   // warp 10: kept regions of the images before `prefix_from`
        uint32_t it = 0;
        for (int w = blockIdx.x; w < p.B * p.G; w += gridDim.x, ++it) {
            const int b = w / p.G, m_base = (w - b * p.G) * n_pad;
            const uint32_t buf = it & 1u;
            for (int pc = 0; pc < num_pc; ++pc) {
                // prefetch the next tile's mask bytes, then wait for this tile's smem to be released
                const bool more = pc + 1 < num_pc;
                const int nw = more ? w : w + (int)gridDim.x;
                if (nw < p.B * p.G) load_tile(nw, more ? pc + 1 : 0, nxt);
                mbar_wait(&bar_bfree[pc], (it & 1u) ^ 1u);          // previous image's MMAs no longer read tile pc
                if (t == 0 && pc == 0) trace_stamp(p, b, 7);
                uint8_t* tile = s_b + (size_t)pc * b_tile_bytes;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int idx = t + kPtCvtThreads * j, m = idx >> 3, cin = idx & 7;
                    if (m < n_pad) {
                        uint32_t w0, w1, w2, w3;
                        int unused = 0;
                        mask4_to_bf16((uint32_t)cur[j], w0, w1, unused, p.one_mul);
                        mask4_to_bf16((uint32_t)(cur[j] >> 32), w2, w3, unused, p.one_mul);
                        *(uint4*)(tile + (size_t)m * 128 + (size_t)((cin ^ (m & 7)) << 4)) = make_uint4(w0, w1, w2, w3);
                    }
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) cur[j] = nxt[j];
                fence_proxy_async();   // generic-proxy smem writes -> visible to the tensor core's async proxy
                named_bar(3, kPtCvtThreads);
                if (t == 0) mbar_arrive(&bar_bfull[pc]);
                if (warp == 10 && pc == num_pc - 1) {  // needed by the epilogue only (bar_mready): after the last tile is handed over
                    // compacted output rows of this image: img_base[b] + number of non-empty regions before m
                    mbar_wait(&bar_mfree[buf], ((it >> 1) & 1u) ^ 1u);  // epilogue of image it-2 released the small arrays
                    if (it == 0) {
                        // every CTA of the (co-resident, <= one per SM) grid has published the counts of its images
                        uint32_t spins = 0;
                        while (ld_acquire_u32(p.arrived) < gridDim.x) {
                            __nanosleep(64);
                            if (++spins == (1u << 24)) {
                                if (lane == 0) printf("rvo: mask_pool grid rendezvous timed out (block %d)\n", blockIdx.x);
                                __trap();
                            }
                        }
                    }
                    // exclusive prefix of the kept counts, carried from this CTA's previous image (<= gridDim.x new terms)
                    int add = 0;
                    for (int i = prefix_from + lane; i < b; i += 32) add += __ldcg(p.counts + i);
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) add += __shfl_xor_sync(0xFFFFFFFFu, add, o);
                    img_base += add;
                    prefix_from = b;
                    const int base = img_base;
                    if (b == p.B - 1 && lane == 0) *p.out_total = img_base + __ldcg(p.counts + b);
                    // non-empty regions of the image BEFORE this work item's group, then the group's own rows (all 64 slots
                    // of the small arrays are rewritten: slots beyond the group or beyond M are marked empty)
                    int run = 0;
                    for (int m0 = 0; m0 < m_base; m0 += 32) {
                        const int m = m0 + lane;
                        const int a = (m < m_base && m < p.M && m < p.lim) ? __ldcg(p.area + (size_t)b * p.M + m) : 0;
                        run += __popc(__ballot_sync(0xFFFFFFFFu, a > 0));
                    }
                    for (int ml0 = 0; ml0 < 64; ml0 += 32) {
                        const int ml = ml0 + lane, m = m_base + ml;
                        const int a = (ml < n_pad && m < p.M && m < p.lim) ? __ldcg(p.area + (size_t)b * p.M + m) : 0;
                        const unsigned bal = __ballot_sync(0xFFFFFFFFu, a > 0);
                        s_outrow[buf * 64 + ml] = a > 0 ? base + run + __popc(bal & ((1u << lane) - 1u)) : -1;
                        s_invarea[buf * 64 + ml] = a > 0 ? 1.0f / (float)a : 0.f;
                        run += __popc(bal);
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bar_mready[buf]);
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

extern void* g_pool_trace;

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// returns RVO_OK when the tensor-core path ran, 1 when the shape is not supported by it (caller falls back to the
// CUDA-core kernel — same results), <0 on error.
int launch_mask_pool_tc(const uint16_t* feats, const uint8_t* masks, int B, int M, int P, int D, int lim, float* out,
                        int32_t* out_counts, int32_t* out_src, int32_t* out_total, int* img_base, int* area,
                        unsigned int* ticket, int sm_count, cudaStream_t stream, uint16_t* db, long long db_row0, int feat_f16) {
    const int n_full = (M + 15) / 16 * 16;
    const int num_pc = (P + 63) / 64, num_slab = D / 128;
    if (D % 128 != 0 || n_full > 64 || num_slab <= 0) return 1;
    // all D/128 slabs of a work item's regions live in TMEM (512 columns): when an image's regions do not fit together
    // (PE-Core-G14: 10 slabs x 64 regions) they are split into groups, one work item each — the feature tiles of the image are
    // streamed once per group (the second pass mostly hits L2)
    int n_pad = n_full;
    if (num_slab * n_pad > 512) n_pad = 512 / num_slab / 16 * 16;
    if (n_pad < 16) return 1;
    const int G = (n_full + n_pad - 1) / n_pad;
    if (num_pc > kPtMaxPc) return 1;
    const size_t b_buf = (size_t)num_pc * n_pad * 128;
    const size_t tail = (128 * 2 + 64 + 256) * 4 + (kPtBars + 1) * 8;
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
    CUresult r = enc(&tm, feat_f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<uint16_t*>(feats), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("mask_pool: cuTensorMapEncodeTiled failed (%d)", (int)r);
        return RVO_E_CUDA;
    }
    PoolTcParams p;
    p.masks = masks;
    p.counts = out_counts;
    p.area = area;
    p.out_total = out_total;
    p.arrived = ticket;
    p.out = out;
    p.db = db;
    p.db_row0 = db_row0;
    p.out_src = out_src;
    p.B = B; p.M = M; p.P = P; p.D = D; p.lim = lim;
    p.n_pad = n_pad; p.num_pc = num_pc; p.num_slab = num_slab; p.num_stages = stages;
    p.G = G;
    p.off_b = (uint32_t)off_b;
    p.b_buf_bytes = (uint32_t)b_buf;
    p.off_misc = (uint32_t)off_misc;
    p.one_mul = feat_f16 ? 120u : 127u;
    p.f16 = feat_f16 ? 1u : 0u;
    p.trace = (unsigned long long*)g_pool_trace;
    RVO_CUDA(cudaFuncSetAttribute(mask_pool_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    RVO_CUDA(cudaFuncSetAttribute(mask_pool_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = B * G < sm_count ? B * G : sm_count;
    // ONE launch, COOPERATIVE: the grid is persistent with at most one CTA per SM and the in-kernel rendezvous on `arrived`
    // needs every CTA resident at once.  The cooperative attribute makes the driver guarantee that (the launch waits for the
    // SMs, or fails if the grid could never fit) — two pooling calls on different streams can no longer hold half of the SMs
    // each and wait for one another.  The launcher zeroes the counter on the same stream.
    RVO_CUDA(cudaMemsetAsync(ticket, 0, sizeof(unsigned int), stream));
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(kPtThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (db) RVO_CUDA(cudaLaunchKernelEx(&cfg, mask_pool_tc_kernel<true>, tm, p));
    else RVO_CUDA(cudaLaunchKernelEx(&cfg, mask_pool_tc_kernel<false>, tm, p));
    RVO_LAUNCHED();
    return RVO_OK;
}

}  // namespace rvo
