// (1) Row L2-normalisation (ingest / query preparation) and
// (2) the exact fp32 CUDA-core scan for batches of <= RVO_SMALL_Q queries — the reference's own
//     operating point (Q = 1, core_system.py:657): scores[q, r] = <q_hat, db[r]>, fp32 FMA over the
//     bf16 DB rows, written densely (4 B per 2 KB row read: <0.2 % extra traffic); HBM-bound.
#include "common.cuh"
#include "prep_scan_small.cuh"

namespace rvo {

// one warp per row: x / ||x||  (core_system.py:407,447; qdrant COSINE upsert/search normalisation)
// dst_bf16: row-major [n, dst_ld] when tiled_row0 < 0, else the tiled DB storage (common.cuh) in which this
// launch writes rows tiled_row0 .. tiled_row0 + n - 1 (dst_ld == d_pad).
__global__ void __launch_bounds__(256) normalize_rows_kernel(const float* __restrict__ src, long long n, int d,
                                                             long long src_ld, uint16_t* __restrict__ dst_bf16,
                                                             long long dst_ld, long long tiled_row0,
                                                             float* __restrict__ dst_f32, long long f32_ld,
                                                             float* __restrict__ margin_out, long long n_pad_rows,
                                                             uint4* __restrict__ zero_base, long long zero_u4) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // the seed scan may pre-launch (it waits for this grid)
    // scratch the search driver needs cleared (candidate counters): folded into this launch instead of a separate memset node
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < zero_u4; i += (long long)gridDim.x * blockDim.x)
        zero_base[i] = make_uint4(0, 0, 0, 0);
    if (row >= n) {
        // padding rows of the query operand (row-major outputs only): all-zero, margin 0
        if (row < n_pad_rows && tiled_row0 < 0) {
            if (dst_bf16)
                for (long long i = lane; i < dst_ld; i += 32) dst_bf16[(size_t)row * (size_t)dst_ld + i] = 0;
            if (dst_f32)
                for (long long i = lane; i < f32_ld; i += 32) dst_f32[(size_t)row * (size_t)f32_ld + i] = 0.f;
            if (margin_out && lane == 0) margin_out[row] = 0.f;
        }
        return;
    }
    const float* s = src + (size_t)row * (size_t)src_ld;
    // sum of squares, scaled by the row's largest magnitude so that rows with huge (but finite) components do not overflow
    // to inf and silently become the zero vector
    float amax = 0.f;
    for (int i = lane; i < d; i += 32) amax = fmaxf(amax, fabsf(s[i]));   // fmaxf drops NaNs: caught below through ss
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xFFFFFFFFu, amax, o));
    const bool rescale = amax > 1e18f || (amax < 1e-18f && amax > 0.f);
    const float pre = rescale ? 1.0f / amax : 1.0f;
    float ss = 0.f;
    for (int i = lane; i < d; i += 32) {
        const float v = s[i] * pre;
        ss = fmaf(v, v, ss);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xFFFFFFFFu, ss, o);
    const float nrm = sqrtf(ss);
    // zero rows stay zero; a row with a NaN / Inf component (norm not finite) is stored as the zero vector as well, so that one
    // bad vector cannot poison every search of the collection with NaN scores (it then scores 0 against everything)
    const bool usable = nrm != 0.f && isfinite(nrm);
    const float inv = usable ? pre / nrm : 0.f;
    if (dst_bf16) {
        if (tiled_row0 < 0) {
            uint16_t* o = dst_bf16 + (size_t)row * (size_t)dst_ld;
            float e2 = 0.f;  // ||bf16(q_hat) - q_hat||^2: the exact rounding error of THIS query operand
            for (long long i = lane; i < dst_ld; i += 32) {
                const float v = (usable && i < d) ? s[i] * inv : 0.f;
                const __nv_bfloat16 b = __float2bfloat16_rn(v);
                const float e = __bfloat162float(b) - v;
                e2 = fmaf(e, e, e2);
                o[i] = __bfloat16_as_ushort(b);
            }
            if (margin_out) {
#pragma unroll
                for (int o2 = 16; o2 > 0; o2 >>= 1) e2 += __shfl_xor_sync(0xFFFFFFFFu, e2, o2);
                if (lane == 0) margin_out[row] = query_margin(sqrtf(e2));
            }
        } else {
            const int nk = (int)(dst_ld / kTileCols);
            for (int i = lane; i < (int)dst_ld; i += 32) {
                const float v = (usable && i < d) ? s[i] * inv : 0.f;
                dst_bf16[tiled_offset(tiled_row0 + row, i, nk)] = __bfloat16_as_ushort(__float2bfloat16_rn(v));
            }
        }
    }
    if (dst_f32) {
        float* o = dst_f32 + (size_t)row * (size_t)f32_ld;
        for (long long i = lane; i < f32_ld; i += 32) o[i] = (usable && i < d) ? s[i] * inv : 0.f;
    }
}

int launch_normalize_rows(const float* src, long long n, int d, long long src_ld, uint16_t* dst_bf16, long long dst_ld,
                          long long tiled_row0, float* dst_f32, long long f32_ld, cudaStream_t stream,
                          float* margin_out, long long n_pad_rows, void* zero_base, size_t zero_bytes) {
    if (n_pad_rows < n) n_pad_rows = n;
    if (n_pad_rows <= 0 && zero_bytes == 0) return RVO_OK;
    long long blocks = (n_pad_rows + 7) / 8;
    if (blocks < 1) blocks = 1;
    normalize_rows_kernel<<<(unsigned)blocks, 256, 0, stream>>>(src, n, d, src_ld, dst_bf16, dst_ld, tiled_row0, dst_f32,
                                                                f32_ld, margin_out, n_pad_rows, (uint4*)zero_base,
                                                                (long long)(zero_bytes / 16));
    RVO_LAUNCHED();
    return RVO_OK;
}

// ---- small-Q scan --------------------------------------------------------------------------------
// One warp per pair of DB rows; lane l owns the 16-byte chunks l, l+32, ... of a row, and keeps the
// matching slices of all NQ queries in registers (no shared-memory traffic in the loop).
// uint4 index of the 16-byte chunk `idx` (0 .. d_pad/8) of row `r` in the tiled DB storage
__device__ __forceinline__ size_t tiled_u4(long long r, int idx, int nk) {
    return (((size_t)(r >> 7) * (size_t)nk + (size_t)(idx >> 3)) * kTileRows + (size_t)(r & 127)) * 8 + (size_t)(idx & 7);
}

template <int NQ, int CPL>
__global__ void __launch_bounds__(256) scan_small_kernel(const uint4* __restrict__ db, long long n_rows,
                                                         int nk, int nchunks, const float* __restrict__ qn,
                                                         long long qn_ld, float* __restrict__ out, long long out_ld) {
    const int lane = threadIdx.x & 31;
    const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
    asm volatile("griddepcontrol.wait;" ::: "memory");               // the normalised queries come from the previous kernel
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // (programmatic dependent launch, see ptx.cuh)

    float qr[NQ][CPL][8];
#pragma unroll
    for (int q = 0; q < NQ; ++q)
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
            const int idx = lane + 32 * c;
#pragma unroll
            for (int i = 0; i < 8; ++i) qr[q][c][i] = idx < nchunks ? qn[(size_t)q * qn_ld + idx * 8 + i] : 0.f;
        }

    for (long long r0 = gw * 2; r0 < n_rows; r0 += nw * 2) {
        const bool two = r0 + 1 < n_rows;
        uint4 a[2][CPL];
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
            const int idx = lane + 32 * c;
            a[0][c] = make_uint4(0, 0, 0, 0);
            a[1][c] = make_uint4(0, 0, 0, 0);
            if (idx < nchunks) {
                a[0][c] = __ldcs(db + tiled_u4(r0, idx, nk));
                if (two) a[1][c] = __ldcs(db + tiled_u4(r0 + 1, idx, nk));
            }
        }
        float acc[2][NQ];
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int q = 0; q < NQ; ++q) acc[r][q] = 0.f;
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int c = 0; c < CPL; ++c) {
                const uint32_t w[4] = {a[r][c].x, a[r][c].y, a[r][c].z, a[r][c].w};
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                    const float lo = __uint_as_float(w[h] << 16);
                    const float hi = __uint_as_float(w[h] & 0xFFFF0000u);
#pragma unroll
                    for (int q = 0; q < NQ; ++q) {
                        acc[r][q] = fmaf(lo, qr[q][c][2 * h], acc[r][q]);
                        acc[r][q] = fmaf(hi, qr[q][c][2 * h + 1], acc[r][q]);
                    }
                }
            }
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                float v = acc[r][q];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
                acc[r][q] = v;
            }
        if (lane == 0) {
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                out[(size_t)q * out_ld + r0] = acc[0][q];
                if (two) out[(size_t)q * out_ld + r0 + 1] = acc[1][q];
            }
        }
    }
}

// Any row length: queries live in shared memory (slower; rows longer than 2048 elements only).
__global__ void __launch_bounds__(256) scan_small_generic_kernel(const uint4* __restrict__ db, long long n_rows,
                                                                 int nk, int nchunks, int nq,
                                                                 const float* __restrict__ qn, long long qn_ld,
                                                                 float* __restrict__ out, long long out_ld) {
    extern __shared__ float sq[];  // [nq][nchunks*8]
    const int dq = nchunks * 8;
    for (int i = threadIdx.x; i < nq * dq; i += blockDim.x) sq[i] = qn[(size_t)(i / dq) * qn_ld + (i % dq)];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long r = gw; r < n_rows; r += nw) {
        float acc[RVO_SMALL_Q];
#pragma unroll
        for (int q = 0; q < RVO_SMALL_Q; ++q) acc[q] = 0.f;
        for (int c = lane; c < nchunks; c += 32) {
            const uint4 v = __ldcs(db + tiled_u4(r, c, nk));
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int q = 0; q < RVO_SMALL_Q; ++q)
                if (q < nq) {
                    const float* qq = sq + q * dq + c * 8;
#pragma unroll
                    for (int h = 0; h < 4; ++h) {
                        acc[q] = fmaf(__uint_as_float(w[h] << 16), qq[2 * h], acc[q]);
                        acc[q] = fmaf(__uint_as_float(w[h] & 0xFFFF0000u), qq[2 * h + 1], acc[q]);
                    }
                }
        }
#pragma unroll
        for (int q = 0; q < RVO_SMALL_Q; ++q) {
            float v = acc[q];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
            if (lane == 0 && q < nq) out[(size_t)q * out_ld + r] = v;
        }
    }
}

template <int NQ, int CPL>
static int launch_small_t(const uint16_t* db, long long n_rows, int d_pad, const float* qn,
                          long long qn_ld, float* out, long long out_ld, int sm_count, cudaStream_t stream) {
    int per_sm = 0;
    RVO_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, scan_small_kernel<NQ, CPL>, 256, 0));
    if (per_sm < 1) per_sm = 1;
    long long want = (n_rows + 15) / 16;  // 8 warps x 2 rows per block per iteration
    long long grid = (long long)sm_count * per_sm;
    if (grid > want) grid = want;
    RVO_CUDA(launch_pdl(scan_small_kernel<NQ, CPL>, dim3((unsigned)grid), dim3(256), 0, stream, (const uint4*)db, n_rows,
                        d_pad / kTileCols, d_pad / 8, qn, qn_ld, out, out_ld));
    RVO_LAUNCHED();
    return RVO_OK;
}

// nq in 1..RVO_SMALL_Q; qn rows beyond nq (up to the instantiated NQ) must exist and be zero.
int launch_scan_small(const uint16_t* db, long long n_rows, int d_pad, const float* qn, long long qn_ld,
                      int nq, float* out, long long out_ld, int sm_count, cudaStream_t stream) {
    if (n_rows <= 0) return RVO_OK;
    const int cpl = (d_pad / 8 + 31) / 32;
#define RVO_SMALL_CASE(NQ_, CPL_)                                                                          \
    return launch_small_t<NQ_, CPL_>(db, n_rows, d_pad, qn, qn_ld, out, out_ld, sm_count, stream)
    if (cpl <= 4) {
        if (nq == 1) RVO_SMALL_CASE(1, 4);
        if (nq == 2) RVO_SMALL_CASE(2, 4);
        RVO_SMALL_CASE(4, 4);
    }
    if (cpl == 5) {
        if (nq == 1) RVO_SMALL_CASE(1, 5);
        if (nq == 2) RVO_SMALL_CASE(2, 5);
        RVO_SMALL_CASE(4, 5);
    }
#undef RVO_SMALL_CASE
    // long rows: generic kernel
    const size_t smem = (size_t)nq * d_pad * 4;
    if (smem > 200 * 1024) {
        set_error("scan_small: row length %d too large", d_pad);
        return RVO_E_INVALID;
    }
    if (smem > 48 * 1024)
        RVO_CUDA(cudaFuncSetAttribute(scan_small_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long grid = (long long)sm_count * 4;
    long long want = (n_rows + 7) / 8;
    if (grid > want) grid = want;
    scan_small_generic_kernel<<<(unsigned)grid, 256, smem, stream>>>((const uint4*)db, n_rows, d_pad / kTileCols, d_pad / 8,
                                                                    nq, qn, qn_ld, out, out_ld);
    RVO_LAUNCHED();
    return RVO_OK;
}

}  // namespace rvo
