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
                                                             uint4* __restrict__ zero_base, long long zero_u4,
        unsigned long long* tr) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    chain_stamp(tr, 0, false);
    // launched behind the previous search's select (option "pdl"): nothing below may be written before that grid has completed
    // (it still reads the counters and margins this launch resets); the seed scan that follows may pre-launch in turn
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
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
        chain_stamp(tr, 0, true);
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
    chain_stamp(tr, 0, true);
}

// Register-resident variant for d_pad <= 32 * NV: the row is read ONCE — NV independent loads per lane, all in flight together —
// and stays in registers for the magnitude, norm, rounding-error and store passes.  The generic kernel above makes four dependent
// passes over global memory (7 us for a 256 x 1024 query batch; this one is bounded by a single DRAM round trip).  Lane l owns
// elements l, l + 32, ... and accumulates them in that order, exactly like the generic kernel: both produce the same bits, so
// a query normalised through either (any alignment, any d) ranks identically.
template <int NV>
__global__ void __launch_bounds__(256) normalize_rows_reg_kernel(const float* __restrict__ src, long long n, int d,
                                                                 long long src_ld, uint16_t* __restrict__ dst_bf16,
                                                                 long long dst_ld, long long tiled_row0,
                                                                 float* __restrict__ dst_f32, long long f32_ld,
                                                                 float* __restrict__ margin_out, long long n_pad_rows,
                                                                 uint4* __restrict__ zero_base, long long zero_u4,
        unsigned long long* tr) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    chain_stamp(tr, 0, false);
    float v[NV];
    const bool live = row < n;
    if (live) {   // the source rows are never written by a kernel of the chain: they may be read before the dependency wait
        const float* s = src + (size_t)row * (size_t)src_ld;
#pragma unroll
        for (int j = 0; j < NV; ++j) v[j] = lane + 32 * j < d ? __ldg(s + lane + 32 * j) : 0.f;
    }
    // launched behind the previous search's select (option "pdl"): no write before that grid has completed (it still reads the
    // counters and margins this launch resets); the seed scan that follows may pre-launch in turn
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < zero_u4; i += (long long)gridDim.x * blockDim.x)
        zero_base[i] = make_uint4(0, 0, 0, 0);
    if (!live) {
        if (row < n_pad_rows && tiled_row0 < 0) {
            if (dst_bf16)
                for (long long i = lane; i < dst_ld; i += 32) dst_bf16[(size_t)row * (size_t)dst_ld + i] = 0;
            if (dst_f32)
                for (long long i = lane; i < f32_ld; i += 32) dst_f32[(size_t)row * (size_t)f32_ld + i] = 0.f;
            if (margin_out && lane == 0) margin_out[row] = 0.f;
        }
        chain_stamp(tr, 0, true);
        return;
    }
    float amax = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) amax = fmaxf(amax, fabsf(v[j]));   // padding elements are 0: no effect
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xFFFFFFFFu, amax, o));
    const bool rescale = amax > 1e18f || (amax < 1e-18f && amax > 0.f);
    const float pre = rescale ? 1.0f / amax : 1.0f;
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const float x = v[j] * pre;
        ss = fmaf(x, x, ss);                                          // fma(0, 0, ss) == ss for the padding elements
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xFFFFFFFFu, ss, o);
    const float nrm = sqrtf(ss);
    const bool usable = nrm != 0.f && isfinite(nrm);
    const float inv = usable ? pre / nrm : 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) v[j] = (usable && lane + 32 * j < d) ? v[j] * inv : 0.f;   // select: NaN * 0 would stay NaN
    if (dst_bf16) {
        const int nk = (int)(dst_ld / kTileCols);
        float e2 = 0.f;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const int c = lane + 32 * j;
            if (c < dst_ld) {
                const __nv_bfloat16 b = __float2bfloat16_rn(v[j]);
                const float e = __bfloat162float(b) - v[j];
                e2 = fmaf(e, e, e2);
                if (tiled_row0 < 0) dst_bf16[(size_t)row * (size_t)dst_ld + c] = __bfloat16_as_ushort(b);
                else dst_bf16[tiled_offset(tiled_row0 + row, c, nk)] = __bfloat16_as_ushort(b);
            }
        }
        if (margin_out && tiled_row0 < 0) {
#pragma unroll
            for (int o2 = 16; o2 > 0; o2 >>= 1) e2 += __shfl_xor_sync(0xFFFFFFFFu, e2, o2);
            if (lane == 0) margin_out[row] = query_margin(sqrtf(e2));
        }
    }
    if (dst_f32) {
        float* o = dst_f32 + (size_t)row * (size_t)f32_ld;
#pragma unroll
        for (int j = 0; j < NV; ++j)
            if (lane + 32 * j < f32_ld) o[lane + 32 * j] = v[j];
    }
    chain_stamp(tr, 0, true);
}

int launch_normalize_rows(const float* src, long long n, int d, long long src_ld, uint16_t* dst_bf16, long long dst_ld,
                          long long tiled_row0, float* dst_f32, long long f32_ld, cudaStream_t stream,
                          float* margin_out, long long n_pad_rows, void* zero_base, size_t zero_bytes) {
    if (n_pad_rows < n) n_pad_rows = n;
    if (n_pad_rows <= 0 && zero_bytes == 0) return RVO_OK;
    long long blocks = (n_pad_rows + 7) / 8;
    if (blocks < 1) blocks = 1;
    // rows of up to 2048 (padded) elements stay in registers; both kernels give the same bits
    const long long widest = dst_ld > f32_ld ? (dst_ld > d ? dst_ld : d) : (f32_ld > d ? f32_ld : d);
    auto kern = widest > 2048 ? normalize_rows_kernel : widest <= 1024 ? normalize_rows_reg_kernel<32> : normalize_rows_reg_kernel<64>;
    RVO_CUDA(launch_pdl(kern, dim3((unsigned)blocks), dim3(256), 0, stream, src, n, d, src_ld, dst_bf16, dst_ld, tiled_row0, dst_f32,
                        f32_ld, margin_out, n_pad_rows, (uint4*)zero_base, (long long)(zero_bytes / 16),
                        (unsigned long long*)(uintptr_t)g_chain_trace.load()));
    RVO_LAUNCHED();
    return RVO_OK;
}

// ---- small-Q scan --------------------------------------------------------------------------------
// One warp per pair of DB rows; lane l owns the 16-byte chunks l, l+32, ... of a row, and keeps the
// matching slices of all NQ queries in registers (no shared-memory traffic in the loop).
//   NORM    the kernel normalises the raw queries itself (every CTA redundantly: NQ <= 4 rows of L2-resident floats; block 0
//           also publishes them for the kernels that follow) — no separate normalise launch on the reference's Q = 1 path
//   DENSE   scores of every `pair_stride`-th row pair go to out[q][2*j + {0,1}] (pair_stride = 1: every row)
//   FILTER  rows whose ordering key (score, row) reaches tau_key[q] — the k-th best key of a row sample, a lower bound of the
//           k-th best key of the shard — are appended to cand[q][seg][cap]; scores never reach HBM
// The arithmetic of a row's score is identical in both modes (same lanes, same order), so a sampled row passes its own filter.
// uint4 index of the 16-byte chunk `idx` (0 .. d_pad/8) of row `r` in the tiled DB storage
__device__ __forceinline__ size_t tiled_u4(long long r, int idx, int nk) {
    return (((size_t)(r >> 7) * (size_t)nk + (size_t)(idx >> 3)) * kTileRows + (size_t)(r & 127)) * 8 + (size_t)(idx & 7);
}

template <int NQ, int CPL, bool DENSE, bool NORM>
__global__ void __launch_bounds__(256) scan_small_kernel(const SmallScanArgs a) {
    __shared__ float s_inv[RVO_SMALL_Q];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    if (NORM) {
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.zero_u4; i += (long long)gridDim.x * blockDim.x)
            a.zero_base[i] = make_uint4(0, 0, 0, 0);
        // x / ||x|| (core_system.py:407 / qdrant COSINE search), same rules as normalize_rows_kernel: zero and non-finite
        // queries become the zero vector; the sum of squares is scaled by the largest magnitude against overflow
        if (warp < NQ) {
            float inv = 0.f;
            if (warp < a.nq) {
                const float* s = a.q + (size_t)warp * (size_t)a.q_ld;
                float amax = 0.f;
                for (int i = lane; i < a.d; i += 32) amax = fmaxf(amax, fabsf(s[i]));
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xFFFFFFFFu, amax, o));
                const bool rescale = amax > 1e18f || (amax < 1e-18f && amax > 0.f);
                const float pre = rescale ? 1.0f / amax : 1.0f;
                float ss = 0.f;
                for (int i = lane; i < a.d; i += 32) {
                    const float v = s[i] * pre;
                    ss = fmaf(v, v, ss);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xFFFFFFFFu, ss, o);
                const float nrm = sqrtf(ss);
                inv = (nrm != 0.f && isfinite(nrm)) ? pre / nrm : 0.f;
            }
            if (lane == 0) s_inv[warp] = inv;
        }
        __syncthreads();
        if (blockIdx.x == 0 && a.qn_out) {
            for (int i = threadIdx.x; i < RVO_SMALL_Q * (int)a.qn_ld; i += blockDim.x) {
                const int q = i / (int)a.qn_ld, c = i - q * (int)a.qn_ld;
                const float inv = q < NQ ? s_inv[q] : 0.f;
                a.qn_out[i] = (q < a.nq && c < a.d && inv != 0.f) ? a.q[(size_t)q * (size_t)a.q_ld + c] * inv : 0.f;
            }
        }
    }

    float qr[NQ][CPL][8];
#pragma unroll
    for (int q = 0; q < NQ; ++q)
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
            const int idx = lane + 32 * c;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float v = 0.f;
                if (idx < a.nchunks) {
                    if (NORM) {
                        const int col = idx * 8 + i;
                        const float inv = s_inv[q];
                        v = (q < a.nq && col < a.d && inv != 0.f) ? a.q[(size_t)q * (size_t)a.q_ld + col] * inv : 0.f;
                    } else {
                        v = a.q[(size_t)q * (size_t)a.q_ld + idx * 8 + i];
                    }
                }
                qr[q][c][i] = v;
            }
        }

    unsigned long long tk[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) tk[q] = (!DENSE && q < a.nq) ? a.tau_key[q] : ~0ull;
    const long long npairs = (a.n_rows + 1) >> 1;
    const long long stride = DENSE ? a.pair_stride : 1;
    const long long nsample = (npairs + stride - 1) / stride;
    const int seg = DENSE ? 0 : (int)(gw % a.nseg);

    for (long long j = gw; j < nsample; j += nw) {
        const long long r0 = j * stride * 2;
        const bool two = r0 + 1 < a.n_rows;
        uint4 v[2][CPL];
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
            const int idx = lane + 32 * c;
            v[0][c] = make_uint4(0, 0, 0, 0);
            v[1][c] = make_uint4(0, 0, 0, 0);
            if (idx < a.nchunks) {
                v[0][c] = __ldcs(a.db + tiled_u4(r0, idx, a.nk));
                if (two) v[1][c] = __ldcs(a.db + tiled_u4(r0 + 1, idx, a.nk));
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
                const uint32_t w[4] = {v[r][c].x, v[r][c].y, v[r][c].z, v[r][c].w};
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
                float x = acc[r][q];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xFFFFFFFFu, x, o);
                acc[r][q] = x;
            }
        if (lane == 0) {
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                if (q >= a.nq) break;
                if (DENSE) {
                    a.out[(size_t)q * a.out_ld + 2 * j] = acc[0][q];
                    if (two) a.out[(size_t)q * a.out_ld + 2 * j + 1] = acc[1][q];
                } else {
#pragma unroll
                    for (int r = 0; r < 2; ++r) {
                        if (r == 1 && !two) break;
                        const unsigned long long key = make_key(acc[r][q], (uint32_t)(r0 + r));
                        if (key >= tk[q]) {
                            const int pos = atomicAdd(a.cnt + q * a.nseg + seg, 1);
                            if (pos < a.cap) a.cand[((size_t)q * a.nseg + seg) * (size_t)a.cap + pos] = key;
                        }
                    }
                }
            }
        }
    }
}

template <int NQ, int CPL>
static int launch_small_t(const SmallScanArgs& a, bool dense, bool norm, int sm_count, cudaStream_t stream) {
    const long long npairs = (a.n_rows + 1) / 2;
    const long long nsample = dense ? (npairs + a.pair_stride - 1) / a.pair_stride : npairs;
    long long want = (nsample + 7) / 8;  // 8 warps, one row pair each, per block per iteration
    if (want < 1) want = 1;
#define RVO_SMALL_LAUNCH(D_, N_)                                                                                     \
    do {                                                                                                             \
        int per_sm = 0;                                                                                              \
        RVO_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, scan_small_kernel<NQ, CPL, D_, N_>, 256, 0)); \
        if (per_sm < 1) per_sm = 1;                                                                                  \
        long long grid = (long long)sm_count * per_sm;                                                               \
        if (grid > want) grid = want;                                                                                \
        RVO_CUDA(launch_pdl_small(scan_small_kernel<NQ, CPL, D_, N_>, dim3((unsigned)grid), dim3(256), 0, stream, a)); \
    } while (0)
    if (dense && norm) RVO_SMALL_LAUNCH(true, true);
    else if (dense) RVO_SMALL_LAUNCH(true, false);
    else RVO_SMALL_LAUNCH(false, false);
#undef RVO_SMALL_LAUNCH
    RVO_LAUNCHED();
    return RVO_OK;
}

// Rows longer than 2048 elements are not instantiated: the tensor path serves them (make_plan).
int launch_scan_small(const SmallScanArgs& a, bool dense, bool norm, int sm_count, cudaStream_t stream) {
    if (a.n_rows <= 0) return RVO_OK;
    const int cpl = (a.nchunks + 31) / 32;
#define RVO_SMALL_CASE(NQ_, CPL_) return launch_small_t<NQ_, CPL_>(a, dense, norm, sm_count, stream)
    if (cpl <= 4) {
        if (a.nq == 1) RVO_SMALL_CASE(1, 4);
        if (a.nq == 2) RVO_SMALL_CASE(2, 4);
        RVO_SMALL_CASE(4, 4);
    }
    if (cpl == 5) {
        if (a.nq == 1) RVO_SMALL_CASE(1, 5);
        if (a.nq == 2) RVO_SMALL_CASE(2, 5);
        RVO_SMALL_CASE(4, 5);
    }
    if (cpl <= 8) {   // long rows (d_pad <= 2048): correct, register-heavy (rare configuration)
        if (a.nq == 1) RVO_SMALL_CASE(1, 8);
        if (a.nq == 2) RVO_SMALL_CASE(2, 8);
        RVO_SMALL_CASE(4, 8);
    }
#undef RVO_SMALL_CASE
    set_error("scan_small: row length %d not covered by the fp32 scan (d_pad <= 2048)", a.nchunks * 8);
    return RVO_E_UNSUPPORTED;
}

bool scan_small_supports(int d_pad) { return d_pad <= 2048; }

}  // namespace rvo
