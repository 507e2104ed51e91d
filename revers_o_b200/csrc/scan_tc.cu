// K2 — query x database scoring on tcgen05 tensor cores with the threshold select fused in the
// epilogue.  Replaces the `vectors @ query` + argsort of qdrant-local search behind
// core_system.py:659-664 for batches of more than RVO_SMALL_Q queries.
//
// Roles inside one persistent CTA (320 threads, one CTA per SM):
//   warp 0 (one elected lane)  TMA producer: DB sub-tiles (128 rows x 64 k, 128B-swizzled) and, unless
//                              the query block is resident, the query k-chunk, into a smem stage ring
//   warp 1 (one elected lane)  tcgen05.mma issuer: D[128 rows x NQ] += A(db) * B(queries)^T, fp32
//                              accumulators in TMEM slots (ring of 512/NQ slots)
//   warps 2-9                  epilogue, two warps per TMEM lane quarter (interleaved 16-column chunks):
//                              tcgen05.ld one TMEM lane (= one DB row) per thread,
//                                FILTER: compare the row's scores with the per-query thresholds tau
//                                        (smem, 128-bit broadcast loads), queue survivors in smem, then
//                                        append them as ordering keys to the per-query candidate lists;
//                                DENSE : write the scores (threshold-seeding pass over sampled tiles).
// The full score matrix never reaches HBM in FILTER mode.
//
// Layout choice: DB rows are the MMA M dimension (TMEM lanes), queries the N dimension
// (TMEM columns), so that (a) a small query batch costs no padded smem traffic and can stay RESIDENT
// in smem for the whole scan (HBM-bound regime: smem fill == DB bytes), and (b) `m_sub` DB sub-tiles
// share one query k-chunk per stage (tensor-bound regime: halves the L2->smem query traffic).
#include "common.cuh"
#include "ptx.cuh"
#include "scan_tc.cuh"

namespace rvo {

using namespace ptx;

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory"); }

template <int MODE>
__global__ void __launch_bounds__(kScanThreads, 1)
scan_tc_kernel(const __grid_constant__ CUtensorMap tmap_db, const __grid_constant__ CUtensorMap tmap_q,
               const ScanParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];  // no static smem in this kernel: base is 1024-aligned

    uint8_t* s_res = smem;                                   // resident query block (optional)
    uint8_t* s_stages = smem + p.off_stages;                 // stage ring
    unsigned long long* s_queue = (unsigned long long*)(smem + p.off_queue);  // [kQueueCap][kEpiThreads]
    float* s_tau = (float*)(smem + p.off_tau);               // [2][256] double-buffered by work item
    uint64_t* s_bars = (uint64_t*)(smem + p.off_bars);
    uint64_t* bar_full = s_bars;                             // [kMaxStages] TMA -> MMA
    uint64_t* bar_empty = s_bars + kMaxStages;               // [kMaxStages] MMA -> TMA
    uint64_t* bar_tfull = s_bars + 2 * kMaxStages;           // [kMaxSlots]  MMA -> epilogue
    uint64_t* bar_tempty = s_bars + 2 * kMaxStages + kMaxSlots;  // [kMaxSlots] epilogue -> MMA
    uint64_t* bar_res = s_bars + 2 * kMaxStages + 2 * kMaxSlots;
    uint32_t* s_tmem = (uint32_t*)(bar_res + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_k = p.d_pad / kBlockK;
    const int nq_blk = p.nq_blk;
    const int m_sub = p.m_sub;
    const int tile_rows = kBlockM * m_sub;
    const uint32_t q_chunk_bytes = (uint32_t)nq_blk * 128u;
    const long long total_work = p.num_super * (long long)p.num_qblk;

    chain_stamp(p.trace, MODE == kModeDense ? 1 : 3, false);
    if (threadIdx.x == 0) {
        if (smem_u32(smem) & 1023u) {
            printf("rvo: dynamic shared memory base not 1024-byte aligned\n");
            __trap();
        }
        for (int i = 0; i < kMaxStages; ++i) {
            mbar_init(&bar_full[i], 1);
            mbar_init(&bar_empty[i], 1);
        }
        for (int i = 0; i < kMaxSlots; ++i) {
            mbar_init(&bar_tfull[i], 1);
            mbar_init(&bar_tempty[i], kEpiWarps);  // lane 0 of each epilogue warp
        }
        mbar_init(bar_res, 1);
        fence_mbar_init();
        fence_proxy_async();
    }
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_db);
        prefetch_tmap(&tmap_q);
    }
    if (warp == 1) {
        tmem_alloc(s_tmem, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;
    grid_dependency_wait();      // thresholds / queries / counters come from the previous kernels of the chain
    grid_launch_dependents();

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (elect_one()) {
            const uint64_t hint_db = (p.num_qblk > 1) ? kEvictNormal : kEvictFirst;
            if (p.resident_q) {
                mbar_expect_tx(bar_res, (uint32_t)num_k * q_chunk_bytes);
                for (int kc = 0; kc < num_k; ++kc)
                    tma_load_2d(&tmap_q, bar_res, s_res + (size_t)kc * q_chunk_bytes, kc * kBlockK, 0, kEvictLast);
            }
            uint32_t stage = 0, phase = 0;
            for (long long w = blockIdx.x; w < total_work; w += gridDim.x) {
                const long long st = w / p.num_qblk;
                const int qb = (int)(w - st * p.num_qblk);
                const long long row0 = st * p.super_stride * tile_rows;
                long long left = (p.n_rows - row0 + kBlockM - 1) / kBlockM;
                const int m_valid = left < m_sub ? (int)left : m_sub;
                const uint32_t bytes = (uint32_t)m_valid * kSubTileBytes + (p.resident_q ? 0u : q_chunk_bytes);
                for (int kc = 0; kc < num_k; ++kc) {
                    mbar_wait(&bar_empty[stage], phase ^ 1);
                    mbar_expect_tx(&bar_full[stage], bytes);
                    uint8_t* sA = s_stages + (size_t)stage * p.stage_bytes;
                    for (int j = 0; j < m_valid; ++j)
                        // tiled DB: tile (row block, k-chunk) is one contiguous 16 KiB box of 128 "virtual rows"
                        tma_load_2d(&tmap_db, &bar_full[stage], sA + j * kSubTileBytes, 0,
                                    (int)(((row0 / kBlockM + j) * num_k + kc) * kBlockM), hint_db);
                    if (!p.resident_q)
                        tma_load_2d(&tmap_q, &bar_full[stage], sA + m_sub * kSubTileBytes, kc * kBlockK, qb * nq_blk,
                                    kEvictLast);
                    if (++stage == (uint32_t)p.num_stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (elect_one()) {
            const uint32_t idesc = make_idesc_bf16(kBlockM, (uint32_t)nq_blk);
            const uint32_t ns = (uint32_t)p.num_slots;
            if (p.resident_q) {
                mbar_wait(bar_res, 0);
                tc_fence_after();
            }
            uint32_t stage = 0, phase = 0, acc = 0;
            for (long long w = blockIdx.x; w < total_work; w += gridDim.x) {
                const long long st = w / p.num_qblk;
                const long long row0 = st * p.super_stride * tile_rows;
                long long left = (p.n_rows - row0 + kBlockM - 1) / kBlockM;
                const int m_valid = left < m_sub ? (int)left : m_sub;
                for (int kc = 0; kc < num_k; ++kc) {
                    mbar_wait(&bar_full[stage], phase);
                    tc_fence_after();
                    const uint32_t sA = smem_u32(s_stages + (size_t)stage * p.stage_bytes);
                    const uint32_t sB = p.resident_q ? smem_u32(s_res + (size_t)kc * q_chunk_bytes)
                                                     : sA + (uint32_t)m_sub * kSubTileBytes;
                    for (int j = 0; j < m_valid; ++j) {
                        const uint32_t pos = acc + (uint32_t)j;
                        const uint32_t slot = pos % ns;
                        if (kc == 0) {
                            mbar_wait(&bar_tempty[slot], ((pos / ns) & 1u) ^ 1u);
                            tc_fence_after();
                        }
                        const uint32_t d_tmem = tmem_base + slot * (uint32_t)p.slot_w;
                        const uint32_t a0 = sA + (uint32_t)j * kSubTileBytes;
#pragma unroll
                        for (int k4 = 0; k4 < kBlockK / 16; ++k4)
                            umma_bf16(d_tmem, make_kmajor_sw128_desc(a0 + k4 * 32), make_kmajor_sw128_desc(sB + k4 * 32),
                                      idesc, (uint32_t)((kc | k4) != 0));
                    }
                    umma_commit(&bar_empty[stage]);  // frees the smem stage once these MMAs retire
                    if (++stage == (uint32_t)p.num_stages) { stage = 0; phase ^= 1; }
                }
                for (int j = 0; j < m_valid; ++j) umma_commit(&bar_tfull[(acc + (uint32_t)j) % ns]);
                acc += (uint32_t)m_valid;
            }
        }
    } else {
        // ===================== epilogue (warps 2..9) =====================
        const int ew = warp - 2;                         // 0..7
        const uint32_t lane_base = (uint32_t)(warp & 3) * 32u;  // TMEM lane quarter this warp may read
        const int half = ew >> 2;                        // which interleaved set of 16-column chunks
        const int et = (int)lane_base + lane;            // 0..127: TMEM lane == DB row inside the sub-tile
        const int qt = half * kBlockM + et;              // 0..255: private queue column
        const uint32_t ns = (uint32_t)p.num_slots;
        const int nchunks = nq_blk / 16;
        uint32_t acc = 0;
        uint32_t wcount = 0;
        int cur_qb = -1;
        // candidate appends in flight from the previous TMEM slot (FILTER mode)
        int pend_n = 0, pend_q0 = 0;
        uint32_t pend_row = 0;
        int pend_pos[4] = {0, 0, 0, 0};
        unsigned long long pend_ent[4] = {0ull, 0ull, 0ull, 0ull};

        for (long long w = blockIdx.x; w < total_work; w += gridDim.x, ++wcount) {
            const long long st = w / p.num_qblk;
            const int qb = (int)(w - st * p.num_qblk);
            const long long row0 = st * p.super_stride * tile_rows;
            long long left = (p.n_rows - row0 + kBlockM - 1) / kBlockM;
            const int m_valid = left < m_sub ? (int)left : m_sub;
            const int q0 = qb * nq_blk;
            // sub-list of every query this super-tile appends to: a function of the ROW range only, so that the
            // 16 sub-lists of a query fill evenly whatever the CTA <-> query-block assignment is
            const int seg = (int)(st % kCandSplit);
            const uint32_t tbuf = p.num_qblk > 1 ? (wcount & 1u) : 0u;
            const float* tau_s = s_tau + tbuf * 256;
            const float* tauh_s = s_tau + 512 + tbuf * 256;
            const int hot_seg = kCandSplit + (int)(st % kHotSplit);   // hot sub-list of this super-tile (search path only)

            if (MODE == kModeFilter && qb != cur_qb) {
                // double-buffered by work item; the barrier also orders "everyone left work w-1" before
                // anyone overwrites that buffer at work w+1
                const int i = ew * 32 + lane;
                if (i < nq_blk) {
                    s_tau[tbuf * 256 + i] = p.tau[q0 + i];
                    s_tau[512 + tbuf * 256 + i] = p.tau_hot ? p.tau_hot[q0 + i] : __int_as_float(0x7f800000);
                }
                epi_bar_sync();
                cur_qb = p.num_qblk > 1 ? -1 : qb;       // several query blocks: reload every work item
            }

            for (int j = 0; j < m_valid; ++j) {
                const uint32_t pos = acc + (uint32_t)j;
                const uint32_t slot = pos % ns;
                mbar_wait(&bar_tfull[slot], (pos / ns) & 1u);
                tc_fence_after();
                const uint32_t taddr = tmem_base + (lane_base << 16) + slot * (uint32_t)p.slot_w;
                const long long row = row0 + (long long)j * kBlockM + et;
                const bool valid = row < p.n_rows;

                if (MODE == kModeDense) {
                    const long long col = st * tile_rows + (long long)j * kBlockM + et;  // sample index
                    for (int c = half; c < nchunks; c += 2) {
                        uint32_t v[16];
                        tmem_ld_x16(taddr + (uint32_t)c * 16u, v);
                        tmem_ld_wait();
                        if (p.dense_max) {   // warp-uniform: one maximum per (query, 32 sample rows)
                            float f[16];
#pragma unroll
                            for (int i = 0; i < 16; ++i) f[i] = valid ? __uint_as_float(v[i]) : -__int_as_float(0x7f800000);
                            const float m = warp_colmax16(f, lane);
                            if ((lane & 1) == 0)
                                p.dense[(size_t)(q0 + c * 16 + ((lane >> 1) & 15)) * (size_t)p.dense_ld + (size_t)(col >> 5)] = m;
                            continue;
                        }
                        {   // rows past the end of the DB (zero-filled by TMA) must never rank: -inf
                            float* o = p.dense + (size_t)(q0 + c * 16) * (size_t)p.dense_ld + (size_t)col;
#pragma unroll
                            for (int i = 0; i < 16; ++i)
                                o[(size_t)i * (size_t)p.dense_ld] = valid ? __uint_as_float(v[i]) : -__int_as_float(0x7f800000);
                        }
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bar_tempty[slot]);
                } else {
                    int n = 0;
                    const uint32_t key_row_bits = 0xFFFFFFFFu - (uint32_t)row;
                    // sub-list of a survivor, decided when its queue entry is read back (NOT in the compare loop, which must stay as
                    // lean as the tensor pipe is fast): hot list if the score also reaches tau_hot, else the row-range's safe list
                    auto with_seg = [&](unsigned long long ent) -> unsigned long long {
                        const int sg = __uint_as_float((uint32_t)(ent >> 32)) >= tauh_s[(int)(ent & 0xFFFFu)] ? hot_seg : seg;
                        return ent | ((unsigned long long)sg << 16);
                    };
                    auto flush_sync = [&](int from, int cnt) {
                        for (int e = from; e < cnt; e += 4) {
                            int slot_pos[4];
                            unsigned long long ent[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u)
                                if (e + u < cnt) {
                                    ent[u] = with_seg(s_queue[(e + u) * kEpiThreads + qt]);
                                    slot_pos[u] = atomicAdd(p.cand_cnt + (q0 + (int)(ent[u] & 0xFFFFu)) * p.nseg + (int)((ent[u] >> 16) & 0xFFu), 1);
                                }
#pragma unroll
                            for (int u = 0; u < 4; ++u)
                                if (e + u < cnt && slot_pos[u] < p.cap) {
                                    const int q = q0 + (int)(ent[u] & 0xFFFFu);
                                    const uint32_t ob = f32_orderable(__uint_as_float((uint32_t)(ent[u] >> 32)));
                                    p.cand[((size_t)q * p.nseg + ((ent[u] >> 16) & 0xFFu)) * (size_t)p.cap + (size_t)slot_pos[u]] =
                                        ((unsigned long long)ob << 32) | (unsigned long long)key_row_bits;
                                }
                        }
                    };
                    for (int c = half; c < nchunks; c += 4) {
                        // two 16-column chunks per TMEM wait
                        uint32_t v[2][16];
                        const bool two = c + 2 < nchunks;
                        tmem_ld_x16(taddr + (uint32_t)c * 16u, v[0]);
                        if (two) tmem_ld_x16(taddr + (uint32_t)(c + 2) * 16u, v[1]);
                        tmem_ld_wait();
                        if (valid) {
#pragma unroll
                            for (int g = 0; g < 2; ++g) {
                                if (g == 1 && !two) break;
                                const int cc = c + 2 * g;
#pragma unroll
                                for (int h = 0; h < 2; ++h) {
                                    if (n > kQueueCap - 8) {  // rare: keep room for 8 more
                                        flush_sync(0, n);
                                        n = 0;
                                    }
                                    const float4 t0 = *(const float4*)(tau_s + cc * 16 + h * 8);
                                    const float4 t1 = *(const float4*)(tau_s + cc * 16 + h * 8 + 4);
                                    const float tt[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
#pragma unroll
                                    for (int i = 0; i < 8; ++i) {
                                        if (__uint_as_float(v[g][h * 8 + i]) >= tt[i]) {
                                            s_queue[n * kEpiThreads + qt] = ((unsigned long long)v[g][h * 8 + i] << 32) |
                                                                            (unsigned)(cc * 16 + h * 8 + i);
                                            ++n;
                                        }
                                    }
                                }
                            }
                        }
                        __syncwarp();  // tcgen05.ld is warp-collective: reconverge before the next one
                    }
                    // TMEM slot is drained: hand it back to the MMA warp.
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bar_tempty[slot]);
                    // Software-pipelined append: the atomics issued for the PREVIOUS slot have had a whole slot of
                    // time to return; finish its stores now, then issue this slot's atomics and move on without
                    // waiting for them (their ~3 us round trip is hidden behind the next slot's scan).
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (u < pend_n && pend_pos[u] < p.cap) {
                            const int q = pend_q0 + (int)(pend_ent[u] & 0xFFFFu);
                            const uint32_t ob = f32_orderable(__uint_as_float((uint32_t)(pend_ent[u] >> 32)));
                            p.cand[((size_t)q * p.nseg + ((pend_ent[u] >> 16) & 0xFFu)) * (size_t)p.cap + (size_t)pend_pos[u]] =
                                ((unsigned long long)ob << 32) | (unsigned long long)pend_row;
                        }
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (u < n) {
                            pend_ent[u] = with_seg(s_queue[u * kEpiThreads + qt]);
                            pend_pos[u] = atomicAdd(p.cand_cnt + (q0 + (int)(pend_ent[u] & 0xFFFFu)) * p.nseg + (int)((pend_ent[u] >> 16) & 0xFFu), 1);
                        }
                    pend_n = n < 4 ? n : 4;
                    pend_q0 = q0;
                    pend_row = key_row_bits;
                    if (n > 4) flush_sync(4, n);
                }
            }
            acc += (uint32_t)m_valid;
        }
        if (MODE == kModeFilter) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (u < pend_n && pend_pos[u] < p.cap) {
                    const int q = pend_q0 + (int)(pend_ent[u] & 0xFFFFu);
                    const uint32_t ob = f32_orderable(__uint_as_float((uint32_t)(pend_ent[u] >> 32)));
                    p.cand[((size_t)q * p.nseg + ((pend_ent[u] >> 16) & 0xFFu)) * (size_t)p.cap + (size_t)pend_pos[u]] =
                        ((unsigned long long)ob << 32) | (unsigned long long)pend_row;
                }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
    chain_stamp(p.trace, MODE == kModeDense ? 1 : 3, true);
}

// ------------------------------------------------------------------------------------------------
// Host side: planning, tensor maps, launch
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* f = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)f;
    }
    return fn;
}

// bf16 [rows, d_pad] row-major matrix; box = 64 elements x box_rows, 128-byte swizzle.
static int make_tmap(CUtensorMap* m, const void* base, long long rows, int d_pad, long long pitch_elems,
                     int box_rows, bool promote) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
        return RVO_E_NO_DEVICE;
    }
    cuuint64_t gdim[2] = {(cuuint64_t)d_pad, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)pitch_elems * 2ull};
    cuuint32_t box[2] = {(cuuint32_t)kBlockK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     promote ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) rows=%lld d_pad=%d pitch=%lld box_rows=%d", (int)r, rows, d_pad,
                  pitch_elems, box_rows);
        return RVO_E_CUDA;
    }
    return RVO_OK;
}

int make_scan_tmap(CUtensorMap* m, const void* base, long long rows, int d_pad, long long pitch_elems, int box_rows) {
    return make_tmap(m, base, rows, d_pad, pitch_elems, box_rows, false);
}

int plan_scan_tc(int nq, int d_pad, int force_m_sub, TcPlan* pl) {
    if (nq < 1 || d_pad < kBlockK || d_pad % kBlockK != 0) {
        set_error("plan_scan_tc: bad nq=%d d_pad=%d", nq, d_pad);
        return RVO_E_INVALID;
    }
    const int nq16 = (nq + 15) / 16 * 16;
    if (nq16 <= 256) {
        pl->num_qblk = 1;
        pl->nq_blk = nq16;
    } else {
        pl->num_qblk = (nq + 255) / 256;
        const int per = (nq + pl->num_qblk - 1) / pl->num_qblk;
        pl->nq_blk = (per + 15) / 16 * 16;
    }
    pl->nq_pad = pl->num_qblk * pl->nq_blk;
    pl->slot_w = (pl->nq_blk + 31) / 32 * 32;
    pl->num_slots = 512 / pl->slot_w;
    if (pl->num_slots > kMaxSlots) pl->num_slots = kMaxSlots;

    const size_t fixed = (size_t)kQueueCap * kEpiThreads * 8 + kTauSmemBytes + 512;
    const size_t budget = (size_t)kSmemLimit - fixed;
    const size_t res_bytes = (size_t)pl->nq_blk * d_pad * 2;

    pl->resident = 0;
    pl->m_sub = 1;
    if (pl->num_qblk == 1 && res_bytes + 3 * (size_t)kSubTileBytes <= budget && force_m_sub <= 0) pl->resident = 1;
    if (!pl->resident) {
        // one query block: TMEM double-buffering (m_sub = 1) beats sharing the query chunk (measured)
        pl->m_sub = pl->num_qblk == 1 ? 1 : 2;
        if (force_m_sub == 1 || force_m_sub == 2 || force_m_sub == 4) pl->m_sub = force_m_sub;
        while (pl->m_sub > pl->num_slots) pl->m_sub >>= 1;
    }
    pl->stage_bytes = (size_t)pl->m_sub * kSubTileBytes + (pl->resident ? 0 : (size_t)pl->nq_blk * 128);
    const size_t res = pl->resident ? res_bytes : 0;
    size_t stages = (budget - res) / pl->stage_bytes;
    if (stages > kMaxStages) stages = kMaxStages;
    if (stages < 2) {
        set_error("plan_scan_tc: not enough shared memory (nq_blk=%d d_pad=%d)", pl->nq_blk, d_pad);
        return RVO_E_INVALID;
    }
    pl->num_stages = (int)stages;
    pl->off_stages = res;  // multiple of 1024: nq_blk % 16 == 0 and d_pad % 64 == 0
    pl->off_queue = pl->off_stages + stages * pl->stage_bytes;
    pl->off_tau = pl->off_queue + (size_t)kQueueCap * kEpiThreads * 8;
    pl->off_bars = pl->off_tau + kTauSmemBytes;
    pl->smem_bytes = pl->off_bars + 512;
    pl->pair = 0;
    // query block too large to stay resident: CTA pairs halve the query operand's L2->SM traffic (scan_tc2.cu)
    if (!pl->resident && force_m_sub <= 0 && pl->nq_blk >= 32) {
        TcPlan p2;
        plan_scan_tc2(*pl, d_pad, &p2);
        p2.pair = 1;
        *pl = p2;
    }
    return RVO_OK;
}

long long scan_tc_sample_rows(long long n_rows, const TcPlan& pl, long long super_stride) {
    const long long T = (long long)kBlockM * pl.m_sub;
    const long long supers = (n_rows + T - 1) / T;
    return (supers + super_stride - 1) / super_stride * T;
}

int launch_scan_tc(int mode, const uint16_t* db, long long n_rows, long long super_stride, int d_pad,
                   const uint16_t* q_bf16, const TcPlan& pl, const float* tau, unsigned long long* cand,
                   int* cand_cnt, int cap, float* dense, long long dense_ld, int sm_count, cudaStream_t stream,
                   const float* tau_hot, int nseg) {
    if (n_rows <= 0) return RVO_OK;
    if (pl.pair)
        return launch_scan_tc2(mode, db, n_rows, super_stride, d_pad, q_bf16, pl, tau, cand, cand_cnt, cap, dense, dense_ld,
                               sm_count, stream, tau_hot, nseg);
    if (n_rows >= (1ll << 31)) {
        set_error("launch_scan_tc: shard too large (%lld rows); shard the DB", n_rows);
        return RVO_E_INVALID;
    }
    if (super_stride < 1) super_stride = 1;
    CUtensorMap tm_db, tm_q;
    // DB storage is tiled [row block of 128][k-chunk][128 rows][64 elements]: as a 2-D tensor of 128-byte
    // "virtual rows" every (row block, k-chunk) box is one contiguous 16 KiB read
    const long long row_blocks = (n_rows + kBlockM - 1) / kBlockM;
    const long long vrows = row_blocks * (d_pad / kBlockK) * kBlockM;
    if (vrows >= (1ll << 31)) {
        set_error("launch_scan_tc: shard too large (%lld rows x %d); shard the DB", n_rows, d_pad);
        return RVO_E_INVALID;
    }
    int rc = make_tmap(&tm_db, db, vrows, kBlockK, kBlockK, kBlockM, false);
    if (rc) return rc;
    rc = make_tmap(&tm_q, q_bf16, pl.nq_pad, d_pad, d_pad, pl.nq_blk, false);
    if (rc) return rc;

    ScanParams p;
    const long long T = (long long)kBlockM * pl.m_sub;
    p.n_rows = n_rows;
    p.super_stride = super_stride;
    p.d_pad = d_pad;
    p.nq_blk = pl.nq_blk;
    p.num_qblk = pl.num_qblk;
    p.m_sub = pl.m_sub;
    p.num_stages = pl.num_stages;
    p.resident_q = pl.resident;
    p.slot_w = pl.slot_w;
    p.num_slots = pl.num_slots;
    p.stage_bytes = (uint32_t)pl.stage_bytes;
    p.off_stages = (uint32_t)pl.off_stages;
    p.off_queue = (uint32_t)pl.off_queue;
    p.off_tau = (uint32_t)pl.off_tau;
    p.off_bars = (uint32_t)pl.off_bars;
    p.num_super = ((n_rows + T - 1) / T + super_stride - 1) / super_stride;
    p.tau = tau;
    p.tau_hot = tau_hot;
    p.cand = cand;
    p.cand_cnt = cand_cnt;
    p.nseg = nseg;
    p.cap = cap;
    p.dense = dense;
    p.dense_ld = dense_ld;
    p.dense_max = mode == kModeDenseMax;
    p.trace = (unsigned long long*)(uintptr_t)g_chain_trace.load();

    const long long total = p.num_super * pl.num_qblk;
    const int grid = (int)(total < sm_count ? total : sm_count);
    if (mode == kModeDense || mode == kModeDenseMax) {
        RVO_CUDA(cudaFuncSetAttribute(scan_tc_kernel<kModeDense>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)pl.smem_bytes));
        RVO_CUDA(launch_pdl(scan_tc_kernel<kModeDense>, dim3(grid), dim3(kScanThreads), pl.smem_bytes, stream, tm_db, tm_q, p));
    } else {
        RVO_CUDA(cudaFuncSetAttribute(scan_tc_kernel<kModeFilter>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)pl.smem_bytes));
        RVO_CUDA(launch_pdl(scan_tc_kernel<kModeFilter>, dim3(grid), dim3(kScanThreads), pl.smem_bytes, stream, tm_db, tm_q, p));
    }
    RVO_LAUNCHED();
    return RVO_OK;
}

}  // namespace rvo
