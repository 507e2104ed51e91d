// K2 scan, CTA-pair variant (tcgen05 cta_group::2) for query blocks that do not fit shared memory.
//
// Two CTAs of a cluster (one SM pair) share each MMA: M = 256 DB rows (128 per CTA), N = the query block.
// Each CTA TMA-loads its own 128 DB rows and only HALF of the query k-chunk; the tensor cores read the other
// half from the peer's shared memory.  Per CTA and k-chunk that is 16 KiB + NQ*64 B instead of 16 KiB + NQ*128 B,
// i.e. the L2->SM traffic of the query operand is halved (3x DB bytes -> 2x at NQ = 256) while every CTA keeps
// two TMEM slots, so the epilogue of tile i still overlaps the MMAs of tile i+1.
//
// Protocol (leader = even CTA of the pair):
//   full[stage]   leader's barrier, 2 arrivals (both producers) + the bytes of all four TMA boxes of the pair
//   empty[stage]  one per CTA, released by the leader's MMA thread with a multicast tcgen05.commit
//   tfull[slot]   one per CTA, multicast commit after the last k-chunk of a tile
//   tempty[slot]  leader's barrier, 16 arrivals: the 8 epilogue warps of BOTH CTAs
// Everything else (thresholds, survivor queue, pipelined candidate append, DENSE mode) is the single-CTA kernel's.
#include "common.cuh"
#include "ptx.cuh"
#include "scan_tc.cuh"

namespace rvo {

using namespace ptx;

__device__ __forceinline__ void epi_bar_sync2() { asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory"); }

template <int MODE>
__global__ void __launch_bounds__(kScanThreads, 1)
scan_tc2_kernel(const __grid_constant__ CUtensorMap tmap_db, const __grid_constant__ CUtensorMap tmap_q,
                const ScanParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* s_stages = smem;
    unsigned long long* s_queue = (unsigned long long*)(smem + p.off_queue);
    float* s_tau = (float*)(smem + p.off_tau);
    uint64_t* s_bars = (uint64_t*)(smem + p.off_bars);
    uint64_t* bar_full = s_bars;
    uint64_t* bar_empty = s_bars + kMaxStages;
    uint64_t* bar_tfull = s_bars + 2 * kMaxStages;
    uint64_t* bar_tempty = s_bars + 2 * kMaxStages + kMaxSlots;
    uint32_t* s_tmem = (uint32_t*)(s_bars + 2 * kMaxStages + 2 * kMaxSlots + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int num_k = p.d_pad / kBlockK;
    const int nq_blk = p.nq_blk;
    const int nq_half = nq_blk >> 1;
    const uint32_t q_half_bytes = (uint32_t)nq_half * 128u;
    const long long pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
    const long long total_work = p.num_super * (long long)p.num_qblk;
    constexpr int kPairRows = 2 * kBlockM;

    chain_stamp(p.trace, MODE == kModeDense ? 1 : 3, false);
    if (threadIdx.x == 0) {
        if (smem_u32(smem) & 1023u) __trap();
        for (int i = 0; i < kMaxStages; ++i) {
            mbar_init(&bar_full[i], 2);
            mbar_init(&bar_empty[i], 1);
        }
        for (int i = 0; i < kMaxSlots; ++i) {
            mbar_init(&bar_tfull[i], 1);
            mbar_init(&bar_tempty[i], 2 * kEpiWarps);
        }
        fence_mbar_init();
        fence_proxy_async();
    }
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_db);
        prefetch_tmap(&tmap_q);
    }
    if (warp == 1) {
        tmem_alloc_2sm(s_tmem, 512);
        tmem_relinquish_2sm();
    }
    tc_fence_before();
    cluster_sync_all();   // barriers of both CTAs initialised before any remote arrive / multicast commit
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;
    grid_dependency_wait();      // thresholds / queries / counters come from the previous kernels of the chain
    grid_launch_dependents();

    if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        if (elect_one()) {
            const uint64_t hint_db = (p.num_qblk > 1) ? kEvictNormal : kEvictFirst;
            const uint32_t pair_bytes = 2u * (kSubTileBytes + q_half_bytes);
            uint32_t stage = 0, phase = 0;
            for (long long w = pair; w < total_work; w += num_pairs) {
                const long long st = w / p.num_qblk;
                const int qb = (int)(w - st * p.num_qblk);
                const long long blk = st * p.super_stride * 2 + rank;   // this CTA's 128-row block of the tiled DB
                for (int kc = 0; kc < num_k; ++kc) {
                    mbar_wait(&bar_empty[stage], phase ^ 1);
                    if (leader) mbar_expect_tx(&bar_full[stage], pair_bytes);
                    else mbar_arrive_leader(&bar_full[stage]);
                    uint8_t* sA = s_stages + (size_t)stage * p.stage_bytes;
                    tma_load_2d_2sm(&tmap_db, &bar_full[stage], sA, 0, (int)((blk * num_k + kc) * kBlockM), hint_db);
                    tma_load_2d_2sm(&tmap_q, &bar_full[stage], sA + kSubTileBytes, kc * kBlockK,
                                    qb * nq_blk + (int)rank * nq_half, kEvictLast);
                    if (++stage == (uint32_t)p.num_stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (leader && elect_one()) {
            const uint32_t idesc = make_idesc_bf16(kPairRows, (uint32_t)nq_blk);
            const uint32_t ns = (uint32_t)p.num_slots;
            uint32_t stage = 0, phase = 0, acc = 0;
            for (long long w = pair; w < total_work; w += num_pairs, ++acc) {
                const uint32_t slot = acc % ns;
                mbar_wait(&bar_tempty[slot], ((acc / ns) & 1u) ^ 1u);   // both CTAs drained this slot
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + slot * (uint32_t)p.slot_w;
                for (int kc = 0; kc < num_k; ++kc) {
                    mbar_wait(&bar_full[stage], phase);
                    tc_fence_after();
                    const uint32_t sA = smem_u32(s_stages + (size_t)stage * p.stage_bytes);
                    const uint32_t sB = sA + kSubTileBytes;
#pragma unroll
                    for (int k4 = 0; k4 < kBlockK / 16; ++k4)
                        umma_bf16_2sm(d_tmem, make_kmajor_sw128_desc(sA + k4 * 32), make_kmajor_sw128_desc(sB + k4 * 32),
                                      idesc, (uint32_t)((kc | k4) != 0));
                    umma_commit_2sm(&bar_empty[stage]);   // frees this stage in BOTH CTAs
                    if (++stage == (uint32_t)p.num_stages) { stage = 0; phase ^= 1; }
                }
                umma_commit_2sm(&bar_tfull[slot]);
            }
        }
    } else {
        // ===================== epilogue (warps 2..9, both CTAs) =====================
        const int ew = warp - 2;
        const uint32_t lane_base = (uint32_t)(warp & 3) * 32u;
        const int half = ew >> 2;
        const int et = (int)lane_base + lane;
        const int qt = half * kBlockM + et;
        const uint32_t ns = (uint32_t)p.num_slots;
        const int nchunks = nq_blk / 16;
        uint32_t acc = 0;
        int cur_qb = -1;
        int pend_n = 0, pend_q0 = 0;
        uint32_t pend_row = 0;
        int pend_pos[4] = {0, 0, 0, 0};
        unsigned long long pend_ent[4] = {0ull, 0ull, 0ull, 0ull};

        for (long long w = pair; w < total_work; w += num_pairs, ++acc) {
            const long long st = w / p.num_qblk;
            const int qb = (int)(w - st * p.num_qblk);
            const long long row0 = (st * p.super_stride * 2 + rank) * kBlockM;
            const int q0 = qb * nq_blk;
            const int seg = (int)(st % kCandSplit);
            const uint32_t tbuf = p.num_qblk > 1 ? (acc & 1u) : 0u;
            const float* tau_s = s_tau + tbuf * 256;
            const float* tauh_s = s_tau + 512 + tbuf * 256;
            const int hot_seg = kCandSplit + (int)(st % kHotSplit);   // hot sub-list of this super-tile (search path only)

            if (MODE == kModeFilter && qb != cur_qb) {
                const int i = ew * 32 + lane;
                if (i < nq_blk) {
                    s_tau[tbuf * 256 + i] = p.tau[q0 + i];
                    s_tau[512 + tbuf * 256 + i] = p.tau_hot ? p.tau_hot[q0 + i] : __int_as_float(0x7f800000);
                }
                epi_bar_sync2();
                cur_qb = p.num_qblk > 1 ? -1 : qb;
            }

            const uint32_t slot = acc % ns;
            mbar_wait(&bar_tfull[slot], (acc / ns) & 1u);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (lane_base << 16) + slot * (uint32_t)p.slot_w;
            const long long row = row0 + et;
            const bool valid = row < p.n_rows;

            if (MODE == kModeDense) {
                const long long col = st * kPairRows + (long long)rank * kBlockM + et;  // sample index
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
                    float* o = p.dense + (size_t)(q0 + c * 16) * (size_t)p.dense_ld + (size_t)col;
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        o[(size_t)i * (size_t)p.dense_ld] = valid ? __uint_as_float(v[i]) : -__int_as_float(0x7f800000);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_leader(&bar_tempty[slot]);
            } else {
                int n = 0;
                const uint32_t key_row_bits = 0xFFFFFFFFu - (uint32_t)row;
                // sub-list of a survivor, decided when its queue entry is read back (see scan_tc.cu)
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
                                if (n > kQueueCap - 8) {
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
                    __syncwarp();
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_leader(&bar_tempty[slot]);
                // software-pipelined append (see scan_tc.cu)
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
    cluster_sync_all();   // the peer may still be reading this CTA's smem / arriving on its barriers
    if (warp == 1) tmem_dealloc_2sm(tmem_base, 512);
    chain_stamp(p.trace, MODE == kModeDense ? 1 : 3, true);
}

// ------------------------------------------------------------------------------------------------
int make_scan_tmap(CUtensorMap* m, const void* base, long long rows, int d_pad, long long pitch_elems, int box_rows);

int plan_scan_tc2(const TcPlan& base, int d_pad, TcPlan* pl) {
    (void)d_pad;
    *pl = base;
    pl->resident = 0;
    pl->m_sub = 2;                                   // a pair covers 2 x 128 rows
    pl->stage_bytes = (size_t)kSubTileBytes + (size_t)base.nq_blk * 64;
    const size_t fixed = (size_t)kQueueCap * kEpiThreads * 8 + kTauSmemBytes + 512;
    size_t stages = ((size_t)kSmemLimit - fixed) / pl->stage_bytes;
    if (stages > kMaxStages) stages = kMaxStages;
    pl->num_stages = (int)stages;
    pl->off_stages = 0;
    pl->off_queue = stages * pl->stage_bytes;
    pl->off_tau = pl->off_queue + (size_t)kQueueCap * kEpiThreads * 8;
    pl->off_bars = pl->off_tau + kTauSmemBytes;
    pl->smem_bytes = pl->off_bars + 512;
    return RVO_OK;
}

int launch_scan_tc2(int mode, const uint16_t* db, long long n_rows, long long super_stride, int d_pad,
                    const uint16_t* q_bf16, const TcPlan& pl, const float* tau, unsigned long long* cand, int* cand_cnt,
                    int cap, float* dense, long long dense_ld, int sm_count, cudaStream_t stream, const float* tau_hot,
                    int nseg) {
    if (n_rows <= 0) return RVO_OK;
    if (super_stride < 1) super_stride = 1;
    const long long row_blocks = (n_rows + kBlockM - 1) / kBlockM;
    const long long vrows = row_blocks * (d_pad / kBlockK) * kBlockM;
    if (vrows >= (1ll << 31)) {
        set_error("launch_scan_tc2: shard too large (%lld rows x %d); shard the DB", n_rows, d_pad);
        return RVO_E_INVALID;
    }
    CUtensorMap tm_db, tm_q;
    int rc = make_scan_tmap(&tm_db, db, vrows, kBlockK, kBlockK, kBlockM);
    if (rc) return rc;
    rc = make_scan_tmap(&tm_q, q_bf16, pl.nq_pad, d_pad, d_pad, pl.nq_blk / 2);
    if (rc) return rc;

    ScanParams p;
    memset(&p, 0, sizeof(p));
    const long long T = 2 * kBlockM;
    p.n_rows = n_rows;
    p.super_stride = super_stride;
    p.d_pad = d_pad;
    p.nq_blk = pl.nq_blk;
    p.num_qblk = pl.num_qblk;
    p.m_sub = 2;
    p.num_stages = pl.num_stages;
    p.resident_q = 0;
    p.slot_w = pl.slot_w;
    p.num_slots = pl.num_slots;
    p.stage_bytes = (uint32_t)pl.stage_bytes;
    p.off_stages = 0;
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
    long long pairs = sm_count / 2;
    if (pairs > total) pairs = total;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(2 * pairs));
    cfg.blockDim = dim3(kScanThreads);
    cfg.dynamicSmemBytes = pl.smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // see ptx.cuh grid_dependency_wait
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g_use_pdl.load() ? 2 : 1;
    if (mode == kModeDense || mode == kModeDenseMax) {
        RVO_CUDA(cudaFuncSetAttribute(scan_tc2_kernel<kModeDense>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)pl.smem_bytes));
        RVO_CUDA(cudaLaunchKernelEx(&cfg, scan_tc2_kernel<kModeDense>, tm_db, tm_q, p));
    } else {
        RVO_CUDA(cudaFuncSetAttribute(scan_tc2_kernel<kModeFilter>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)pl.smem_bytes));
        RVO_CUDA(cudaLaunchKernelEx(&cfg, scan_tc2_kernel<kModeFilter>, tm_db, tm_q, p));
    }
    RVO_LAUNCHED();
    return RVO_OK;
}

}  // namespace rvo
