// scan_tc.cu -- tcgen05 / TMA streaming scorer: the bandwidth-bound scan for query batches.
//
// One persistent CTA per SM, warp-specialised (192 threads; 288 with the threshold warp, see below):
//   warp 0   TMA producer   streams the store as 128-row x 128-byte boxes (SWIZZLE_128B) through a
//                           ring of 16 KB shared-memory stages, L2 evict-first; loads the
//                           normalised queries once (evict-last) -- no thread ever touches a row
//   warp 1   MMA issuer     one thread issues tcgen05.mma (kind::tf32 for an fp32 store, kind::f16
//                           for bf16): D[128 rows x nq] += A[128 x 32B] * Q[nq x 32B]^T, accumulators
//                           in TMEM (8 stages x nq columns), tcgen05.commit frees stages / publishes D
//   warps 2-5 epilogue      tcgen05.ld their 32-lane TMEM quadrant (lane = store row, column = query),
//                           multiply by the cached 1/||row|| (fused normalisation), compare with the
//                           per-query running threshold, push the rare survivors into per-query
//                           queues; thread q then folds them into query q's unsorted top-kp set in
//                           shared memory (replace the minimum, rescan with pipelined loads).  While a
//                           tile overflows a queue its accumulator simply stays in TMEM and is re-read
//                           after the drain.
// Threshold seeding: each CTA starts without a threshold, so for its first tiles every score would
// be a "survivor".  Instead every CTA first publishes, per query, the maximum score of its first
// tile; the kp-th largest of those per-CTA maxima is a valid lower bound on the shard's kp-th best
// score (kp distinct rows reach it), and with ~148 CTAs it is as tight as the kp-th best of the
// first 19 K rows.  All CTAs adopt it as their initial threshold (bounded wait, no grid barrier
// semantics needed: a late CTA only makes the bound looser), while TMA/MMA keep streaming into the
// 8 TMEM stages.  This removes the per-CTA warm-up flood that otherwise dominates small shards.
// Threshold warp (template TW, narrow tiles): a CTA's own kp-th best is a weak filter -- it has seen
// 1/148 of the shard -- so ~every tile still produced a survivor and paid the two-barrier drain, and the
// epilogue, not HBM, paced bf16 scans.  With TW the drains keep each CTA's RUNNING maximum per query
// current in the seed table and one extra warp per CTA runs a distributed bound service: the CTA that
// owns query q (q mod G) keeps re-deriving the exact kp-th largest of all CTAs' maxima (same proof as the
// first seed: kp distinct rows reach it) and publishes it with an atomic max; every CTA adopts the
// published bounds of all queries with a CAS-max on its shared-memory thresholds.  The bound then follows
// the shard-wide top-kp (~40th best of everything scanned so far), survivors become rare, and the epilogue
// drops off the critical path.  The first seed is derived the same way (one select per owner CTA, a second
// short bounded wait) instead of 64 selects in every CTA.
// HBM traffic = the store bytes exactly once per batch of <= 64 queries; the per-CTA lists are
// merged and exactly rescored by select.cu.
// Collect mode (second pass for queries the first pass could not certify): thresholds are fixed per
// query (exact k-th candidate score - 2 eps; +inf for queries that need nothing) and every row at or
// above its query's threshold is appended to that query's global buffer -- no lists, no drains.
// Dump mode (template DUMP, stores of <= 9472 rows): with one or two tiles per CTA there is no
// threshold to filter with -- every row of a first tile is a "survivor" and the serial per-query
// drain dominates (measured 82 us for 5000 rows x 30 queries).  Instead the epilogue writes the
// key of EVERY row, [tile][query][128], and select.cu ranks all of them per query; nothing is
// dropped before the exact rescoring, so the certification bound is the kp-th best approximate score.
#include "tc_common.cuh"
#include <stdlib.h>

namespace vm {
using namespace tc;

static constexpr int TC_BLOCK_M = 128;
static constexpr int TC_STAGE_BYTES = TC_BLOCK_M * 128;
static constexpr int TC_ACC = 8;      // accumulator stages in TMEM (8 x 64 columns = all 512)
static constexpr int TC_QCAP = 32;
static constexpr int TC_THREADS_TW = 288;  // with the threshold warp (TW)
static constexpr int TC_THREADS = 192;     // without: TMA warp, MMA warp, 4 epilogue warps
// 288:  // TMA warp, MMA warp, 4 epilogue warps, (2 idle), threshold warp 8 -> scheduler 0,
                                         // away from the schedulers of the warps that run the drains (2 and 3)
static constexpr int TC_TMEM_COLS = 512;
static constexpr int TC_MAX_STAGES = 8;
static constexpr int SEED_TAB_WORDS = 256 * 64;  // per-CTA maxima [<= 256 CTAs][<= 64 queries]; the published bounds follow

// second-pass ("collect") arguments; thr == nullptr selects the normal top-kp mode
struct CollectArgs {
    const float *thr;      // [nq] per-query score threshold (+inf: ignore the query)
    uint64_t *buf;         // [nq][cap] collected keys
    int *cnt;              // [nq] number of rows at or above the threshold (may exceed cap)
    int cap;
    const int *pending;    // device counter of uncertified queries: 0 -> the kernel exits at once
};

struct TcLayout {
    int nq_pad, KB, stages, kp;
    uint32_t off_b, off_a, off_list, off_queue, off_tauk, off_tauf, off_seedf, off_qcnt, off_flags, off_bars, off_tmem, total;
};

static TcLayout make_layout(int dtype, int ld, int nq, int kp)
{
    TcLayout L{};
    const int es = dtype == VM_F32 ? 4 : 2;
    L.nq_pad = (nq + 15) & ~15;
    L.KB = (ld * es + 127) / 128;
    L.kp = kp;
    uint32_t o = 0;
    L.off_b = o; o += (uint32_t)L.KB * L.nq_pad * 128;
    uint32_t epi = (uint32_t)kp * L.nq_pad * 8 + (uint32_t)TC_QCAP * L.nq_pad * 8 + (uint32_t)L.nq_pad * (8 + 4 + 4 + 4) + 64 + 512;
    int64_t room = 227 * 1024 - 1024 - (int64_t)o - epi;
    L.stages = (int)(room / TC_STAGE_BYTES);
    if (L.stages > TC_MAX_STAGES) L.stages = TC_MAX_STAGES;
    if (L.stages < 0) L.stages = 0;
    L.off_a = o; o += (uint32_t)L.stages * TC_STAGE_BYTES;
    L.off_list = o; o += (uint32_t)kp * L.nq_pad * 8;
    L.off_queue = o; o += (uint32_t)TC_QCAP * L.nq_pad * 8;
    L.off_tauk = o; o += (uint32_t)L.nq_pad * 8;
    L.off_tauf = o; o += (uint32_t)L.nq_pad * 4;
    L.off_seedf = o; o += (uint32_t)L.nq_pad * 4;
    L.off_qcnt = o; o += (uint32_t)L.nq_pad * 4;
    L.off_flags = o; o += 64;
    L.off_bars = o; o += 8 * (2 * TC_MAX_STAGES + 1 + 2 * TC_ACC);
    L.off_tmem = o; o += 16;
    L.total = o + 1024;  // slack for the 1024-byte alignment of the swizzled tiles
    return L;
}


// monotone update of a shared-memory float (several writers, values only grow)
__device__ __forceinline__ void smem_fmax(float *addr, float v)
{
    uint32_t *a = reinterpret_cast<uint32_t *>(addr);
    uint32_t old = *reinterpret_cast<volatile uint32_t *>(a);
    while (!(__uint_as_float(old) >= v)) {
        const uint32_t prev = atomicCAS(a, old, __float_as_uint(v));
        if (prev == old) break;
        old = prev;
    }
}

// Cheap lower bound on the kp-th largest of the per-CTA maxima of query q (one warp, kp <= 64 <= G <= 256... or
// kp <= 32 <= G): lane l holds the entries of CTAs l, l+32, ...; the minimum over the lanes of each lane's
// largest (kp <= 32) or second largest (kp <= 64) entry is reached by at least kp distinct CTAs.  ~40 cycles
// after the loads instead of ~1000 for the exact select; typically the ~2*kp-th largest instead of the kp-th.
__device__ __forceinline__ uint32_t seed_bound_fast(const uint32_t *seed_tab, int G, int nq_pad, int q, int kp, int lane)
{
    uint32_t m1 = 0, m2 = 0;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        const int i = lane + 32 * t;
        const uint32_t x = i < G ? __ldcg(seed_tab + (size_t)i * nq_pad + q) : 0u;
        m2 = max(m2, min(m1, x));
        m1 = max(m1, x);
    }
    return __reduce_min_sync(0xffffffffu, kp > 32 ? m2 : m1);
}
__device__ __forceinline__ float seed_from_ordered(uint32_t o)
{
    const float sd = f32_from_ordered(o);
    return sd > -INFINITY ? sd : -INFINITY;  // 0 (an empty slot) maps to a NaN pattern
}

// kp-th largest of the per-CTA maxima of query q (one warp, G <= 256 CTAs): a lower bound on the shard's
// kp-th best score, because kp distinct rows (one per CTA) reach it.  16 radix bits; truncation rounds down.
__device__ __forceinline__ float seed_select(const uint32_t *seed_tab, int G, int nq_pad, int q, int kp, int lane)
{
    uint32_t vals[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        const int i = lane + 32 * t;
        vals[t] = i < G ? __ldcg(seed_tab + (size_t)i * nq_pad + q) : 0u;
    }
    uint32_t prefix = 0;
    int need = kp;
    for (int bit = 31; bit >= 16; --bit) {
        const uint32_t hi = bit == 31 ? 0u : (0xFFFFFFFFu << (bit + 1));
        int cnt = 0;
#pragma unroll
        for (int t = 0; t < 8; ++t) cnt += ((vals[t] & hi) == prefix && ((vals[t] >> bit) & 1u)) ? 1 : 0;
        const int tot = __reduce_add_sync(0xffffffffu, cnt);
        if (tot >= need) prefix |= 1u << bit;
        else need -= tot;
    }
    float sd = f32_from_ordered(prefix);
    if (!(sd > -INFINITY) || G < kp) sd = -INFINITY;  // also catches NaN patterns
    return sd;
}

template <bool TF32, bool DUMP, bool TW>
__global__ void __launch_bounds__(TW ? TC_THREADS_TW : TC_THREADS, 1)
scan_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const float *__restrict__ inv_norms, int64_t n, int num_tiles, int nq, TcLayout L, uint64_t *__restrict__ cand,
               uint32_t *__restrict__ seed_tab, int *__restrict__ seed_ctr, CollectArgs col, int tw_sleep)
{
    const bool collect = !DUMP && col.thr != nullptr;
    if (collect && *col.pending == 0) return;  // nothing left to refine (uniform across the grid)

    extern __shared__ __align__(16) uint8_t smem_raw[];
    // align to 1024 B with pointer arithmetic on the __shared__ array (an integer round trip would
    // demote every later access to generic LD/ST)
    uint8_t *base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t *sB = base + L.off_b;
    uint8_t *sA = base + L.off_a;
    uint64_t *list = (uint64_t *)(base + L.off_list);    // [kp][nq_pad] unsorted top-kp set per query
    uint64_t *queue = (uint64_t *)(base + L.off_queue);  // [QCAP][nq_pad]
    uint64_t *tauk = (uint64_t *)(base + L.off_tauk);
    float *tauf = (float *)(base + L.off_tauf);
    float *seedf = (float *)(base + L.off_seedf);         // [nq_pad] seeded lower bound per query (only grows)
    int *qcnt = (int *)(base + L.off_qcnt);
    volatile int *s_hit = (volatile int *)(base + L.off_flags);  // [2]
    volatile int *s_ovf = s_hit + 2;                             // [2]
    volatile int *s_done = s_hit + 4;                            // epilogue finished (stops the threshold warp)
    uint64_t *bars = (uint64_t *)(base + L.off_bars);
    uint64_t *full = bars, *empty = bars + TC_MAX_STAGES, *qfull = bars + 2 * TC_MAX_STAGES;
    uint64_t *tfull = qfull + 1, *tempty = tfull + TC_ACC;
    uint32_t *tmem_slot = (uint32_t *)(base + L.off_tmem);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    const int nq_pad = L.nq_pad, KB = L.KB, stages = L.stages, kp = L.kp;
    constexpr int ELEMS = TF32 ? 32 : 64;  // elements per 128-byte K block

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(qfull, 1);
        for (int a = 0; a < TC_ACC; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 4); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, TC_TMEM_COLS);
    if (warp >= 2) {
        const int e = threadIdx.x - 64;
        if (e < nq_pad) {
            for (int j = 0; j < kp; ++j) list[j * nq_pad + e] = 0;
            tauk[e] = 0; qcnt[e] = 0; seedf[e] = -INFINITY;
            tauf[e] = collect ? (e < nq ? col.thr[e] : INFINITY) : -INFINITY;
        }
        if (e < 5) s_hit[e] = 0;  // s_hit[0..1], s_ovf[0..1], s_done
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================================ TMA producer ================================
        if (lane == 0) {
            mbar_arrive_expect_tx(qfull, (uint32_t)KB * nq_pad * 128);
            for (int kb = 0; kb < KB; ++kb) tma_load_2d(&tmB, qfull, sB + (size_t)kb * nq_pad * 128, kb * ELEMS, 0, L2_EVICT_LAST);
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                for (int kb = 0; kb < KB; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    mbar_arrive_expect_tx(&full[stage], TC_STAGE_BYTES);
                    tma_load_2d(&tmA, &full[stage], sA + (size_t)stage * TC_STAGE_BYTES, kb * ELEMS, tile * TC_BLOCK_M, L2_EVICT_FIRST);
                    if (++stage == stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ==================================
        if (lane == 0) {
            const uint32_t idesc = make_idesc(TF32 ? 2u : 1u, TC_BLOCK_M, (uint32_t)nq_pad);
            mbar_wait(qfull, 0);
            tc_fence_after();
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0;
            const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                mbar_wait(&tempty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * nq_pad);
                for (int kb = 0; kb < KB; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t a_addr = a0 + (uint32_t)stage * TC_STAGE_BYTES;
                    const uint32_t b_addr = b0 + (uint32_t)kb * nq_pad * 128;
#pragma unroll
                    for (int j = 0; j < 4; ++j)  // 4 x 32-byte K slices per 128-byte swizzle row
                        umma<TF32>(d_tmem, make_smem_desc_sw128(a_addr + j * 32), make_smem_desc_sw128(b_addr + j * 32), idesc,
                                   (uint32_t)((kb | j) != 0));
                    umma_commit(&empty[stage]);  // stage reusable once these MMAs have read it
                    if (++stage == stages) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tfull[acc]);  // accumulator complete
                if (++acc == TC_ACC) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp < 6) {
        // ================================ epilogue =====================================
        const int e = threadIdx.x - 64;   // 0..127
        const int quad = warp & 3;        // TMEM lane quadrant this warp may read
        const int row_in_tile = quad * 32 + lane;
        uint64_t my_tau = 0;              // threshold key (= minimum of the set) of query e (threads e < nq_pad)
        int my_min = 0;                   // its position in the set
        uint32_t my_max = 0;              // best score (ordered bits) in query e's set == this CTA's entry in seed_tab
        int acc = 0, par = 0;
        uint32_t acc_phase = 0;
        if constexpr (DUMP) {
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int64_t r = (int64_t)tile * TC_BLOCK_M + row_in_tile;
                const float iv = r < n ? __ldg(inv_norms + r) : -1.0f;
                mbar_wait(&tfull[acc], acc_phase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * nq_pad);
                uint64_t *out = cand + (int64_t)tile * nq * TC_BLOCK_M + row_in_tile;
                for (int c = 0; c < nq_pad; c += 16) {
                    uint32_t v[16];
                    tmem_ld16(taddr + c, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (c + j < nq)  // coalesced: 32 consecutive rows of one query per warp store
                            out[(int64_t)(c + j) * TC_BLOCK_M] = iv >= 0.0f ? make_key(__uint_as_float(v[j]) * iv, (uint32_t)r) : 0ull;
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[acc]);
                if (++acc == TC_ACC) { acc = 0; acc_phase ^= 1; }
            }
        } else {
        if (seed_tab != nullptr && !collect) {
            // ---- cooperative threshold seeding from the first tile (see file header) ----
            uint32_t *wmax = reinterpret_cast<uint32_t *>(queue);  // [4][nq_pad] scratch (queue is idle now)
            const int64_t row0 = (int64_t)blockIdx.x * TC_BLOCK_M + row_in_tile;
            const float inv0 = row0 < n ? __ldg(inv_norms + row0) : -1.0f;
            mbar_wait(&tfull[0], 0);
            tc_fence_after();
            const uint32_t taddr0 = tmem_base + ((uint32_t)(quad * 32) << 16);
            for (int c = 0; c < nq_pad; c += 16) {
                uint32_t v[16];
                tmem_ld16(taddr0 + c, v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float sc = inv0 >= 0.0f ? __uint_as_float(v[j]) * inv0 : -INFINITY;
                    const uint32_t mx = __reduce_max_sync(0xffffffffu, f32_ordered(sc));
                    if (lane == 0) wmax[(warp - 2) * nq_pad + c + j] = mx;
                }
            }
            named_bar_sync(1, 128);
            if (e < nq_pad) {
                uint32_t m = wmax[e];
#pragma unroll
                for (int w = 1; w < 4; ++w) m = max(m, wmax[w * nq_pad + e]);
                seed_tab[(size_t)blockIdx.x * nq_pad + e] = m;
                my_max = m;
            }
            __threadfence();
            named_bar_sync(1, 128);
            if (e == 0) {
                atomicAdd(seed_ctr, 1);
                const long long t0 = clock64();
                while (*(volatile int *)seed_ctr < (int)gridDim.x)
                    if (clock64() - t0 > 400000) break;  // ~0.2 ms: late CTAs only loosen the bound
                __threadfence();
            }
            named_bar_sync(1, 128);
            // The bound of query q is derived once, by the CTA that owns it (q mod G), and published; a second,
            // short bounded wait lets every CTA adopt all nq bounds (64 exact selects in every CTA cost ~15 us).
            const int G = (int)gridDim.x;
            uint32_t *gbound = seed_tab + SEED_TAB_WORDS;
            int *pub_ctr = reinterpret_cast<int *>(gbound + 64);
            if (warp == 2) {
                for (int q = blockIdx.x; q < nq; q += G) {
                    const float sd = seed_select(seed_tab, G, nq_pad, q, kp, lane);
                    if (lane == 0) {
                        if (sd > -INFINITY) atomicMax(gbound + q, f32_ordered(sd));
                        __threadfence();
                        atomicAdd(pub_ctr, 1);
                    }
                }
                if (lane == 0) {
                    const long long t0 = clock64();
                    while (*(volatile int *)pub_ctr < nq)
                        if (clock64() - t0 > 200000) break;  // ~0.1 ms: a missing bound only means "no threshold yet"
                    __threadfence();
                }
            }
            named_bar_sync(1, 128);
            if (e < nq) {
                const uint32_t o = __ldcg(gbound + e);
                if (o != 0u) {
                    const float f = f32_from_ordered(o);
                    smem_fmax(&seedf[e], f);
                    smem_fmax(&tauf[e], f);
                }
            }
            named_bar_sync(1, 128);
        }
        int64_t row = (int64_t)blockIdx.x * TC_BLOCK_M + row_in_tile;
        float inv = row < n ? __ldg(inv_norms + row) : -1.0f;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            // prefetch the next tile's inverse norm so its DRAM latency hides behind this tile
            const int64_t row_next = row + (int64_t)gridDim.x * TC_BLOCK_M;
            const float inv_next = (tile + (int)gridDim.x < num_tiles && row_next < n) ? __ldg(inv_norms + row_next) : -1.0f;
            const bool valid = inv >= 0.0f;  // < 0: beyond the shard or a skipped row
            mbar_wait(&tfull[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * nq_pad);
            uint64_t pushed = 0;
            for (;;) {
                bool hit_any = false, ovf = false;
                // 16 queries (TMEM columns) at a time; the load of chunk c+1 is issued before chunk c is
                // processed so its TMEM latency hides behind the compares
                auto process = [&](const uint32_t (&v)[16], int c) {
                    // thresholds of these 16 queries, loaded up front (a stale, lower value only lets an
                    // extra candidate through to the drain, which re-checks against the live key)
                    float tf[16];
#pragma unroll
                    for (int j4 = 0; j4 < 4; ++j4) {
                        const float4 t4 = *reinterpret_cast<const float4 *>(tauf + c + 4 * j4);
                        tf[4 * j4] = t4.x; tf[4 * j4 + 1] = t4.y; tf[4 * j4 + 2] = t4.z; tf[4 * j4 + 3] = t4.w;
                    }
                    if (valid) {
                        uint32_t hits = 0;
#pragma unroll
                        for (int j = 0; j < 16; ++j) hits |= (__uint_as_float(v[j]) * inv >= tf[j]) ? (1u << j) : 0u;
                        hits &= ~(uint32_t)(pushed >> c) & (nq - c >= 16 ? 0xFFFFu : ((1u << (nq - c > 0 ? nq - c : 0)) - 1u));
                        while (hits) {  // rare path
                            const int j = __ffs(hits) - 1;
                            hits &= hits - 1;
                            const int q = c + j;
                            float s = 0.0f;
#pragma unroll
                            for (int jj = 0; jj < 16; ++jj) s = (jj == j) ? __uint_as_float(v[jj]) * inv : s;
                            const uint64_t key = make_key(s, (uint32_t)row);
                            if (collect) {
                                const int pos = atomicAdd(col.cnt + q, 1);
                                if (pos < col.cap) col.buf[(size_t)q * col.cap + pos] = key;
                            } else if (key > tauk[q]) {
                                hit_any = true;
                                const int pos = atomicAdd(&qcnt[q], 1);
                                if (pos < TC_QCAP) { queue[pos * nq_pad + q] = key; pushed |= 1ull << q; }
                                else ovf = true;
                            }
                        }
                    }
                };
                uint32_t va[16], vb[16];
                tmem_ld16(taddr, va);
                for (int c = 0; c < nq_pad; c += 32) {
                    tmem_ld_wait();
                    if (c + 16 < nq_pad) tmem_ld16(taddr + c + 16, vb);
                    process(va, c);
                    if (c + 16 < nq_pad) {
                        tmem_ld_wait();
                        if (c + 32 < nq_pad) tmem_ld16(taddr + c + 32, va);
                        process(vb, c + 16);
                    }
                }
                if (hit_any) s_hit[par] = 1;
                if (ovf) s_ovf[par] = 1;
                if (e == 0) { s_hit[par ^ 1] = 0; s_ovf[par ^ 1] = 0; }  // arm the next round's flags
                named_bar_sync(1, 128);
                const int h = s_hit[par], o = s_ovf[par];
                par ^= 1;
                if (!h) break;
                if (e < nq_pad) {
                    // drain query e's queue: replace the current minimum of the (unsorted) top-kp set,
                    // then rescan for the new minimum with independent, pipelined loads
                    int cnt = qcnt[e];
                    cnt = cnt < TC_QCAP ? cnt : TC_QCAP;
                    const uint32_t max_before = my_max;
                    for (int i = 0; i < cnt; ++i) {
                        const uint64_t key = queue[i * nq_pad + e];
                        if (key > my_tau) {
                            if (TW) my_max = max(my_max, (uint32_t)(key >> 32));
                            list[my_min * nq_pad + e] = key;
                            uint64_t m = ~0ull;
                            int p = 0;
#pragma unroll 8
                            for (int j = 0; j < kp; ++j) {
                                const uint64_t x = list[j * nq_pad + e];
                                if (x < m) { m = x; p = j; }
                            }
                            my_tau = m;
                            my_min = p;
                        }
                    }
                    qcnt[e] = 0;
                    tauk[e] = my_tau;
                    const float sd = seedf[e];
                    if (TW) {
                        smem_fmax(&tauf[e], my_tau ? fmaxf(key_score(my_tau), sd) : sd);
                        if (seed_tab != nullptr && my_max != max_before) __stcg(seed_tab + (size_t)blockIdx.x * nq_pad + e, my_max);
                    } else {
                        tauf[e] = my_tau ? fmaxf(key_score(my_tau), sd) : sd;
                    }
                }
                named_bar_sync(1, 128);
                if (!o) break;
            }
            // release the accumulator stage to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
            if (++acc == TC_ACC) { acc = 0; acc_phase ^= 1; }
            row = row_next;
            inv = inv_next;
        }
        named_bar_sync(1, 128);
        for (int i = e; i < nq * kp && !collect; i += 128) {
            const int q = i / kp, j = i - q * kp;
            cand[((int64_t)blockIdx.x * nq + q) * kp + j] = list[j * nq_pad + q];
        }
        }  // !DUMP
        if (e == 0) *s_done = 1;
    } else {
        // ================================ threshold warp ================================
        // Keeps re-deriving, query by query, the kp-th largest of all CTAs' RUNNING maxima (seed_tab, kept
        // current by the drains) and raises the query's threshold to it: the bound then follows the
        // shard-wide top-kp instead of this CTA's own, so survivors -- and the per-tile drains they cause --
        // become rare.  Off the epilogue's critical path; a stale table entry only loosens the bound.
        if (TW && !DUMP && warp == 8 && seed_tab != nullptr && !collect) {
            // A distributed bound service: CTA b owns queries b, b+G, ... (at most one with G >= nq), keeps
            // re-deriving their exact kp-th largest running maximum and publishes it with an atomic max; every
            // CTA adopts the published bounds of all queries.  One table column per owner per round instead of
            // every column in every CTA keeps the shared 38 KB table from becoming an L2 hot spot.
            const int G = (int)gridDim.x;
            uint32_t *gbound = seed_tab + SEED_TAB_WORDS;
            // the exit test is a warp vote: a lane that lags behind must not leave the loop while the others
            // are already inside the next round's warp reductions
            while (!__any_sync(0xffffffffu, *s_done != 0)) {
                for (int q = blockIdx.x; q < nq; q += G) {
                    const float sd = seed_select(seed_tab, G, nq_pad, q, kp, lane);
                    if (lane == 0 && sd > -INFINITY) atomicMax(gbound + q, f32_ordered(sd));
                }
                for (int q = lane; q < nq; q += 32) {
                    const uint32_t o = __ldcg(gbound + q);
                    if (o != 0u) {
                        const float f = f32_from_ordered(o);
                        if (f > seedf[q]) { smem_fmax(&seedf[q], f); smem_fmax(&tauf[q], f); }
                    }
                }
                __nanosleep(tw_sleep);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TC_TMEM_COLS);
    }
}

bool scan_tc_supported(int dtype, int dim, int nq, int kp)
{
    if (dtype != VM_F32 && dtype != VM_BF16) return false;
    if (nq < 1 || nq > 64 || kp < 1 || kp > 64) return false;
    TcLayout L = make_layout(dtype, ld_for_dim(dim), nq, kp);
    return L.stages >= 3;
}

// queries_store_dtype: the normalised queries [nq_pad][ld] in the STORE dtype (fp32 for the tf32
// path, bf16 for the bf16 path).
int launch_scan_tc(const ScanArgs &a, const void *queries_store_dtype, uint32_t *seed_tab, int *seed_ctr, const ScanCollect *sc,
                   ScanInfo *info)
{
    VM_REQUIRE(a.n >= 1 && a.n < 0x7FFFFF00LL, VM_ERR_UNSUPPORTED, "tcgen05 scan: shard rows %lld outside [1, 2^31)", (long long)a.n);
    TcLayout L = make_layout(a.dtype, a.ld, a.nq, a.kp);
    VM_REQUIRE(L.stages >= 3, VM_ERR_UNSUPPORTED, "tcgen05 scan: dim %d x %d queries does not fit shared memory", a.dim, a.nq);
    CUtensorMap tmA, tmB;
    int rc = make_tmap_2d_cached(&tmA, a.rows, a.dtype, (uint64_t)a.n, (uint64_t)a.ld, (uint64_t)a.ld, TC_BLOCK_M);
    if (rc != VM_OK) return rc;
    rc = make_tmap_2d_cached(&tmB, queries_store_dtype, a.dtype, (uint64_t)L.nq_pad, (uint64_t)a.ld, (uint64_t)a.ld, (uint32_t)L.nq_pad);
    if (rc != VM_OK) return rc;
    const int num_tiles = (int)((a.n + TC_BLOCK_M - 1) / TC_BLOCK_M);
    int dev_idx = 0;
    VM_CUDA_CHECK(cudaGetDevice(&dev_idx));
    dev_idx &= 63;
    // seeding needs at least kp CTAs (kp first-tile maxima); it pays off even with a single tile per CTA,
    // because a first tile without a threshold costs ~80 us of serial drains (stores below that use dump mode)
    const bool seed = seed_tab && seed_ctr && a.ctas >= a.kp && a.ctas <= 256;
    if (!seed || sc) { seed_tab = nullptr; seed_ctr = nullptr; }
    CollectArgs col{};
    if (sc) { col.thr = sc->thr; col.buf = sc->buf; col.cnt = sc->cnt; col.cap = sc->cap; col.pending = sc->pending; }
    // Threshold warp: pays off where the epilogue, not HBM, paces the scan -- tiles of <= 128 KB (bf16 rows
    // of 384 dims stream in 2.2 us; an fp32 tile takes 4.4 us and hides the drains, and measures 3 % slower
    // with the extra warp) -- once a CTA streams enough tiles for the shared bound to matter.  Measured on
    // 64-query batches, bf16: 1M rows 0.247 -> 0.172 ms, 4M rows 0.68 -> 0.49 ms, 12.5M rows 1.68 -> 1.49 ms.
    constexpr int tw_min_tiles = 16, tw_max_tile_kb = 128;
    const int tw_sleep = 200;
    // (a single-query scan has no epilogue pressure: C5 bf16 measured 1.11 ms without vs 1.16 ms with it)
    const bool tw = seed_tab != nullptr && !sc && !a.dump && a.nq > 16 && (int64_t)num_tiles >= (int64_t)tw_min_tiles * a.ctas &&
                    (int64_t)TC_BLOCK_M * a.ld * (a.dtype == VM_F32 ? 4 : 2) <= (int64_t)tw_max_tile_kb * 1024;
    const bool dump = a.dump && !sc;
    if (dump) VM_REQUIRE(TC_BLOCK_M == SCAN_DUMP_TILE && (int64_t)num_tiles * TC_BLOCK_M <= SCAN_DUMP_MAX_KEYS && a.ctas == num_tiles,
                         VM_ERR_UNSUPPORTED, "tcgen05 scan: dump mode needs one CTA per tile and at most %d rows", SCAN_DUMP_MAX_KEYS);
    if (info && !sc) { info->stages = L.stages; info->variant = dump ? 1 : (tw ? 2 : 0); }
#define LAUNCH_TC(TF, DU, TWV)                                                                                                \
    do {                                                                                                                   \
        static bool set[64] = {}; /* the attribute is per device */                                                        \
        if (!set[dev_idx]) {                                                                                               \
            VM_CUDA_CHECK(cudaFuncSetAttribute(scan_tc_kernel<TF, DU, TWV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); \
            set[dev_idx] = true;                                                                                           \
        }                                                                                                                  \
        scan_tc_kernel<TF, DU, TWV><<<a.ctas, TWV ? TC_THREADS_TW : TC_THREADS, L.total, a.stream>>>(tmA, tmB, a.inv_norms, a.n, num_tiles, a.nq, L, a.cand, \
                                                                          seed_tab, seed_ctr, col, tw_sleep);             \
    } while (0)
    if (a.dtype == VM_F32) { if (dump) LAUNCH_TC(true, true, false); else if (tw) LAUNCH_TC(true, false, true); else LAUNCH_TC(true, false, false); }
    else { if (dump) LAUNCH_TC(false, true, false); else if (tw) LAUNCH_TC(false, false, true); else LAUNCH_TC(false, false, false); }
#undef LAUNCH_TC
    VM_CUDA_CHECK(cudaGetLastError());
    return VM_OK;
}

}  // namespace vm
