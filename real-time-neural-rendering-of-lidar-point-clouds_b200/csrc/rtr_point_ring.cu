// The two point passes as persistent, TMA-fed kernels (sm_100a) — the default for culled frames since round 1h.
//
//   zmin_ring_kernel   <- minDepthPass   (render.cu:53-83)
//   blend_ring_kernel  <- accumulatePass (render.cu:85-130)
//
// What bounded the passes on B200 (ncu, profiles/r01g): not HBM but load latency — with the cloud coming in through
// per-thread LDG.128 every warp serialised "chunk id -> 16-byte records -> z-buffer gather -> atomic" and sat on the
// long scoreboard — and the SM's reduction path (REDs leave an SM at about one 32-byte sector per clock).  Here
//
//  * whole 1024-record chunks (16 KB) are streamed into a 6-stage shared-memory ring with cp.async.bulk (TMA, L2
//    evict-first) signalled through mbarriers: a CTA is two groups of 256 threads, each group consumes every other
//    tile and its thread 0 refills a stage the moment the group has copied it into registers, so the HBM stream runs
//    ahead of the arithmetic and costs no registers or LSU issue slots (2 CTAs = 32 warps and 12 chunks = 192 KB in
//    flight per SM; 3 CTAs x 4 stages and a dedicated producer warp measured the same);
//  * each thread of a group takes FOUR CONSECUTIVE records of the chunk.  The cloud is Morton-ordered, so these are
//    spatial neighbours and mostly project to the same pixel wherever the scan is denser than the pixel grid (5
//    points per pixel on average at C3): they are merged in registers (min of the depth bits / sums of the colour
//    bytes — both exact and order-free) before anything touches memory, which removes REDs rather than speeding them
//    up;
//  * the survivors do the early depth test through L1 and one REDG each.
//
// The same kernels serve the culled frame (tiles = the frame's visible-chunk list; what option ring = 1 uses them
// for) and, with ring = 2, the stream-all frame (every chunk, visited in a low-discrepancy order so that the CTAs in
// flight hold a mix of in-frustum and out-of-frustum chunks; the per-thread LDG.128 kernels stream 4 % faster there
// and stay the default for it).
#include <type_traits>

#include "rtr_kernels.h"

namespace rtr {

constexpr int kRingStages = 6;
constexpr int kRingGroups = 2;                        // consumer groups per CTA, each takes every kRingGroups-th tile
constexpr int kRingConsumers = kPointBlock;           // a group: 256 threads x 4 consecutive records = one chunk
constexpr int kRingThreads = kRingGroups * kRingConsumers;  // 512: no dedicated producer warp, thread 0 of a group refills its stages
constexpr int kRingCtasPerSm = 2;                     // 2 x (96 KB ring + barriers) per SM: 32 warps, up to 64 registers
// What __launch_bounds__ is told.  3 (shared memory admits 2 anyway) would cap the kernels at 40 registers without a
// spill and leave 24 K instead of 16 K registers per SM to the other frame's image kernels: measured 3 % SLOWER with two
// frames in flight (profiles/r01j_exp_ring_dynamic.json, section 7), so the kernels keep their 48.
constexpr int kRingMinCtas = 2;
constexpr int kRingPerThread = kChunkPoints / kRingConsumers;
static_assert(kRingStages % kRingGroups == 0, "every group must own a fixed subset of the stages");
static_assert(kRingPerThread == 4, "the bank-conflict-free XOR swizzle below assumes 4 records per thread");

struct RingSmem {
    PointRecord rec[kRingStages][kChunkPoints];  // 6 x 16 KB
    unsigned long long full[kRingStages];        // producer -> consumers: the chunk's bytes have landed
    unsigned long long empty[kRingStages];       // consumers -> producer: every warp of the group has its records in registers
    uint32_t chunk[kRingStages];                 // chunk id staged in the slot
};

// ---------------------------------------------------------------- mbarrier / bulk-copy PTX
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "RTR_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra RTR_DONE;\n"
        "bra RTR_WAIT;\n"
        "RTR_DONE:\n"
        "}\n" ::"r"(smem_addr(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
// 1-D bulk copy global -> shared, completion counted in bytes on `bar` (SASS: UBLKCP.S.G).
__device__ __forceinline__ void bulk_load(void* dst_smem, const void* src, uint32_t bytes, unsigned long long* bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_addr(dst_smem)),
        "l"(src), "r"(bytes), "r"(smem_addr(bar)), "l"(policy)
        : "memory");
}

// Claim the next tile of a launch.  atom.inc with a bound below 2^32 - 1, not atom.add: ptxas warp-aggregates an add
// (and an inc that can be rewritten as one) to a uniform address — vote + ATOMG by one lane + SHFL of the result — and
// that shuffle waits for the atomic's round trip on the spot, whereas this result is first read one ring iteration
// later.  SASS: ATOMG.E.INC, no SHFL.
__device__ __forceinline__ uint32_t claim_tile(uint32_t* counter) {
    uint32_t v;
    asm volatile("atom.relaxed.gpu.global.inc.u32 %0, [%1], 0x7FFFFFFF;" : "=r"(v) : "l"(counter) : "memory");
    return v;
}

// Which chunk tile number t of this launch is.  l2 = true: read the list through L2 (used before the PDL wait, when
// this SM's L1 may still hold last frame's list).
template <bool LIST>
__device__ __forceinline__ uint32_t tile_chunk(const RingSchedule& sc, uint32_t t, bool l2 = false) {
    if constexpr (LIST) return l2 ? __ldcg(sc.vis_list + t) : __ldg(sc.vis_list + t);
    else return uint32_t((uint64_t(t) * sc.perm_mul) % sc.n_chunks);
}

// Producer / consumer skeleton shared by both passes.  consume(p, chunk, first, rot, valid): p[s] is record
// chunk * kChunkPoints + first + (s ^ rot) of the cloud and exists iff (s ^ rot) < valid.
//
// The k-th tile a CTA takes lands in stage k % kRingStages.  Its first kRingStages tiles are tiles blockIdx.x +
// k * gridDim.x of the launch.  What it takes after those is
//   * round-robin (tile blockIdx.x + k * gridDim.x) when sc.tile_counter is null — stream-all passes, ring_dynamic = 0;
//   * claimed (list passes, default): ncu of the round-robin kernels showed an SM busy for 74 K ... 102 K of the 110 K
//     cycles of a z-min pass although every CTA gets an even sample of the list — the SMs do not run equally fast — so
//     tiles should go to whoever is free.  The remaining tiles are dealt into sc.n_queues queues (tile 6G + c * Q + q is
//     entry c of queue q), each with its own counter on its own 128-byte line; a consumer group belongs to queue
//     (2 * blockIdx.x + group) mod Q and its refilling thread claims entry c ONE ITERATION BEFORE the refill that
//     streams it, so the counter's round trip is covered by a whole tile of arithmetic.  Two things made earlier
//     attempts slower than round-robin (profiles/r01j_exp_ring_dynamic.json): ptxas warp-aggregates atomicAdd on a
//     uniform address (vote + ATOMG + SHFL of the result: the shuffle waits for the round trip on the spot), and
//     12.5 K claims per pass on ONE address serialise in L2 until the refills run late (44 % of the blend pass's stall
//     samples on the wait for the tile's bytes).  With 4-16 queues: z-min 57 -> 54 us, blend 59 -> 56 us on C3.
// A stage that gets no tile because the launch has run out of them is marked kNoTile and its `full` barrier completed
// by a plain arrive; a group stops at the first such stage (claims are handed out in increasing order, so every
// tile streamed for the group lies before it in the ring).
// (Leaving the z-min pass's last chunks in L2 for a backwards blend pass was measured slower: the streamed lines
// displace the z-buffer / accumulator lines the REDs need.  profiles/r01h_exp_ring_dynamic.json)
//
// early: the tile list is older than the previous grid (the blend pass: the list was built before the z-min pass),
// so the first kRingStages chunks of the CTA are requested BEFORE the PDL wait and land while the previous grid drains.
// Either way this function executes the PDL prologue exactly once for every thread.
constexpr uint32_t kNoTile = 0xFFFFFFFFu;

template <bool LIST, typename Consume>
__device__ __forceinline__ void ring_walk(const PointRecord* __restrict__ pts, uint64_t n, const RingSchedule& sc,
                                          RingSmem& sm, const bool early, Consume&& consume) {
    if (!early) pdl_prologue();
    uint32_t n_tiles = sc.n_chunks;
    if constexpr (LIST) {
        const uint32_t* nv = sc.cull->n_visible;
        n_tiles = early ? __ldcg(nv + (__ldcg(&sc.cull->parity) & 1u)) : cull_count(sc.cull);
    }
    const uint32_t G = gridDim.x;
    // Group g takes k = g, g + kRingGroups, ...; its thread 0 is also the producer of those tiles: once every warp of
    // the group has copied a tile's records into registers (the stage's `empty` barrier) it streams another tile
    // into the freed stage, so each group always has kRingStages / kRingGroups chunks in flight or landed.
    const uint32_t group = threadIdx.x / kRingConsumers, tid = threadIdx.x % kRingConsumers;
    const uint32_t lane = threadIdx.x & 31u, rot = (lane >> 1) & 3u;
    const bool leader = tid == 0;
    // Claiming pays once a CTA has many tiles to even out; with a dozen or fewer (C2: 20 M points, 1280x720) the claims'
    // atomics cost 0.5 us per pass and buy nothing (profiles/r01j_exp_ring_dynamic.json, section 9).  n_tiles is the same
    // word for every CTA of the launch, so they all decide alike.
    const bool claim_tiles = LIST && sc.tile_counter != nullptr && n_tiles > sc.claim_min_tiles_per_cta * G;
    uint64_t policy = 0;
    auto issue = [&](uint32_t stage, uint32_t chunk) {
        sm.chunk[stage] = chunk;
        const uint64_t first = uint64_t(chunk) * kChunkPoints;
        const uint64_t left = n - first;
        const uint32_t bytes = uint32_t(left < uint64_t(kChunkPoints) ? left : uint64_t(kChunkPoints)) * uint32_t(sizeof(PointRecord));
        mbar_arrive_expect_tx(&sm.full[stage], bytes);
        bulk_load(&sm.rec[stage][0], pts + first, bytes, &sm.full[stage], policy);
    };
    auto end_mark = [&](uint32_t stage) {
        sm.chunk[stage] = kNoTile;
        mbar_arrive(&sm.full[stage]);
    };
    if (leader) {
        policy = l2_policy_evict_first();
#pragma unroll
        for (uint32_t k = group; k < uint32_t(kRingStages); k += kRingGroups) {
            const uint32_t t = blockIdx.x + k * G;
            if (t < n_tiles) issue(k, tile_chunk<LIST>(sc, t, early));
            else end_mark(k);
        }
    }
    if (early) pdl_prologue();
    // The tiles beyond the CTAs' first ring-fulls are dealt round-robin into n_queues queues, each with its own counter
    // (a different 128-byte line each: atomics on one address serialise in L2, and 12 K claims per pass through one
    // counter held the refills up); a group claims from queue (2 * blockIdx.x + group) mod n_queues only, so a queue
    // is shared by groups of 592 / n_queues different CTAs spread over the chip.
    // claim = how many tiles of its queue had been claimed before this group's next one.
    const uint32_t Q = sc.n_queues, queue = ring_queue_of(blockIdx.x, group, kRingGroups, Q);
    uint32_t* const counter = sc.tile_counter + queue * kTileQueueStride;
    uint32_t claim = 0u;
    if (claim_tiles && leader) claim = claim_tile(counter);
    // Thread i of a group owns records 4i..4i+3 of the chunk.  A 128-bit LDS is served 8 lanes at a time; lane l reads
    // its record s ^ ((l >> 1) & 3) at step s, so that the 8 lanes of a phase touch 8 different 16-byte bank groups
    // (address/16 mod 8 = 4(l&1) + (s ^ (l>>1 & 3))): conflict-free without padding, one XOR per load.
    const uint32_t lds0 = smem_addr(&sm.rec[0][0]) + ((tid * kRingPerThread + rot) << 4);
    const uint32_t last_chunk = uint32_t((n - 1) / kChunkPoints);
    // how many of this thread's four records exist in the LAST chunk of the cloud (stale bytes follow them in the stage)
    const uint64_t tail_first = uint64_t(last_chunk) * kChunkPoints + tid * kRingPerThread;
    const uint32_t tail_valid = tail_first + kRingPerThread <= n ? uint32_t(kRingPerThread) : (tail_first < n ? uint32_t(n - tail_first) : 0u);
    for (uint32_t k = group;; k += kRingGroups) {
        const uint32_t stage = k % kRingStages, parity = (k / kRingStages) & 1u;
        uint32_t t_refill = kNoTile, refill_chunk = 0u;
        if (leader) {
            if (claim_tiles) {
                if (claim < n_tiles) t_refill = ring_claimed_tile(G, kRingStages, Q, queue, claim);
            } else {
                t_refill = blockIdx.x + (k + uint32_t(kRingStages)) * G;
            }
            if (t_refill < n_tiles) {
                refill_chunk = tile_chunk<LIST>(sc, t_refill);               // in flight during the wait below
                if (claim_tiles) claim = claim_tile(counter);    // for the refill after this one: back by then
            } else {
                t_refill = kNoTile;  // and nothing more to claim: every later refill of this group is an end mark too
                claim = kNoTile;
            }
        }
        mbar_wait(&sm.full[stage], parity);
        const uint32_t chunk = sm.chunk[stage];
        if (chunk == kNoTile) break;  // the whole group reads the same word
        const uint32_t a = lds0 + stage * uint32_t(kChunkPoints * sizeof(PointRecord));
        PointRecord p[kRingPerThread];
#pragma unroll
        for (int s = 0; s < kRingPerThread; ++s) {
            uint4 v;
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a ^ (uint32_t(s) << 4)));
            p[s].x = __uint_as_float(v.x); p[s].y = __uint_as_float(v.y); p[s].z = __uint_as_float(v.z); p[s].bgra = v.w;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.empty[stage]);
        if (leader) {
            mbar_wait(&sm.empty[stage], parity);  // the group's other warps are a few instructions behind at most
            if (t_refill != kNoTile) issue(stage, refill_chunk);
            else end_mark(stage);
        }
        consume(p, chunk, tid * kRingPerThread, rot, chunk == last_chunk ? tail_valid : uint32_t(kRingPerThread));
    }
}

__device__ __forceinline__ RingSmem& ring_setup() {
    extern __shared__ __align__(128) unsigned char ring_raw[];
    RingSmem& sm = *reinterpret_cast<RingSmem*>(ring_raw);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < kRingStages; ++s) {
            mbar_init(&sm.full[s], 1u);                     // the producer's arrive.expect_tx
            mbar_init(&sm.empty[s], kRingConsumers / 32u);  // one arrival per warp of the group that consumes the stage
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    return sm;
}

// ---------------------------------------------------------------- z-min
// VARIANT bit 0: early depth test, bit 2: through L1 (ld.ca), bit 3: measurement only — no RED issued,
// bit 5: measurement only — no in-register merge of same-pixel neighbours.
template <int VARIANT, bool DISTORT, int KEY64, bool LIST>
__global__ void __launch_bounds__(kRingThreads, kRingMinCtas) zmin_ring_kernel(const PointRecord* __restrict__ pts, uint64_t n,
                                                                 uint64_t index_base,
                                                                 const __grid_constant__ ProjParams pp,
                                                                 const __grid_constant__ RingSchedule sc,
                                                                 uint32_t* __restrict__ zbuf,
                                                                 unsigned long long* __restrict__ zkey) {
    RingSmem& sm = ring_setup();  // touches shared memory only: overlaps the previous grid's tail
    ring_walk<LIST>(pts, n, sc, sm, false, [&](const PointRecord (&p)[kRingPerThread], uint32_t chunk, uint32_t first, uint32_t rot, uint32_t valid) {
        uint32_t pix[kRingPerThread];
        using Key = std::conditional_t<KEY64 != 0, unsigned long long, uint32_t>;
        Key key[kRingPerThread];  // KEY64: (depth bits << 32) | global index ; else the depth bits
        bool live[kRingPerThread];
        {
            const float x[4] = {p[0].x, p[1].x, p[2].x, p[3].x}, y[4] = {p[0].y, p[1].y, p[2].y, p[3].y}, z[4] = {p[0].z, p[1].z, p[2].z, p[3].z};
            float depth[4];
            project4<DISTORT>(pp, x, y, z, pix, depth, live);
#pragma unroll
            for (int s = 0; s < kRingPerThread; ++s) {
                const uint32_t slot = uint32_t(s) ^ rot;
                live[s] = live[s] & (slot < valid);
                key[s] = __float_as_uint(depth[s]);
                if constexpr (KEY64)
                    key[s] = (key[s] << 32) | static_cast<unsigned long long>(uint32_t(index_base + uint64_t(chunk) * kChunkPoints + first + slot));
            }
            (void)index_base; (void)chunk; (void)first;
        }
        // a warp whose 128 records all fell outside the frustum is done (most warps of a stream-all pass, the rim of a culled one)
        if (!__any_sync(0xFFFFFFFFu, live[0] | live[1] | live[2] | live[3])) return;
        // neighbours that landed in the same pixel: keep the smallest key in the first of them
#pragma unroll
        for (int j = 1; j < ((VARIANT & 32) ? 0 : kRingPerThread); ++j) {
#pragma unroll
            for (int i = 0; i < j; ++i) {
                const bool same = live[i] & live[j] & (pix[i] == pix[j]);
                if (same) key[i] = key[j] < key[i] ? key[j] : key[i];
                live[j] = live[j] & !same;
            }
        }
        if constexpr (KEY64) {
            unsigned long long cur[kRingPerThread];
#pragma unroll
            for (int s = 0; s < kRingPerThread; ++s) {
                cur[s] = ~0ull;
                if ((VARIANT & 1) && live[s]) cur[s] = (VARIANT & 4) ? __ldca(zkey + pix[s]) : __ldcg(zkey + pix[s]);
            }
#pragma unroll
            for (int s = 0; s < kRingPerThread; ++s)
                if (live[s] && key[s] < cur[s] && !(VARIANT & 8)) red_min_u64(zkey + pix[s], key[s]);
        } else {
            uint32_t cur[kRingPerThread];
#pragma unroll
            for (int s = 0; s < kRingPerThread; ++s) {
                cur[s] = 0xFFFFFFFFu;
                if ((VARIANT & 1) && live[s]) cur[s] = (VARIANT & 4) ? __ldca(zbuf + pix[s]) : __ldcg(zbuf + pix[s]);
            }
#pragma unroll
            for (int s = 0; s < kRingPerThread; ++s) {
                if (live[s] && key[s] < cur[s]) {
                    if constexpr (VARIANT & 8) {  // measurement only (results are wrong)
                        if (key[s] == 0x12345678u && pix[s] == 0xFFFFFFFFu) zbuf[0] = 0u;
                    } else {
                        red_min_u32(zbuf + pix[s], key[s]);
                    }
                }
            }
        }
    });
}

// ---------------------------------------------------------------- blend
// VARIANT bit 2: float accumulators, one RED.ADD.F32x4 per (thread, pixel); else two RED.ADD.64 on the
// reference's 4 x u32 layout.
template <int VARIANT, bool DISTORT, bool LIST>
__global__ void __launch_bounds__(kRingThreads, kRingMinCtas) blend_ring_kernel(const PointRecord* __restrict__ pts, uint64_t n,
                                                                  const __grid_constant__ ProjParams pp,
                                                                  const __grid_constant__ RingSchedule sc,
                                                                  const uint32_t* __restrict__ zbuf,
                                                                  unsigned long long* __restrict__ accum2) {
    RingSmem& sm = ring_setup();
    // early: the visible list this pass walks is older than the grid in front of it (the z-min pass or the merge of
    // the same frame), so its first chunks are requested before the PDL wait
    ring_walk<LIST>(pts, n, sc, sm, LIST && sc.early != 0u, [&](const PointRecord (&p)[kRingPerThread], uint32_t, uint32_t, uint32_t rot, uint32_t valid) {
        uint32_t pix[kRingPerThread];
        float depth[kRingPerThread];
        bool live[kRingPerThread];
        {
            const float x[4] = {p[0].x, p[1].x, p[2].x, p[3].x}, y[4] = {p[0].y, p[1].y, p[2].y, p[3].y}, z[4] = {p[0].z, p[1].z, p[2].z, p[3].z};
            project4<DISTORT>(pp, x, y, z, pix, depth, live);
#pragma unroll
            for (int s = 0; s < kRingPerThread; ++s) live[s] = live[s] & ((uint32_t(s) ^ rot) < valid);
        }
        if (!__any_sync(0xFFFFFFFFu, live[0] | live[1] | live[2] | live[3])) return;  // nothing of this warp is in the frustum
        uint32_t zmin[kRingPerThread];
#pragma unroll
        for (int s = 0; s < kRingPerThread; ++s) {
            zmin[s] = 0u;
            if (live[s]) zmin[s] = __ldg(zbuf + pix[s]);
        }
        uint32_t b[kRingPerThread], g[kRingPerThread], r[kRingPerThread], c[kRingPerThread];
#pragma unroll
        for (int s = 0; s < kRingPerThread; ++s) {
            const float lim = __fadd_rn(__uint_as_float(zmin[s]), kDepthWindow);
            live[s] = live[s] & !(depth[s] > lim);  // render.cu:106 (NaN depth is accepted, as there)
            b[s] = p[s].bgra & 0xFFu; g[s] = (p[s].bgra >> 8) & 0xFFu; r[s] = (p[s].bgra >> 16) & 0xFFu; c[s] = 1u;
        }
        // accepted neighbours of the same pixel: sum their bytes into the first of them (integer, exact)
#pragma unroll
        for (int j = 1; j < ((VARIANT & 32) ? 0 : kRingPerThread); ++j) {
#pragma unroll
            for (int i = 0; i < j; ++i) {
                const bool same = live[i] & live[j] & (pix[i] == pix[j]);
                if (same) { b[i] += b[j]; g[i] += g[j]; r[i] += r[j]; c[i] += c[j]; }
                live[j] = live[j] & !same;
            }
        }
#pragma unroll
        for (int s = 0; s < kRingPerThread; ++s) {
            if (live[s]) {
                unsigned long long* a = accum2 + uint64_t(pix[s]) * 2;
                if constexpr (VARIANT & 4) {
                    red_add_f32x4(a, float(b[s]), float(g[s]), float(r[s]), float(c[s]));
                } else {
                    red_add_u64(a + 0, static_cast<unsigned long long>(b[s]) | (static_cast<unsigned long long>(g[s]) << 32));
                    red_add_u64(a + 1, static_cast<unsigned long long>(r[s]) | (static_cast<unsigned long long>(c[s]) << 32));
                }
            }
        }
    });
}

// ---------------------------------------------------------------- host launchers
RingSchedule make_ring_schedule(uint64_t n_points, const CullState* cull, const uint32_t* vis_list) {
    RingSchedule sc;
    sc.cull = cull;
    sc.vis_list = vis_list;
    sc.early = 1;
    sc.tile_counter = nullptr;
    sc.n_queues = 1;
    sc.ctas_per_sm = kRingCtasPerSm;
    sc.claim_min_tiles_per_cta = 12;
    sc.n_chunks = uint32_t((n_points + kChunkPoints - 1) / kChunkPoints);
    // golden-ratio stride, made coprime with n_chunks: t -> (t * mul) mod n_chunks is a permutation whose every
    // window of consecutive t is spread evenly over the cloud
    uint64_t mul = uint64_t(double(sc.n_chunks) * 0.6180339887498949) | 1u;
    auto gcd = [](uint64_t a, uint64_t b) { while (b) { const uint64_t t = a % b; a = b; b = t; } return a; };
    while (sc.n_chunks > 1 && gcd(mul, sc.n_chunks) != 1) mul += 2;
    sc.perm_mul = sc.n_chunks > 1 ? uint32_t(mul % sc.n_chunks) : 0u;
    if (sc.n_chunks > 1 && sc.perm_mul == 0) sc.perm_mul = 1;
    return sc;
}

// The opt-in to > 48 KB of dynamic shared memory is per (kernel, device): once for each.
template <typename K>
static cudaError_t ring_attr(K kernel, bool* done) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64 && done[dev]) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sizeof(RingSmem)));
    if (e == cudaSuccess && dev >= 0 && dev < 64) done[dev] = true;
    return e;
}
static unsigned ring_grid(int sm_count, const RingSchedule& sc, bool list) {
    unsigned grid = unsigned(sm_count) * (sc.ctas_per_sm >= 1u && sc.ctas_per_sm <= unsigned(kRingCtasPerSm) ? sc.ctas_per_sm : unsigned(kRingCtasPerSm));
    if (!list && sc.n_chunks < grid) grid = sc.n_chunks ? sc.n_chunks : 1u;
    return grid;
}

#define RTR_RING_LAUNCH(KERNEL, ...)                                                                            \
    do {                                                                                                        \
        static bool attr_done[64] = {false};                                                                    \
        const cudaError_t attr_status = ring_attr(KERNEL, attr_done);                                           \
        if (attr_status != cudaSuccess) return attr_status;                                                     \
        launch_pdl_smem(KERNEL, dim3(grid), dim3(kRingThreads), sizeof(RingSmem), s, __VA_ARGS__);              \
    } while (0)

template <int VARIANT>
static cudaError_t launch_zmin_ring_v(cudaStream_t s, unsigned grid, const PointRecord* pts, uint64_t n, uint64_t index_base,
                                      const ProjParams& pp, const RingSchedule& sc, bool list, uint32_t* zbuf,
                                      unsigned long long* zkey) {
    if (list && pp.distort) {
        if (zkey) RTR_RING_LAUNCH((zmin_ring_kernel<VARIANT, true, 1, true>), pts, n, index_base, pp, sc, zbuf, zkey);
        else RTR_RING_LAUNCH((zmin_ring_kernel<VARIANT, true, 0, true>), pts, n, index_base, pp, sc, zbuf, zkey);
    } else if (list) {
        if (zkey) RTR_RING_LAUNCH((zmin_ring_kernel<VARIANT, false, 1, true>), pts, n, index_base, pp, sc, zbuf, zkey);
        else RTR_RING_LAUNCH((zmin_ring_kernel<VARIANT, false, 0, true>), pts, n, index_base, pp, sc, zbuf, zkey);
    } else if (pp.distort) {
        if (zkey) RTR_RING_LAUNCH((zmin_ring_kernel<VARIANT, true, 1, false>), pts, n, index_base, pp, sc, zbuf, zkey);
        else RTR_RING_LAUNCH((zmin_ring_kernel<VARIANT, true, 0, false>), pts, n, index_base, pp, sc, zbuf, zkey);
    } else {
        if (zkey) RTR_RING_LAUNCH((zmin_ring_kernel<VARIANT, false, 1, false>), pts, n, index_base, pp, sc, zbuf, zkey);
        else RTR_RING_LAUNCH((zmin_ring_kernel<VARIANT, false, 0, false>), pts, n, index_base, pp, sc, zbuf, zkey);
    }
    return cudaGetLastError();
}

void ring_geometry(uint32_t* stages, uint32_t* groups_per_cta, uint32_t* ctas_per_sm) {
    *stages = kRingStages;
    *groups_per_cta = kRingGroups;
    *ctas_per_sm = kRingCtasPerSm;
}

cudaError_t launch_zmin_ring(cudaStream_t s, int sm_count, int variant, const PointRecord* pts, uint64_t n,
                             uint64_t index_base, const ProjParams& pp, const RingSchedule& sc_in, bool list, uint32_t* zbuf,
                             unsigned long long* zkey) {
    if (n == 0) return cudaSuccess;
    const unsigned grid = ring_grid(sm_count, sc_in, list);
    RingSchedule sc = sc_in;
    sc.n_queues = ring_effective_queues(grid, kRingGroups, sc.n_queues);
    switch (variant & 45) {  // bit 1 (warp aggregation) has no ring form: the in-register merge replaces it
        case 37: return launch_zmin_ring_v<37>(s, grid, pts, n, index_base, pp, sc, list, zbuf, zkey);
        case 0: return launch_zmin_ring_v<0>(s, grid, pts, n, index_base, pp, sc, list, zbuf, zkey);
        case 1: return launch_zmin_ring_v<1>(s, grid, pts, n, index_base, pp, sc, list, zbuf, zkey);
        case 5: return launch_zmin_ring_v<5>(s, grid, pts, n, index_base, pp, sc, list, zbuf, zkey);
        case 8: return launch_zmin_ring_v<8>(s, grid, pts, n, index_base, pp, sc, list, zbuf, zkey);
        case 9: return launch_zmin_ring_v<9>(s, grid, pts, n, index_base, pp, sc, list, zbuf, zkey);
        case 13: return launch_zmin_ring_v<13>(s, grid, pts, n, index_base, pp, sc, list, zbuf, zkey);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_blend_ring(cudaStream_t s, int sm_count, int variant, const PointRecord* pts, uint64_t n,
                              const ProjParams& pp, const RingSchedule& sc_in, bool list, const uint32_t* zbuf, uint32_t* accum) {
    if (n == 0) return cudaSuccess;
    const unsigned grid = ring_grid(sm_count, sc_in, list);
    RingSchedule sc = sc_in;
    sc.n_queues = ring_effective_queues(grid, kRingGroups, sc.n_queues);
    unsigned long long* a2 = reinterpret_cast<unsigned long long*>(accum);
    const bool f32 = (variant & 4) != 0;
    if (list && !pp.distort && f32 && (variant & 32)) {  // measurement: no in-register merge
        RTR_RING_LAUNCH((blend_ring_kernel<36, false, true>), pts, n, pp, sc, zbuf, a2);
        return cudaGetLastError();
    }
    if (list && pp.distort) {
        if (f32) RTR_RING_LAUNCH((blend_ring_kernel<4, true, true>), pts, n, pp, sc, zbuf, a2);
        else RTR_RING_LAUNCH((blend_ring_kernel<0, true, true>), pts, n, pp, sc, zbuf, a2);
    } else if (list) {
        if (f32) RTR_RING_LAUNCH((blend_ring_kernel<4, false, true>), pts, n, pp, sc, zbuf, a2);
        else RTR_RING_LAUNCH((blend_ring_kernel<0, false, true>), pts, n, pp, sc, zbuf, a2);
    } else if (pp.distort) {
        if (f32) RTR_RING_LAUNCH((blend_ring_kernel<4, true, false>), pts, n, pp, sc, zbuf, a2);
        else RTR_RING_LAUNCH((blend_ring_kernel<0, true, false>), pts, n, pp, sc, zbuf, a2);
    } else {
        if (f32) RTR_RING_LAUNCH((blend_ring_kernel<4, false, false>), pts, n, pp, sc, zbuf, a2);
        else RTR_RING_LAUNCH((blend_ring_kernel<0, false, false>), pts, n, pp, sc, zbuf, a2);
    }
    return cudaGetLastError();
}

}  // namespace rtr
