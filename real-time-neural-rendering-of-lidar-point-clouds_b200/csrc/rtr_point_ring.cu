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
#include <cstdlib>
#include <type_traits>

#include "rtr_kernels.h"

namespace rtr {

#ifndef RTR_RING_STAGES
#define RTR_RING_STAGES 6
#endif
constexpr int kRingStages = RTR_RING_STAGES;
// Lane -> record mapping of a warp's 128 records (build knob; A/B in profiles/r02Q_exp_lane_map.json):
//   0: thread l holds records 4l .. 4l+3 (Morton neighbours, merged in registers; a warp instruction carries records
//      4 apart, spread over the whole 128-record footprint);
//   1: thread l holds records l, 32+l, 64+l, 96+l: a warp instruction carries 32 CONSECUTIVE records — about half as
//      many distinct 32-byte sectors per gather / RED instruction (profiles/r01h_merge_group_sizes.json: 6.7 vs 12.1 for
//      the z-buffer, 12.7 vs 21.4 for the colour sums) — and same-pixel neighbours sit in adjacent LANES, where a
//      segmented shuffle scan over runs of at most RTR_RUN_MAX lanes merges them (min / integer sums: exact, order-free).
#ifndef RTR_LANE_MAP
#define RTR_LANE_MAP 0
#endif
#ifndef RTR_RUN_MAX
#define RTR_RUN_MAX 8
#endif
static_assert(RTR_RUN_MAX == 1 || RTR_RUN_MAX == 2 || RTR_RUN_MAX == 4 || RTR_RUN_MAX == 8 || RTR_RUN_MAX == 16 || RTR_RUN_MAX == 32, "run length cap: a power of two");
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
__device__ __forceinline__ void mbar_arrive_at(uint32_t bar_smem_addr) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_smem_addr) : "memory");
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
        sm.chunk[stage] = chunk;  // the list entry as it is: two-camera lists carry their pass flags in the top bits
        const uint64_t first = uint64_t(chunk & kTileIdMask) * kChunkPoints;
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
#if RTR_LANE_MAP
    const uint32_t first_rec = (tid & ~31u) * kRingPerThread + lane;   // slot s is record first_rec + 32 s of the chunk
    const uint32_t lds0 = smem_addr(&sm.rec[0][0]) + (first_rec << 4);   // 32 lanes x 16 B in a row: conflict-free as it is
#else
    const uint32_t lds0 = smem_addr(&sm.rec[0][0]) + ((tid * kRingPerThread + rot) << 4);
#endif
    const uint32_t last_chunk = uint32_t((n - 1) / kChunkPoints);
    // how many of this thread's four records exist in the LAST chunk of the cloud (stale bytes follow them in the stage)
#if RTR_LANE_MAP
    const uint32_t tail_valid = uint32_t(n - uint64_t(last_chunk) * kChunkPoints);   // records of the last chunk: slot s exists iff first_rec + 32 s < valid
#else
    const uint64_t tail_first = uint64_t(last_chunk) * kChunkPoints + tid * kRingPerThread;
    const uint32_t tail_valid = tail_first + kRingPerThread <= n ? uint32_t(kRingPerThread) : (tail_first < n ? uint32_t(n - tail_first) : 0u);
#endif
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
#if RTR_LANE_MAP
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a + (uint32_t(s) << 9)));
#else
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a ^ (uint32_t(s) << 4)));
#endif
            p[s].x = __uint_as_float(v.x); p[s].y = __uint_as_float(v.y); p[s].z = __uint_as_float(v.z); p[s].bgra = v.w;
        }
        // The stage may only be handed back once the records are IN the registers.  Issuing the four LDS is not enough:
        // the mbarrier arrive overtakes them when the SM's load/store queue is backed up by scattered REDs and gathers
        // (an unsorted cloud: hundreds of sector operations per warp instruction), the leader sees eight arrivals, re-arms
        // the stage, and the bulk copy of the NEXT tile lands under loads that have not read the banks yet — a few
        // records of the next tile processed twice, a few of this one never (found as frames that differ from run to run
        // in a handful of pixels, tools/diag_repeat.py: 9 % of the frames of an unsorted 16 M-point cloud, 92 % with every
        // chunk streamed; never with the per-thread kernels).  The arrive's address therefore DEPENDS on the loaded
        // words — and-ed with a kernel parameter that is always 0, which ptxas cannot fold away — so the instruction waits
        // on the loads' scoreboard (SASS: LOP3 on the four LDS destinations in front of SYNCS.ARRIVE).
        const uint32_t landed = (p[0].bgra ^ p[1].bgra ^ p[2].bgra ^ p[3].bgra) & sc.zero;
        __syncwarp();
        if (lane == 0) mbar_arrive_at(smem_addr(&sm.empty[stage]) + landed);
        if (leader) {
            mbar_wait(&sm.empty[stage], parity);  // the group's other warps are a few instructions behind at most
            if (t_refill != kNoTile) issue(stage, refill_chunk);
            else end_mark(stage);
        }
#if RTR_LANE_MAP
        consume(p, chunk, first_rec, 0u, (chunk & kTileIdMask) == last_chunk ? tail_valid : uint32_t(kChunkPoints));
#else
        consume(p, chunk, tid * kRingPerThread, rot, (chunk & kTileIdMask) == last_chunk ? tail_valid : uint32_t(kRingPerThread));
#endif
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

// ---------------------------------------------------------------- the two passes' per-thread bodies
// Each body is split into "project" (arithmetic + in-register merge), "gather" (the z-buffer loads are issued) and
// "commit" (compare + REDs), so that the fused kernel can have both passes' gathers in flight before either is consumed.
//
// z-min.  VARIANT bit 0: early depth test, bit 2: through L1 (ld.ca); RTR_EXPERIMENTS builds only: bit 3: no RED
// issued (results are wrong), bit 5: no in-register merge of same-pixel neighbours.
// index of slot s within its chunk, and whether the cloud has such a record (see ring_walk's consume)
__device__ __forceinline__ uint32_t slot_record(uint32_t first, uint32_t rot, int s) {
#if RTR_LANE_MAP
    (void)rot;
    return first + 32u * uint32_t(s);
#else
    return first + (uint32_t(s) ^ rot);
#endif
}
__device__ __forceinline__ bool slot_exists(uint32_t first, uint32_t rot, uint32_t valid, int s) {
#if RTR_LANE_MAP
    (void)rot;
    return first + 32u * uint32_t(s) < valid;
#else
    (void)first;
    return (uint32_t(s) ^ rot) < valid;
#endif
}

#if RTR_LANE_MAP
// Runs of adjacent lanes that carry the same pixel in one warp instruction (cut at multiples of RTR_RUN_MAX lanes so
// that log2(RTR_RUN_MAX) scan steps cover a run): dist = lanes between this lane and its run's first, last = this lane
// ends its run.  A lane that is not `on` is a run of its own.  Executed by the whole warp.
struct LaneRun { uint32_t dist; bool last; };
__device__ __forceinline__ LaneRun lane_run(uint32_t pix, bool on) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t p = on ? pix : 0xFFFFFFFFu;
    const uint32_t prev = __shfl_up_sync(0xFFFFFFFFu, p, 1);
    const bool head = !on | (prev != p) | ((lane & uint32_t(RTR_RUN_MAX - 1)) == 0u);
    const uint32_t heads = __ballot_sync(0xFFFFFFFFu, head);   // bit 0 is always set
    LaneRun r;
    r.dist = lane - (31u - uint32_t(__clz(int(heads & (0xFFFFFFFFu >> (31u - lane))))));
    r.last = lane == 31u || ((heads >> (lane + 1u)) & 1u) != 0u;
    return r;
}
#endif

template <int KEY64>
struct ZminLanes {
    using Key = std::conditional_t<KEY64 != 0, unsigned long long, uint32_t>;
    uint32_t pix[kRingPerThread];
    Key key[kRingPerThread], cur[kRingPerThread];  // KEY64: (depth bits << 32) | global index ; else the depth bits
    bool live[kRingPerThread];
    bool any;  // some record of the WARP is in the frustum
};
template <int VARIANT, bool DISTORT, int KEY64>
__device__ __forceinline__ void zmin_project(ZminLanes<KEY64>& z, const PointRecord (&p)[kRingPerThread], const ProjParams& pp,
                                             uint32_t chunk, uint32_t first, uint32_t rot, uint32_t valid, uint64_t index_base) {
    const float x[4] = {p[0].x, p[1].x, p[2].x, p[3].x}, y[4] = {p[0].y, p[1].y, p[2].y, p[3].y}, zz[4] = {p[0].z, p[1].z, p[2].z, p[3].z};
    float depth[4];
    project4<DISTORT>(pp, x, y, zz, z.pix, depth, z.live);
#pragma unroll
    for (int s = 0; s < kRingPerThread; ++s) {
        z.live[s] = z.live[s] & slot_exists(first, rot, valid, s);
        z.key[s] = __float_as_uint(depth[s]);
        if constexpr (KEY64)
            z.key[s] = (z.key[s] << 32) | static_cast<unsigned long long>(uint32_t(index_base + uint64_t(chunk) * kChunkPoints + slot_record(first, rot, s)));
    }
    (void)index_base; (void)chunk; (void)first;
    // a warp whose 128 records all fell outside the frustum is done (most warps of a stream-all pass, the rim of a culled one)
    z.any = __any_sync(0xFFFFFFFFu, z.live[0] | z.live[1] | z.live[2] | z.live[3]);
    if (!z.any) return;
#if RTR_LANE_MAP
    // neighbours that landed in the same pixel sit in adjacent lanes: the last lane of a run keeps the run's smallest key
    if constexpr (!(VARIANT & 32) && RTR_RUN_MAX > 1) {
#pragma unroll
        for (int s = 0; s < kRingPerThread; ++s) {
            const LaneRun run = lane_run(z.pix[s], z.live[s]);
#pragma unroll
            for (uint32_t d = 1; d < uint32_t(RTR_RUN_MAX); d <<= 1) {
                const typename ZminLanes<KEY64>::Key o = __shfl_up_sync(0xFFFFFFFFu, z.key[s], d);
                if (run.dist >= d) z.key[s] = o < z.key[s] ? o : z.key[s];
            }
            z.live[s] = z.live[s] & run.last;
        }
    }
#endif
    // neighbours that landed in the same pixel: keep the smallest key in the first of them
#pragma unroll
    for (int j = 1; j < (((VARIANT & 32) || RTR_LANE_MAP) ? 0 : kRingPerThread); ++j) {
#pragma unroll
        for (int i = 0; i < j; ++i) {
            const bool same = z.live[i] & z.live[j] & (z.pix[i] == z.pix[j]);
            if (same) z.key[i] = z.key[j] < z.key[i] ? z.key[j] : z.key[i];
            z.live[j] = z.live[j] & !same;
        }
    }
}
template <int VARIANT, int KEY64>
__device__ __forceinline__ void zmin_gather(ZminLanes<KEY64>& z, const uint32_t* __restrict__ zbuf, const unsigned long long* __restrict__ zkey) {
    if (!z.any) return;
#pragma unroll
    for (int s = 0; s < kRingPerThread; ++s) {
        z.cur[s] = ~typename ZminLanes<KEY64>::Key(0);
        if constexpr (KEY64) {
            if ((VARIANT & 1) && z.live[s]) z.cur[s] = (VARIANT & 4) ? __ldca(zkey + z.pix[s]) : __ldcg(zkey + z.pix[s]);
        } else {
            if ((VARIANT & 1) && z.live[s]) z.cur[s] = (VARIANT & 4) ? __ldca(zbuf + z.pix[s]) : __ldcg(zbuf + z.pix[s]);
        }
    }
}
template <int VARIANT, int KEY64>
__device__ __forceinline__ void zmin_commit(const ZminLanes<KEY64>& z, uint32_t* __restrict__ zbuf, unsigned long long* __restrict__ zkey) {
    if (!z.any) return;
#pragma unroll
    for (int s = 0; s < kRingPerThread; ++s) {
        if (z.live[s] && z.key[s] < z.cur[s]) {
            if constexpr (KEY64) {
                if (!(VARIANT & 8)) red_min_u64(zkey + z.pix[s], z.key[s]);
            } else if constexpr ((VARIANT & 8) != 0) {  // RTR_EXPERIMENTS only: no RED (results are wrong)
                if (z.key[s] == 0x12345678u && z.pix[s] == 0xFFFFFFFFu) zbuf[0] = 0u;
            } else {
                red_min_u32(zbuf + z.pix[s], z.key[s]);
            }
        }
    }
}

// blend.  VARIANT bit 2: float accumulators, one RED.ADD.F32x4 per (thread, pixel); else two RED.ADD.64 on the
// reference's 4 x u32 layout.  RTR_EXPERIMENTS builds only: bit 5: no in-register merge.
struct BlendLanes {
    uint32_t pix[kRingPerThread], zmin[kRingPerThread];
    float depth[kRingPerThread];
    bool live[kRingPerThread];
    bool any;
};
template <bool DISTORT>
__device__ __forceinline__ void blend_project(BlendLanes& b, const PointRecord (&p)[kRingPerThread], const ProjParams& pp, uint32_t first, uint32_t rot, uint32_t valid) {
    const float x[4] = {p[0].x, p[1].x, p[2].x, p[3].x}, y[4] = {p[0].y, p[1].y, p[2].y, p[3].y}, z[4] = {p[0].z, p[1].z, p[2].z, p[3].z};
    project4<DISTORT>(pp, x, y, z, b.pix, b.depth, b.live);
#pragma unroll
    for (int s = 0; s < kRingPerThread; ++s) b.live[s] = b.live[s] & slot_exists(first, rot, valid, s);
    b.any = __any_sync(0xFFFFFFFFu, b.live[0] | b.live[1] | b.live[2] | b.live[3]);  // false: nothing of this warp is in the frustum
}
__device__ __forceinline__ void blend_gather(BlendLanes& b, const uint32_t* __restrict__ zbuf) {
    if (!b.any) return;
#pragma unroll
    for (int s = 0; s < kRingPerThread; ++s) {
        b.zmin[s] = 0u;
        if (b.live[s]) b.zmin[s] = __ldg(zbuf + b.pix[s]);
    }
}
template <int VARIANT>
__device__ __forceinline__ void blend_commit(BlendLanes& l, const PointRecord (&p)[kRingPerThread], unsigned long long* __restrict__ accum2) {
    if (!l.any) return;
    uint32_t b[kRingPerThread], g[kRingPerThread], r[kRingPerThread], c[kRingPerThread];
#pragma unroll
    for (int s = 0; s < kRingPerThread; ++s) {
        const float lim = __fadd_rn(__uint_as_float(l.zmin[s]), kDepthWindow);
        l.live[s] = l.live[s] & !(l.depth[s] > lim);  // render.cu:106 (NaN depth is accepted, as there)
        b[s] = p[s].bgra & 0xFFu; g[s] = (p[s].bgra >> 8) & 0xFFu; r[s] = (p[s].bgra >> 16) & 0xFFu; c[s] = 1u;
    }
#if RTR_LANE_MAP
    // accepted neighbours of the same pixel sit in adjacent lanes: segmented sums (integers: exact, order-free) over each
    // run, two channels per 32-bit word (a run's sums stay below 32 * 255 < 2^16); the run's last lane issues the RED
    if constexpr (!(VARIANT & 32) && RTR_RUN_MAX > 1) {
#pragma unroll
        for (int s = 0; s < kRingPerThread; ++s) {
            const LaneRun run = lane_run(l.pix[s], l.live[s]);
            uint32_t bg = b[s] | (g[s] << 16), rc = r[s] | (c[s] << 16);
#pragma unroll
            for (uint32_t d = 1; d < uint32_t(RTR_RUN_MAX); d <<= 1) {
                const uint32_t o0 = __shfl_up_sync(0xFFFFFFFFu, bg, d), o1 = __shfl_up_sync(0xFFFFFFFFu, rc, d);
                if (run.dist >= d) { bg += o0; rc += o1; }
            }
            b[s] = bg & 0xFFFFu; g[s] = bg >> 16; r[s] = rc & 0xFFFFu; c[s] = rc >> 16;
            l.live[s] = l.live[s] & run.last;
        }
    }
#endif
    // accepted neighbours of the same pixel: sum their bytes into the first of them (integer, exact)
#pragma unroll
    for (int j = 1; j < (((VARIANT & 32) || RTR_LANE_MAP) ? 0 : kRingPerThread); ++j) {
#pragma unroll
        for (int i = 0; i < j; ++i) {
            const bool same = l.live[i] & l.live[j] & (l.pix[i] == l.pix[j]);
            if (same) { b[i] += b[j]; g[i] += g[j]; r[i] += r[j]; c[i] += c[j]; }
            l.live[j] = l.live[j] & !same;
        }
    }
#pragma unroll
    for (int s = 0; s < kRingPerThread; ++s) {
        if (l.live[s]) {
            unsigned long long* a = accum2 + uint64_t(l.pix[s]) * 2;
            if constexpr ((VARIANT & 4) != 0) {
                red_add_f32x4(a, float(b[s]), float(g[s]), float(r[s]), float(c[s]));
            } else {
                red_add_u64(a + 0, static_cast<unsigned long long>(b[s]) | (static_cast<unsigned long long>(g[s]) << 32));
                red_add_u64(a + 1, static_cast<unsigned long long>(r[s]) | (static_cast<unsigned long long>(c[s]) << 32));
            }
        }
    }
}

// ---------------------------------------------------------------- z-min
template <int VARIANT, bool DISTORT, int KEY64, bool LIST>
__global__ void __launch_bounds__(kRingThreads, kRingMinCtas) zmin_ring_kernel(const PointRecord* __restrict__ pts, uint64_t n,
                                                                 uint64_t index_base,
                                                                 const __grid_constant__ ProjParams pp,
                                                                 const __grid_constant__ RingSchedule sc,
                                                                 uint32_t* __restrict__ zbuf,
                                                                 unsigned long long* __restrict__ zkey) {
    RingSmem& sm = ring_setup();  // touches shared memory only: overlaps the previous grid's tail
    ring_walk<LIST>(pts, n, sc, sm, false, [&](const PointRecord (&p)[kRingPerThread], uint32_t entry, uint32_t first, uint32_t rot, uint32_t valid) {
        ZminLanes<KEY64> z;
        zmin_project<VARIANT, DISTORT, KEY64>(z, p, pp, entry & kTileIdMask, first, rot, valid, index_base);
        zmin_gather<VARIANT, KEY64>(z, zbuf, zkey);
        zmin_commit<VARIANT, KEY64>(z, zbuf, zkey);
    });
}

// ---------------------------------------------------------------- z-min with shared-memory tile pre-reduction
// north_star names "shared-memory tile pre-reduction ... to cut L2 contention"; this is that design, as zmin_variant
// bit 6 (results identical, tests/test_gpu_parity.py VARIANTS).  Per tile (one chunk, one consumer group of 256 threads):
//   1. project + merge in registers as above; the group's pixel bounding box by redux.min/max per warp and four
//      shared-memory atomics per warp;
//   2. bar.sync (group).  If the box fits a 32 x 32 window: atom.shared.min of every surviving record into the
//      window, else the direct path (early test + RED to global memory) for this tile;
//   3. bar.sync (group); every thread takes four consecutive window pixels: early depth test against the global
//      z-buffer and ONE red.global.min per touched pixel — a warp's REDs cover 4 window rows of 128 B.
// Windows and boxes are double-buffered per group, so a tile costs two group barriers.  What it buys and costs is
// measured in profiles/r02_exp_smem_tile.json.
constexpr int kWinDim = 32;
struct alignas(16) TileWindows {   // (the windows are initialised and flushed with 16-byte accesses)
    uint32_t win[kRingGroups][2][kWinDim * kWinDim];  // depth bits, 0xFFFFFFFF = untouched
    uint32_t box[kRingGroups][2][4];                  // u min, v min, u max, v max of the tile's live records
};
struct RingSmemTile {
    RingSmem ring;
    TileWindows tw;
};
__device__ __forceinline__ void group_barrier(uint32_t group) {
    asm volatile("bar.sync %0, %1;" ::"r"(1u + group), "r"(uint32_t(kRingConsumers)) : "memory");
}

template <int VARIANT, bool LIST>
__global__ void __launch_bounds__(kRingThreads, kRingMinCtas) zmin_ring_smem_kernel(const PointRecord* __restrict__ pts, uint64_t n,
                                                                      const __grid_constant__ ProjParams pp,
                                                                      const __grid_constant__ RingSchedule sc,
                                                                      uint32_t* __restrict__ zbuf,
                                                                      unsigned long long* __restrict__ stats) {
    extern __shared__ __align__(128) unsigned char ring_raw[];
    RingSmemTile& st = *reinterpret_cast<RingSmemTile*>(ring_raw);
    RingSmem& sm = ring_setup();  // (the ring is the first member)
    const uint32_t group = threadIdx.x / kRingConsumers, tid = threadIdx.x % kRingConsumers, lane = threadIdx.x & 31u;
    const uint32_t W = uint32_t(pp.W);
    // both windows untouched, both boxes empty
    for (int b = 0; b < 2; ++b) {
        reinterpret_cast<uint4*>(st.tw.win[group][b])[tid] = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
        if (tid < 4) st.tw.box[group][b][tid] = tid < 2 ? 0xFFFFFFFFu : 0u;
    }
    group_barrier(group);
    uint32_t it = 0;
    unsigned long long n_fit = 0, n_direct = 0, n_flushed = 0, n_entered = 0;
    ring_walk<LIST>(pts, n, sc, sm, false, [&](const PointRecord (&p)[kRingPerThread], uint32_t entry, uint32_t first, uint32_t rot, uint32_t valid) {
        const uint32_t cur_buf = it & 1u, nxt_buf = cur_buf ^ 1u;
        ++it;
        uint32_t* win = st.tw.win[group][cur_buf];
        uint32_t* box = st.tw.box[group][cur_buf];
        ZminLanes<0> z;
        zmin_project<VARIANT, false, 0>(z, p, pp, entry & kTileIdMask, first, rot, valid, 0ull);   // (z.any false: nothing of the warp is live)
        uint32_t u[kRingPerThread], v[kRingPerThread];
        uint32_t umin = 0xFFFFFFFFu, vmin = 0xFFFFFFFFu, umax = 0u, vmax = 0u;
#pragma unroll
        for (int s = 0; s < kRingPerThread; ++s) {
            z.live[s] = z.live[s] & z.any;
            v[s] = z.pix[s] / W;
            u[s] = z.pix[s] - v[s] * W;
            if (z.live[s]) { umin = min(umin, u[s]); vmin = min(vmin, v[s]); umax = max(umax, u[s]); vmax = max(vmax, v[s]); }
        }
        if (z.any) {
            umin = __reduce_min_sync(0xFFFFFFFFu, umin); vmin = __reduce_min_sync(0xFFFFFFFFu, vmin);
            umax = __reduce_max_sync(0xFFFFFFFFu, umax); vmax = __reduce_max_sync(0xFFFFFFFFu, vmax);
            if (lane == 0) { atomicMin(box + 0, umin); atomicMin(box + 1, vmin); atomicMax(box + 2, umax); atomicMax(box + 3, vmax); }
        }
        group_barrier(group);
        const uint32_t u0 = box[0], v0 = box[1], u1 = box[2], v1 = box[3];
        const bool some = u1 >= u0 && v1 >= v0;                       // the tile has a live record at all
        const bool fits = some && (u1 - u0) < uint32_t(kWinDim) && (v1 - v0) < uint32_t(kWinDim);
        if (fits) {
#pragma unroll
            for (int s = 0; s < kRingPerThread; ++s)
                if (z.live[s]) { atomicMin(win + (v[s] - v0) * kWinDim + (u[s] - u0), z.key[s]); ++n_entered; }
        } else if (some) {
            zmin_gather<VARIANT, 0>(z, zbuf, nullptr);
            zmin_commit<VARIANT, 0>(z, zbuf, nullptr);
        }
        // the other window / box were last read in the previous tile's flush, which every thread of the group finished
        // before it arrived at this tile's first barrier: re-arm them for the next tile
        reinterpret_cast<uint4*>(st.tw.win[group][nxt_buf])[tid] = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
        if (tid < 4) st.tw.box[group][nxt_buf][tid] = tid < 2 ? 0xFFFFFFFFu : 0u;
        group_barrier(group);
        if (tid == 0) { n_fit += fits ? 1u : 0u; n_direct += (some && !fits) ? 1u : 0u; }
        if (fits) {
            // flush: thread t owns window pixels 4t .. 4t+3 (row t / 8, columns 4 (t % 8) ...)
            const uint4 w4 = reinterpret_cast<const uint4*>(win)[tid];
            const uint32_t wk[4] = {w4.x, w4.y, w4.z, w4.w};
            const uint32_t row = v0 + tid / (kWinDim / 4), col = u0 + (tid % (kWinDim / 4)) * 4;
            uint32_t cur[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                cur[k] = 0xFFFFFFFFu;
                if ((VARIANT & 1) && wk[k] != 0xFFFFFFFFu) cur[k] = (VARIANT & 4) ? __ldca(zbuf + row * W + col + k) : __ldcg(zbuf + row * W + col + k);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (wk[k] != 0xFFFFFFFFu) {
                    ++n_flushed;
                    if (wk[k] < cur[k]) red_min_u32(zbuf + row * W + col + k, wk[k]);
                }
        }
    });
    if (stats) {   // measurement support: a handful of atomics per CTA
        n_flushed = __reduce_add_sync(0xFFFFFFFFu, uint32_t(n_flushed));
        n_entered = __reduce_add_sync(0xFFFFFFFFu, uint32_t(n_entered));
        if (tid == 0) { atomicAdd(stats + 0, n_fit); atomicAdd(stats + 1, n_direct); }
        if (lane == 0) { atomicAdd(stats + 2, n_flushed); atomicAdd(stats + 3, n_entered); }
    }
}

// ---------------------------------------------------------------- blend
template <int VARIANT, bool DISTORT, bool LIST>
__global__ void __launch_bounds__(kRingThreads, kRingMinCtas) blend_ring_kernel(const PointRecord* __restrict__ pts, uint64_t n,
                                                                  const __grid_constant__ ProjParams pp,
                                                                  const __grid_constant__ RingSchedule sc,
                                                                  const uint32_t* __restrict__ zbuf,
                                                                  unsigned long long* __restrict__ accum2) {
    RingSmem& sm = ring_setup();
    // early: the visible list this pass walks is older than the grid in front of it (the z-min pass or the merge of
    // the same frame), so its first chunks are requested before the PDL wait
    ring_walk<LIST>(pts, n, sc, sm, LIST && sc.early != 0u, [&](const PointRecord (&p)[kRingPerThread], uint32_t, uint32_t first, uint32_t rot, uint32_t valid) {
        BlendLanes b;
        blend_project<DISTORT>(b, p, pp, first, rot, valid);
        blend_gather(b, zbuf);
        blend_commit<VARIANT>(b, p, accum2);
    });
}

// ---------------------------------------------------------------- both passes over ONE stream of chunks
// Frame k-1's blend and frame k's z-min share the chunks they read (consecutive poses of a trajectory see nearly the
// same part of the cloud), so a frame sequence streams the union of the two visible lists once per frame: the tile's
// records are loaded from the ring once and projected for both cameras.  Per tile (warp-uniform flags of the list
// entry, classify_pair_kernel): kTileZmin -> z-min for pp_zmin into zbuf_zmin; kTileBlend -> blend for pp_blend against
// zbuf_blend (complete: its z-min ran in the previous launch) into accum_blend.  The two frames use different frame
// sets, so nothing one half writes is read by the other.  Both gathers are issued before either is consumed.
// Registers: 2 CTAs x 512 threads x 64 registers are the SM's whole register file: nothing else runs beside the pass.
// Capping the kernel lower (RTR_FUSED_REGS = 56 / 48 would leave room for a CTA of the clear / classification / image
// kernels) was measured SLOWER: the spills land on the gathers' results and expose their latency (fused pass 88 -> 97 ->
// 120 us, frames/s 8 900 -> 8 180 -> 6 970; profiles/r02c_exp_fused_ab.json, ncu: 25 % of the stall samples on two STL).
#ifndef RTR_FUSED_REGS
#define RTR_FUSED_REGS 64
#endif
#ifndef RTR_FUSED_SEQ
#define RTR_FUSED_SEQ 0
#endif
#ifndef RTR_L2_HINTS
#define RTR_L2_HINTS 0
#endif
#if RTR_FUSED_REGS >= 64
#define RTR_FUSED_BOUNDS __launch_bounds__(kRingThreads, kRingMinCtas)
#else
#define RTR_FUSED_BOUNDS __maxnreg__(RTR_FUSED_REGS)
#endif
template <int ZV, int BV, bool DISTORT>
__global__ void RTR_FUSED_BOUNDS fused_ring_kernel(const PointRecord* __restrict__ pts, uint64_t n,
                                                                  const __grid_constant__ ProjParams pp_blend,
                                                                  const __grid_constant__ ProjParams pp_zmin,
                                                                  const __grid_constant__ RingSchedule sc,
                                                                  const uint32_t* __restrict__ zbuf_blend,
                                                                  unsigned long long* __restrict__ accum_blend,
                                                                  uint32_t* __restrict__ zbuf_zmin,
                                                                  const __grid_constant__ ClearTarget clr) {
    RingSmem& sm = ring_setup();
    if (sc.tiles_hint && blockIdx.x == 0 && threadIdx.x == 0) {  // (the list is older than this grid's PDL wait only in flush passes; a stale hint is harmless)
        asm volatile("griddepcontrol.wait;" ::: "memory");
        *reinterpret_cast<volatile uint32_t*>(sc.tiles_hint) = __ldcg(sc.cull->n_visible + (__ldcg(&sc.cull->parity) & 1u));
    }
    ring_walk<true>(pts, n, sc, sm, false, [&](const PointRecord (&p)[kRingPerThread], uint32_t entry, uint32_t first, uint32_t rot, uint32_t valid) {
#if RTR_FUSED_SEQ
        // (experiment build: one half after the other — fewer live registers, two exposed gather latencies per tile)
        if (entry & kTileZmin) {
            ZminLanes<0> z;
            zmin_project<ZV, DISTORT, 0>(z, p, pp_zmin, 0u, first, rot, valid, 0ull);
            zmin_gather<ZV, 0>(z, zbuf_zmin, nullptr);
            zmin_commit<ZV, 0>(z, zbuf_zmin, nullptr);
        }
        if (entry & kTileBlend) {
            BlendLanes b;
            blend_project<DISTORT>(b, p, pp_blend, first, rot, valid);
            blend_gather(b, zbuf_blend);
            blend_commit<BV>(b, p, accum_blend);
        }
#else
        ZminLanes<0> z;
        BlendLanes b;
        z.any = false;
        b.any = false;
        if (entry & kTileZmin) {
            zmin_project<ZV, DISTORT, 0>(z, p, pp_zmin, 0u, first, rot, valid, 0ull);
            zmin_gather<ZV, 0>(z, zbuf_zmin, nullptr);
        }
        if (entry & kTileBlend) {
            blend_project<DISTORT>(b, p, pp_blend, first, rot, valid);
            blend_gather(b, zbuf_blend);
        }
        zmin_commit<ZV, 0>(z, zbuf_zmin, nullptr);
        blend_commit<BV>(b, p, accum_blend);
#endif
    });
    // Out of tiles: clear this CTA's slice of the frame set the frame after next will use (fillBuffer + cudaMemset of
    // the reference, render.cu:16-31, project_cloud.cu:316-317).  Nobody reads that set any more (the launch waited for
    // its last frame's image passes and D2H), the stores are fire-and-forget, and the groups that finish first do
    // them while the others still stream: no clear kernel, no launch on anybody's critical path.
    if (clr.accum) {
        const uint64_t t = uint64_t(blockIdx.x) * kRingThreads + threadIdx.x, stride = uint64_t(gridDim.x) * kRingThreads;
#if RTR_L2_HINTS & 4
        for (uint64_t i = t; i < clr.n_px; i += stride) __stcs(clr.accum + i, make_uint4(0u, 0u, 0u, 0u));
#else
        for (uint64_t i = t; i < clr.n_px; i += stride) clr.accum[i] = make_uint4(0u, 0u, 0u, 0u);
#endif
        const uint64_t cov4 = clr.cov >> 2;
        uint4* z4 = reinterpret_cast<uint4*>(clr.zbuf);
#if RTR_L2_HINTS & 8
        for (uint64_t i = t; i < cov4; i += stride) __stcs(z4 + i, make_uint4(kEmptyDepthBits, kEmptyDepthBits, kEmptyDepthBits, kEmptyDepthBits));
#else
        for (uint64_t i = t; i < cov4; i += stride) z4[i] = make_uint4(kEmptyDepthBits, kEmptyDepthBits, kEmptyDepthBits, kEmptyDepthBits);
#endif
        for (uint64_t i = (cov4 << 2) + t; i < clr.cov; i += stride) clr.zbuf[i] = kEmptyDepthBits;
        if (t == 0) { clr.minmax[0] = 0xFFFFFFFFu; clr.minmax[1] = 0u; clr.minmax[2] = 0u; clr.minmax[3] = 0u; }
    }
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
    sc.grid_override = 0;
    sc.tiles_hint = nullptr;
    sc.zero = 0;
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
#ifdef RTR_EXPERIMENTS
    if (const char* c = std::getenv("RTR_RING_CARVEOUT"))  // measurement: shared-memory carve-out in percent (the rest of the 256 KB is L1)
        if (e == cudaSuccess) e = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, std::atoi(c));
#endif
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

template <typename K>
static cudaError_t launch_smem_tile(K kernel, bool* done, unsigned grid, cudaStream_t s, const PointRecord* pts, uint64_t n, const ProjParams& pp,
                                    const RingSchedule& sc, uint32_t* zbuf, unsigned long long* stats) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (!(dev >= 0 && dev < 64 && done[dev])) {
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sizeof(RingSmemTile)));
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) done[dev] = true;
    }
    launch_pdl_smem(kernel, dim3(grid), dim3(kRingThreads), sizeof(RingSmemTile), s, pts, n, pp, sc, zbuf, stats);
    return cudaGetLastError();
}

cudaError_t launch_zmin_ring(cudaStream_t s, int sm_count, int variant, const PointRecord* pts, uint64_t n,
                             uint64_t index_base, const ProjParams& pp, const RingSchedule& sc_in, bool list, uint32_t* zbuf,
                             unsigned long long* zkey) {
    if (n == 0) return cudaSuccess;
    const unsigned grid = ring_grid(sm_count, sc_in, list);
    RingSchedule sc = sc_in;
    sc.n_queues = ring_effective_queues(grid, kRingGroups, sc.n_queues);
    if ((variant & 64) && !zkey && !pp.distort) {  // shared-memory tile pre-reduction (pinhole, 32-bit z-buffer)
        static bool done_l[64] = {false}, done_a[64] = {false};
        unsigned long long* stats = sc.cull ? smem_tile_stats(const_cast<CullState*>(sc.cull)) : nullptr;
        if (list) return launch_smem_tile(zmin_ring_smem_kernel<5, true>, done_l, grid, s, pts, n, pp, sc, zbuf, stats);
        return launch_smem_tile(zmin_ring_smem_kernel<5, false>, done_a, grid, s, pts, n, pp, sc, zbuf, stats);
    }
    switch (variant & 45) {  // bit 1 (warp aggregation) has no ring form: the in-register merge replaces it
        case 0: return launch_zmin_ring_v<0>(s, grid, pts, n, index_base, pp, sc, list, zbuf, zkey);
        case 1: return launch_zmin_ring_v<1>(s, grid, pts, n, index_base, pp, sc, list, zbuf, zkey);
        case 5: return launch_zmin_ring_v<5>(s, grid, pts, n, index_base, pp, sc, list, zbuf, zkey);
#ifdef RTR_EXPERIMENTS  // measurement-only kernels (no RED / no merge: wrong frames) exist in experiment builds only
        case 37: return launch_zmin_ring_v<37>(s, grid, pts, n, index_base, pp, sc, list, zbuf, zkey);
        case 8: return launch_zmin_ring_v<8>(s, grid, pts, n, index_base, pp, sc, list, zbuf, zkey);
        case 9: return launch_zmin_ring_v<9>(s, grid, pts, n, index_base, pp, sc, list, zbuf, zkey);
        case 13: return launch_zmin_ring_v<13>(s, grid, pts, n, index_base, pp, sc, list, zbuf, zkey);
#endif
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
#ifdef RTR_EXPERIMENTS
    if (list && !pp.distort && f32 && (variant & 32)) {  // measurement: no in-register merge
        RTR_RING_LAUNCH((blend_ring_kernel<36, false, true>), pts, n, pp, sc, zbuf, a2);
        return cudaGetLastError();
    }
#endif
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


cudaError_t launch_fused_ring(cudaStream_t s, int sm_count, int zmin_variant, int blend_variant, const PointRecord* pts,
                              uint64_t n, const ProjParams& pp_blend, const ProjParams& pp_zmin, const RingSchedule& sc_in,
                              const uint32_t* zbuf_blend, uint32_t* accum_blend, uint32_t* zbuf_zmin, const ClearTarget& clear) {
    if (n == 0) return cudaSuccess;
    unsigned grid = ring_grid(sm_count, sc_in, true);
    RingSchedule sc = sc_in;
    if (sc.grid_override > grid) {  // short-lived CTAs, tiles dealt round-robin (no claims)
        grid = sc.grid_override;
        sc.tile_counter = nullptr;
    }
    sc.n_queues = ring_effective_queues(grid, kRingGroups, sc.n_queues);
    unsigned long long* a2 = reinterpret_cast<unsigned long long*>(accum_blend);
    const bool f32 = (blend_variant & 4) != 0, distort = pp_zmin.distort != 0;
    // z-min variants 0 / 1 / 5 (early test off / through L2 / through L1); everything else maps to the default 5
    const int zv = (zmin_variant & 5) == 0 ? 0 : ((zmin_variant & 5) == 1 ? 1 : 5);
#define RTR_FUSED(ZV)                                                                                                      \
    do {                                                                                                                   \
        if (distort) {                                                                                                     \
            if (f32) RTR_RING_LAUNCH((fused_ring_kernel<ZV, 4, true>), pts, n, pp_blend, pp_zmin, sc, zbuf_blend, a2, zbuf_zmin, clear);  \
            else RTR_RING_LAUNCH((fused_ring_kernel<ZV, 0, true>), pts, n, pp_blend, pp_zmin, sc, zbuf_blend, a2, zbuf_zmin, clear);      \
        } else {                                                                                                           \
            if (f32) RTR_RING_LAUNCH((fused_ring_kernel<ZV, 4, false>), pts, n, pp_blend, pp_zmin, sc, zbuf_blend, a2, zbuf_zmin, clear); \
            else RTR_RING_LAUNCH((fused_ring_kernel<ZV, 0, false>), pts, n, pp_blend, pp_zmin, sc, zbuf_blend, a2, zbuf_zmin, clear);     \
        }                                                                                                                  \
    } while (0)
    if (zv == 0) RTR_FUSED(0);
    else if (zv == 1) RTR_FUSED(1);
    else RTR_FUSED(5);
#undef RTR_FUSED
    return cudaGetLastError();
}

}  // namespace rtr
