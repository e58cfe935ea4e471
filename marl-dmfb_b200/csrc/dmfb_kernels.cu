// dmfb_kernels.cu — batched DMFB environment step for sm_100a (B200).
//
// Replaces, for N independent chips at once, the reference call tree
//   DMFBenv.step (env/DMFB/dmfb.py:560-587) -> RoutingTaskManager.moveDroplets (:253-299)
//   -> moveOneDroplet (:325-359) -> addUsage (:459-463) -> getObs (:622-626) -> getOneObs (:395-457)
// plus DMFBenv.reset (:589-597), getglobalobs (:368-392) and restart (:599-605).
//
// Design:
//  * one CTA per tile of E consecutive envs; E is chosen so that the tile's observation span
//    E*A*(3*fov^2+2) bytes is a multiple of 16 -> the CTA's output is one contiguous, 16-byte aligned
//    range of the [N,A,D] int8 tensor even though a single 245-byte row is not;
//  * one THREAD PER DROPLET, G = next power of two >= A lanes per env: the sequential, order dependent
//    move/revert loop of the reference runs as A rounds of shuffle + ballot inside the lane group; the
//    pairwise fluidic-constraint counts, rewards, dones and the usage update are per lane, per-env
//    scalars are handled by the group leader; droplet state lives in registers from load to store;
//  * the observation tile is staged in shared memory: 16-byte zero fill, boundary layer expanded from
//    bit masks (4 output bytes per multiply), sparse byte scatter for droplet ids / clipped goals /
//    direction bytes, then ONE TMA bulk store (cp.async.bulk.global.shared::cta) per tile;
//  * HBM-bound integer/byte work: no tensor cores.
#include <cstddef>

#include "common.cuh"

namespace dmfb {

thread_local char g_last_error[256] = "";
std::atomic<uint64_t> g_launches{0};

namespace {

constexpr int kMaxThreads = 512;
constexpr unsigned kFull = 0xFFFFFFFFu;

// Shared-memory carve-up of one tile, computed identically on host and device.
struct TileLayout {
    int E, A, D, nw, ncodes;
    uint32_t tile_bytes, off_l2row, off_l2col, off_flag, off_loglen, off_dirx, off_diry, off_sink, off_sets, total;
    __host__ __device__ TileLayout(int E_, int A_, int fov, int W, int Lc, int D_) {
        E = E_; A = A_; D = D_; nw = (fov * fov + 31) / 32; ncodes = 2 * (fov / 2) + 1;
        tile_bytes = ((uint32_t)(E * A * D) + 15u) & ~15u;
        uint32_t o = tile_bytes;
        off_l2row = o; o += (uint32_t)(ncodes * nw) * 4u;
        off_l2col = o; o += (uint32_t)(ncodes * nw) * 4u;
        off_flag = o; o += ((uint32_t)E + 3u) & ~3u;
        off_loglen = o; o += (uint32_t)E * 4u;
        off_dirx = o; o += ((uint32_t)(2 * W) + 3u) & ~3u;
        off_diry = o; o += ((uint32_t)(2 * Lc) + 3u) & ~3u;
        off_sink = o; o += 16u;
        off_sets = o; o += (uint32_t)E * 4u * (((uint32_t)A + 3u) & ~3u);   // CoordSets: 16-byte aligned per env
        total = (o + 15u) & ~15u;
    }
    __host__ __device__ TileLayout(const dmfb_cfg_t& c, int E_)
        : TileLayout(E_, c.n_agents, c.fov, c.width, c.length, c.obs_dim) {}
};

// S.flag values
constexpr uint8_t kFlagSelected = 1;  // masked reset: env selected (row must be stored)
constexpr uint8_t kFlagNewTask = 4;   // a new task was generated for the env: updateHealth still to run

struct TileSmem {
    int8_t* tile;
    uint32_t* l2row;  // [ncodes][nw]
    uint32_t* l2col;
    uint8_t* flag;    // [E]
    int32_t* loglen;  // [E] usage-log entries of the envs flagged kFlagNewTask (fused auto-reset)
    int8_t* dirx;
    int8_t* diry;
    int8_t* sink;     // predicated-off byte stores go here (keeps the paint code branch-free)
    uint8_t* sets;    // [E][4][slots] coordinate sets of the specialised instances (CoordSets)
    __device__ TileSmem(unsigned char* base, const TileLayout& L) {
        tile = reinterpret_cast<int8_t*>(base);
        l2row = reinterpret_cast<uint32_t*>(base + L.off_l2row);
        l2col = reinterpret_cast<uint32_t*>(base + L.off_l2col);
        flag = reinterpret_cast<uint8_t*>(base + L.off_flag);
        loglen = reinterpret_cast<int32_t*>(base + L.off_loglen);
        dirx = reinterpret_cast<int8_t*>(base + L.off_dirx);
        diry = reinterpret_cast<int8_t*>(base + L.off_diry);
        sink = reinterpret_cast<int8_t*>(base + L.off_sink);
        sets = base + L.off_sets;
    }
};

static_assert(offsetof(dmfb_cfg_t, dir_x) % 4 == 0 && offsetof(dmfb_cfg_t, dir_y) % 4 == 0, "word-wise table loads");

// tid / nthreads: the (sub)set of CTA threads that cooperates on the fill
__device__ __forceinline__ void load_tables(const dmfb_cfg_t& cfg, const TileLayout& L, const TileSmem& S, int tid,
                                            int nthreads)
{
    const int nt = L.ncodes * L.nw;
#pragma unroll 1
    for (int k = tid; k < nt; k += nthreads) {
        const int c = k / L.nw, j = k - c * L.nw;
        S.l2row[k] = cfg.l2_row[c][j];
        S.l2col[k] = cfg.l2_col[c][j];
    }
    // 4 table bytes per (lane-divergent, hence serialised) constant-bank load; both tables are 4-byte aligned
    const uint32_t* dx4 = reinterpret_cast<const uint32_t*>(cfg.dir_x);
    const uint32_t* dy4 = reinterpret_cast<const uint32_t*>(cfg.dir_y);
#pragma unroll 1
    for (int k = tid; k < (2 * cfg.width + 3) / 4; k += nthreads) reinterpret_cast<uint32_t*>(S.dirx)[k] = dx4[k];
#pragma unroll 1
    for (int k = tid; k < (2 * cfg.length + 3) / 4; k += nthreads) reinterpret_cast<uint32_t*>(S.diry)[k] = dy4[k];
}

__device__ __forceinline__ void zero_tile(const TileLayout& L, const TileSmem& S, int tid, int nthreads)
{
    uint4* t4 = reinterpret_cast<uint4*>(S.tile);
    const int n16 = (int)(L.tile_bytes >> 4);
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    int k = tid;
    const int stride = nthreads;
    for (; k + 3 * stride < n16; k += 4 * stride) {
        t4[k] = z; t4[k + stride] = z; t4[k + 2 * stride] = z; t4[k + 3 * stride] = z;
    }
    for (; k < n16; k += stride) t4[k] = z;
}

// Compile-time sized zero fill: THREADS threads, BYTES % 16 == 0; fully unrolled STS.128 with immediate offsets.
template <int BYTES, int THREADS>
__device__ __forceinline__ void zero_tile_static(int8_t* tile, int tid)
{
    constexpr int n16 = BYTES / 16;
    uint4* t4 = reinterpret_cast<uint4*>(tile) + tid;
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (int k = 0; k < (n16 + THREADS - 1) / THREADS; ++k)
        if (k * THREADS + THREADS <= n16 || tid + k * THREADS < n16) t4[k * THREADS] = z;
}

// Two "within one cell" tests in one VABSDIFF4: a and b each pack two cells (x0 | y0<<8 | x1<<16 | y1<<24);
// bit 0 of the result: cells 0 are neighbours (or equal), bit 1: cells 1 are.  |dx| <= 1 && |dy| <= 1 is, on the
// integer grid, both "Euclid < 2" (dmfb.py:258,268) and "squared distance <= 2" (dmfb.py:220).
__device__ __forceinline__ uint32_t near_pair(uint32_t a, uint32_t b)
{
    const uint32_t t = __vabsdiffu4(a, b) & 0xFEFEFEFEu;   // a byte is zero iff that |difference| <= 1
    return (uint32_t)((t & 0xFFFFu) == 0u) | ((uint32_t)((t >> 16) == 0u) << 1);
}
__device__ __forceinline__ uint32_t dup_lo(uint32_t w) { return __byte_perm(w, 0, 0x1010); }  // cell0 | cell0<<16
__device__ __forceinline__ uint32_t dup_hi(uint32_t w) { return __byte_perm(w, 0, 0x3232); }  // cell1 | cell1<<16

// Coordinate sets of one env in shared memory (instances with a compile-time droplet count): the x and the y
// coordinates of all droplets as byte arrays, padded to whole words - CX, CY after the moves, PX, PY before.  A lane
// then tests its droplet against FOUR others per VABSDIFF4 (the byte-replicated own coordinate against a word of the
// set) instead of one per shuffle: the pairwise constraint counts and the visibility test of the observation drop from
// ~17 and 4 instructions per PAIR to ~9 and 3 per WORD.  Pad slots hold 128 + fov/2: further than any window reaches
// and more than one cell from every real coordinate (< 128), yet small enough that "difference + bias" stays a byte.
template <int A_T>
struct CoordSets {
    static constexpr int kSlots = (A_T + 3) & ~3, kWords = kSlots / 4, kBytes = 4 * kSlots;
    uint8_t* p;                                                  // the env's block: [CX | CY | PX | PY], kSlots bytes each
    uint32_t x[kWords], y[kWords];                               // CX / CY once loaded
    __device__ CoordSets(uint8_t* sets, int env_in_tile) : p(sets + env_in_tile * kBytes) {}
    // lane i publishes its coordinates in the sets sx / sy (0: CX, 1: CY, 2: PX, 3: PY); lanes i < kSlots - A_T also pad
    __device__ __forceinline__ void put(int sx, int sy, int i, uint32_t vx, uint32_t vy, uint32_t pad_byte) const {
        p[sx * kSlots + i] = (uint8_t)vx;
        p[sy * kSlots + i] = (uint8_t)vy;
        if (A_T + i < kSlots) {
            p[sx * kSlots + A_T + i] = (uint8_t)pad_byte;
            p[sy * kSlots + A_T + i] = (uint8_t)pad_byte;
        }
    }
    __device__ __forceinline__ uint32_t word(int s, int w) const { return reinterpret_cast<const uint32_t*>(p)[s * kWords + w]; }
    __device__ __forceinline__ void load_current() {
#pragma unroll
        for (int w = 0; w < kWords; ++w) { x[w] = word(0, w); y[w] = word(1, w); }
    }
    // number of slots within one cell (|dx| <= 1 and |dy| <= 1) of (vx, vy) among the words (sx, sy); pads never are
    __device__ __forceinline__ static int count_near(uint32_t vx, uint32_t vy, const uint32_t* sx, const uint32_t* sy) {
        const uint32_t mx = vx * 0x01010101u, my = vy * 0x01010101u;
        int far = 0;
#pragma unroll
        for (int w = 0; w < kWords; ++w) {
            const uint32_t t = __vabsdiffu4(mx, sx[w]) | __vabsdiffu4(my, sy[w]);
            const uint32_t h = (t >> 1) & 0x7F7F7F7Fu;            // a byte is zero iff both differences are <= 1
            far += __popc((h + 0x7F7F7F7Fu) & 0x80808080u);       // bit 7 of a byte: nonzero
        }
        return kSlots - far;
    }
    // bit j set iff droplet j (current sets) lies in the window of half-width hf around (vx, vy):
    // 2|dx| < fov and 2|dy| < fov for odd fov
    __device__ __forceinline__ uint32_t visible(uint32_t vx, uint32_t vy, int hf) const {
        const uint32_t mx = vx * 0x01010101u, my = vy * 0x01010101u;
        const uint32_t bias = (uint32_t)(0x7F - hf) * 0x01010101u; // byte + bias sets bit 7 iff byte > hf (no carry: byte <= 128 + hf)
        uint32_t hidden = 0;
#pragma unroll
        for (int w = 0; w < kWords; ++w) {
            const uint32_t nv = ((__vabsdiffu4(mx, x[w]) + bias) | (__vabsdiffu4(my, y[w]) + bias)) & 0x80808080u;
            hidden |= (((nv >> 7) * 0x01020408u) >> 24) << (4 * w);   // bits 0, 8, 16, 24 -> one nibble
        }
        return ~hidden & ((1u << A_T) - 1u);
    }
};
template <>
struct CoordSets<0> {                                            // run-time droplet count: shuffles, no sets
    __device__ CoordSets(uint8_t*, int) {}
};

// A lane group = the G consecutive lanes of a warp that hold the droplets of one env.  G is a power of two
// (4, 8, 16, 32) or, to avoid idle lanes, exactly the droplet count (G = 10: three envs per warp, lanes 30-31 idle).
template <int G>
struct Group {
    static constexpr bool kPow2 = (G & (G - 1)) == 0;
    static constexpr int kPerWarp = 32 / G;                      // envs per warp
    static constexpr unsigned kBits = (G == 32) ? 0xFFFFFFFFu : ((1u << G) - 1u);
    int lane, base, i, idx;
    bool valid;                                                  // false for the spare lanes of a non-power-of-two G
    __device__ explicit Group(int tid) {
        lane = tid & 31;
        idx = lane / G;                                          // group index inside the warp
        valid = idx < kPerWarp;
        i = lane - idx * G;
        base = valid ? idx * G : 0;
    }
    // env (inside the tile) held by this lane's group
    __device__ __forceinline__ int env_in_tile(int tid) const { return (tid >> 5) * kPerWarp + idx; }
    // value of lane j of my group (all 32 lanes must call)
    template <typename T>
    __device__ __forceinline__ T get(T v, int j) const {
        if constexpr (kPow2) return __shfl_sync(kFull, v, j, G);
        else return __shfl_sync(kFull, v, base + j);
    }
    // bits of my group from a warp ballot
    __device__ __forceinline__ unsigned ballot(bool p) const { return (__ballot_sync(kFull, p) >> base) & kBits; }
    // group sum (all 32 lanes must call).  (One REDUX per integer sum and doubling windows for the float sum - 4
    // shuffles instead of 10 for G = 10 - were measured and dropped: C2 33.2 -> 34.6 us.)
    template <typename T>
    __device__ __forceinline__ T sum(T v) const {
        if constexpr (kPow2) {
#pragma unroll
            for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
            return v;
        } else {
            T s = 0;
#pragma unroll
            for (int j = 0; j < G; ++j) s += __shfl_sync(kFull, v, base + j);
            return s;
        }
    }
};

// threads per CTA for E envs with G lanes per env
__host__ __device__ constexpr int cta_threads(int E, int G) { return ((E + 32 / G - 1) / (32 / G)) * 32; }

// ---- _Generate_Start_End (dmfb.py:207-226) -------------------------------------------------------------------------
// 2A uniform cells; the whole set is redrawn until every pairwise squared distance is > 2.  Attempt number k for
// (seed, env, episode) is a pure function of those values, and the task is the FIRST accepted attempt - the same
// accept/reject rule on the same proposal distribution as the reference, so tasks are distributed exactly like the
// reference's, independent of sharding (cfg.env_base) and of who examines which attempt when.
//
// The search is organised in ROUNDS: the whole warp examines kAttemptsPerRound consecutive attempts of ONE env and
// hands the lowest accepted one to the lanes of the group `dst` that holds that env (lane i gets the word of droplet
// i).  Two flavours, chosen by A alone (so every kernel instance draws the same task for the same key):
//   A == 10  one attempt per WARP and round: lane = point, the 190 pair tests by shuffles with early exit;
//   else     one attempt per lane GROUP, points exchanged by shuffles (32/G per round).
constexpr uint32_t kTaskReady = 0x80000000u;
constexpr uint32_t kStepNoRunAhead = 0x80000000u;   // internal step flag: the run-ahead search is a kernel of its own   // dmfb_state_t.next_cursor: next_task holds the next episode's task

template <int G>
__device__ __forceinline__ int attempts_per_round(int A) { return A == 10 ? 1 : Group<G>::kPerWarp; }
// rounds of the run-ahead search per warp and step (fused auto-reset): enough to keep up with the demand - three envs
// per warp that each need about 70 attempts (10 droplets, 20x20) per episode of 80 steps = 2.6 per step
__device__ __forceinline__ int prefetch_rounds(int A) { return A == 10 ? 4 : 1; }

// One round for env `env` (global index), episode `epi`, attempts [first, first + attempts_per_round).  All 32 lanes
// call with warp-uniform arguments.  Returns true if an attempt was accepted; then `word` of the lanes of group `dst`
// (lane i < A) is the task, other lanes keep theirs.
//
// A == 10: ONE attempt per round, spread over the warp.  Lane p < 20 draws point p (point 2j = start of droplet j,
// 2j+1 = its goal) and tests it against the points p+1 .. p+10 (mod 20) - all 190 pairs - stopping at the first
// offset at which any lane saw a conflict (an attempt on a 20x20 chip fails with 98.6 %, after 2.5 offsets on average).
// A round is ~60 instructions: what a step that runs ahead adds is small and the same for every warp.  (One attempt
// per LANE with all points in registers, 32 per round, costs the same per attempt but ~3,000 instructions per round:
// the few warps that ran a round in a step kept the whole launch waiting, +11 us at 64K envs of 10 droplets.)
// Stream of the attempts of (seed, env, episode) for the one-attempt-per-warp flavour
__device__ __forceinline__ uint64_t layout_stream(uint64_t seed, int64_t env, uint32_t epi)
{
    uint64_t base = seed ^ (0x9E3779B97F4A7C15ull * (uint64_t)(kStreamLayout + 1));
    base += (uint64_t)env * 0xD1342543DE82EF95ull + ((uint64_t)epi << 32) * 0xDA942042E4DD58B5ull;
    return mix64(base ^ 0xA5A5A5A5A5A5A5A5ull);
}

// Attempts first, first + 1, ... (at most `max_rounds`) of the stream `base`, one per iteration; stops at the first
// accepted one.  Returns the number of attempts examined; `hit` says whether the last one was accepted.
template <int G, int A_T>
__device__ __forceinline__ uint32_t sample_rounds_warp(const dmfb_cfg_t& cfg, const Group<G>& g, uint64_t base,
                                                       uint32_t first, uint32_t max_rounds, int dst, uint32_t& word, bool& hit)
{
    constexpr int P = 2 * A_T;                                    // points of one attempt, P <= 32
    const uint32_t W = (uint32_t)cfg.width, Lc = (uint32_t)cfg.length;
    const int p = g.lane;
    const bool on = p < P;
    hit = false;
    uint32_t done = 0;
#pragma unroll 1
    while (done < max_rounds) {
        const uint64_t ctr = base + (uint64_t)(first + done) * (uint64_t)P * 0x9E3779B97F4A7C15ull;
        ++done;
        const uint64_t z = mix64(ctr + (uint64_t)(p + 1) * 0x9E3779B97F4A7C15ull);
        // idle lanes hold distinct far-away cells
        const uint32_t cell = on ? (__umulhi((uint32_t)z, W) | (__umulhi((uint32_t)(z >> 32), Lc) << 8)) : (0xFF00u | (uint32_t)p);
        bool bad = false;
#pragma unroll 1
        for (int o = 1; o <= P / 2 && !bad; ++o) {
            int q = p + o;
            if (q >= P) q -= P;
            const uint32_t other = __shfl_sync(kFull, cell, on ? q : p);
            bad = __any_sync(kFull, on && (__vabsdiffu4(cell, other) & 0xFEFEu) == 0u);   // |dx| <= 1 && |dy| <= 1
#ifdef DMFB_WHATIF_SEARCH_FREE
            bad = false;
            break;
#endif
        }
        if (bad) continue;
        // accepted: droplet i = start (point 2i) | goal (point 2i+1) << 16
        const int i2 = (g.i < A_T) ? 2 * g.i : 0;
        const uint32_t w = __shfl_sync(kFull, cell, i2) | (__shfl_sync(kFull, cell, i2 + 1) << 16);
        if (g.idx == dst && g.i < A_T) word = w;
        hit = true;
        break;
    }
    return done;
}

template <int G>
__device__ __forceinline__ bool sample_round(const dmfb_cfg_t& cfg, const Group<G>& g, int A, uint64_t seed, int64_t env,
                                             uint32_t epi, uint32_t first, int dst, uint32_t& word)
{
    if (A == 10) {
        bool hit;
        sample_rounds_warp<G, 10>(cfg, g, layout_stream(seed, env, epi), first, 1u, dst, word, hit);
        return hit;
    }
    const uint32_t W = (uint32_t)cfg.width, Lc = (uint32_t)cfg.length;
    const bool lane_in = g.valid && g.i < A;
    // per-(env, episode, droplet) stream; attempt k uses counters 2k+1, 2k+2
    uint64_t base = seed ^ (0x9E3779B97F4A7C15ull * (uint64_t)(kStreamLayout + 1));
    base += (uint64_t)env * 0xD1342543DE82EF95ull + ((uint64_t)epi << 32 | (uint32_t)g.i) * 0xDA942042E4DD58B5ull;
    base = mix64(base);
    const uint64_t k = (uint64_t)first + (uint64_t)g.idx;
    uint32_t w = 0;
    if (lane_in) {
        const uint64_t z0 = mix64(base + (2 * k + 1) * 0x9E3779B97F4A7C15ull);
        const uint64_t z1 = mix64(base + (2 * k + 2) * 0x9E3779B97F4A7C15ull);
        w = __umulhi((uint32_t)z0, W) | (__umulhi((uint32_t)(z0 >> 32), Lc) << 8) |
            (__umulhi((uint32_t)z1, W) << 16) | (__umulhi((uint32_t)(z1 >> 32), Lc) << 24);
    }
    uint32_t bad = near_pair(w, w >> 16) & 1u;                // own start vs own goal
    for (int j = 0; j < A; ++j) {
        const uint32_t o = g.get(w, j);
        const uint32_t hit = near_pair(w, dup_lo(o)) | near_pair(w, dup_hi(o));  // my 2 points vs theirs
        if (j != g.i) bad |= hit;
    }
    const unsigned gb = g.ballot(bad != 0u && lane_in);
    const unsigned okm = __ballot_sync(kFull, g.valid && gb == 0u && g.i == 0);   // leaders of accepting groups
    if (okm == 0u) return false;
    const int win = __ffs(okm) - 1;                           // lowest attempt number of this round
    const uint32_t wsel = __shfl_sync(kFull, w, win + g.i);
    if (g.idx == dst) word = wsel;
    return true;
}

// Complete searches for the groups whose leader has `want` set, each starting at attempt `first` (uniform per group;
// 0 or the cursor of a search that already ran ahead).  Every lane of the warp must call; env0 = global env index of
// the warp's group 0.  A density that cannot be placed makes the reference loop for ever (dmfb.py:212-224); here the
// search gives up after kMaxSamplerRounds rounds, raises DMFB_STATUS_SAMPLER_GAVE_UP and keeps the previous layout.
template <int G>
__device__ __forceinline__ uint32_t generate_layout(const dmfb_cfg_t& cfg, const dmfb_state_t& st, const Group<G>& g, int A,
                                                    uint64_t seed, int64_t env0, uint32_t episode, bool want,
                                                    uint32_t first, uint32_t keep)
{
    uint32_t word = keep;
    unsigned todo = __ballot_sync(kFull, want && g.i == 0);      // leader lanes of the requesting groups
    if (todo == 0u) return word;
    const uint32_t per_round = (uint32_t)attempts_per_round<G>(A);
    while (todo) {
        const int src = __ffs(todo) - 1;                          // leader lane of the group served now
        todo &= todo - 1;
        const uint32_t epi = __shfl_sync(kFull, episode, src);
        uint32_t at = __shfl_sync(kFull, first, src);
        if (A == 10) {
            bool hit;
            sample_rounds_warp<G, 10>(cfg, g, layout_stream(seed, env0 + src / G, epi), at, kMaxSamplerRounds, src / G, word, hit);
            if (!hit && st.gen_status && g.lane == 0) atomicOr(st.gen_status, DMFB_STATUS_SAMPLER_GAVE_UP);
            continue;
        }
        for (uint32_t round = 0;; ++round, at += per_round) {
            if (round >= kMaxSamplerRounds) {                     // density that cannot be placed
                if (st.gen_status && g.lane == 0) atomicOr(st.gen_status, DMFB_STATUS_SAMPLER_GAVE_UP);
                break;
            }
            if (sample_round<G>(cfg, g, A, seed, env0 + src / G, epi, at, src / G, word)) break;
        }
    }
    return word;
}

// GenRandomBlocks (dmfb.py:228-251) for the envs whose group has `want` set: every 2x2 block is drawn with x_min
// uniform in [0, W-4] and y_min uniform in [0, L-4] and redrawn while it covers a start / goal cell of the task
// `word` (each lane checks its own droplet) or overlaps an earlier block (checked by the leader lane, which also
// stores the result).  Every lane of the warp must call; `want` / `episode` are uniform per group.
template <int G>
__device__ __forceinline__ void generate_blocks(const dmfb_cfg_t& cfg, const dmfb_state_t& st, const Group<G>& g,
                                                uint64_t seed, int64_t n, uint32_t episode, bool want, bool lane_on,
                                                uint32_t word)
{
    const int nb = cfg.n_blocks;
    if (nb == 0 || !__any_sync(kFull, want)) return;
    uint8_t* blocks = st.blocks + (size_t)n * nb * 2;     // only dereferenced by leaders of groups with want
    uint64_t state = seed ^ (0x9E3779B97F4A7C15ull * (uint64_t)(kStreamBlocks + 1));
    state += (uint64_t)(cfg.env_base + n) * 0xD1342543DE82EF95ull + ((uint64_t)episode << 32) * 0xDA942042E4DD58B5ull;
    state = mix64(state);
    uint8_t bx[DMFB_MAX_BLOCKS], by[DMFB_MAX_BLOCKS];
    const int sx = word & 255u, sy = (word >> 8) & 255u, tx = (word >> 16) & 255u, ty = word >> 24;
    for (int b = 0; b < nb; ++b) {
        bool pending = want;
        uint32_t rounds = 0;
        while (__any_sync(kFull, pending)) {
            if (++rounds >= kMaxSamplerRounds) {                  // obstacles that cannot be placed: previous ones stay
                if (st.gen_status && g.lane == 0) atomicOr(st.gen_status, DMFB_STATUS_SAMPLER_GAVE_UP);
                return;
            }
            uint32_t cand = 0;
            if (pending && g.i == 0) {
                const uint64_t z = mix64(state += 0x9E3779B97F4A7C15ull);
                cand = __umulhi((uint32_t)z, (uint32_t)(cfg.width - 3)) |
                       (__umulhi((uint32_t)(z >> 32), (uint32_t)(cfg.length - 3)) << 8);
            }
            cand = g.get(cand, 0);
            const int x = cand & 255u, y = cand >> 8;
            const bool covers = lane_on && (((unsigned)(sx - x) <= 1u && (unsigned)(sy - y) <= 1u) ||
                                            ((unsigned)(tx - x) <= 1u && (unsigned)(ty - y) <= 1u));
            const unsigned any_cover = g.ballot(covers);
            int ok = 0;
            if (pending && g.i == 0 && any_cover == 0u) {
                ok = 1;
                for (int k = 0; k < b; ++k)   // isBlockOverlap (:56-69): inclusive ranges intersect on both axes
                    if (!(x > bx[k] + 1 || bx[k] > x + 1) && !(y > by[k] + 1 || by[k] > y + 1)) ok = 0;
                if (ok) {
                    bx[b] = (uint8_t)x; by[b] = (uint8_t)y;
                    blocks[2 * b] = (uint8_t)x; blocks[2 * b + 1] = (uint8_t)y;
                }
            }
            if (g.get(ok, 0)) pending = false;
        }
    }
}

// ---- observation painting: getOneObs (dmfb.py:395-457) of one agent into the zero filled tile --------
// get(j) returns the packed word of droplet j of the same env; every lane must call it (it shuffles).
// Only lanes with `on` store.  Byte ranges of different agents never overlap: layer 2 is written as whole
// 4-byte words only where the word lies inside this agent's record (its first word may cover the last
// <= 3 bytes of the agent's own layer 1, which is painted afterwards), the rest byte by byte.
template <int FOV_T, int A_T, typename GetWord>
__device__ __forceinline__ void paint_agent(const dmfb_cfg_t& cfg, const TileLayout& L, const TileSmem& S,
                                            int agent_in_tile, int i, uint32_t me, bool on, GetWord get,
                                            const CoordSets<A_T>& cs, const uint8_t* __restrict__ env_blocks = nullptr)
{
    const int fov = FOV_T ? FOV_T : cfg.fov;
    const int hf = fov >> 1, f2 = fov * fov;
    const int A = A_T ? A_T : L.A, D = FOV_T ? 3 * FOV_T * FOV_T + 2 : L.D;
    const int nw = FOV_T ? (FOV_T * FOV_T + 31) / 32 : L.nw;
    const int W = cfg.width, Lc = cfg.length;
    const int x = me & 255u, y = (me >> 8) & 255u, gx = (me >> 16) & 255u, gy = me >> 24;
    int8_t* rec = S.tile + agent_in_tile * D;

    // ---- layer 2: off-chip boundary (:428-439) expanded from bit masks, 4 output bytes per multiply ----
    if (on) {
        const int lb = hf - x, rb = hf - (W - 1 - x);
        const int ub = hf - y, db = hf - (Lc - 1 - y);
        const int rc = lb > 0 ? lb : (rb > 0 ? hf + rb : 0);
        const int cc = ub > 0 ? ub : (db > 0 ? hf + db : 0);
        if (rc | cc) {
            const uint32_t* rowm = S.l2row + rc * nw;
            const uint32_t* colm = S.l2col + cc * nw;
            const int base = agent_in_tile * D + 2 * f2;  // byte offset of the layer inside the tile
            const int s = base & 3;                       // misalignment of the layer start
            uint32_t* wptr = reinterpret_cast<uint32_t*>(S.tile + (base & ~3));
            const int nfull_min = f2 >> 2;                // words that are whole for every alignment
            uint32_t prev = 0;
#pragma unroll
            for (int j = 0; j <= nw; ++j) {
                if (8 * j >= nfull_min) break;
                const uint32_t m = (j < nw) ? (rowm[j] | colm[j]) : 0u;
                const uint32_t sh = __funnelshift_l(prev, m, s);  // mask bits shifted up by s
                prev = m;
                const uint32_t even = sh & 0x0F0F0F0Fu, odd = (sh >> 4) & 0x0F0F0F0Fu;  // nibbles 0,2,4,6 / 1,3,5,7
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    const int k = 8 * j + t;
                    // nibble t isolated with one PRMT, then 4 bits -> 4 bytes: bit b -> byte b
                    // (multiply by 1 + 2^7 + 2^14 + 2^21: 16 distinct partial products, no carries)
                    const uint32_t nib = __byte_perm((t & 1) ? odd : even, 0u, 0x4440u + (uint32_t)(t >> 1));
                    if (k < nfull_min) wptr[k] = (nib * 0x00204081u) & 0x01010101u;
                }
            }
            // the last (f2 & 3) + s <= 6 bytes: one more whole word if they reach 4, then single bytes
            const int q0 = 4 * nfull_min - s;             // first layer bit not yet written
            const int wi = q0 >> 5;
            const uint32_t lo = (wi < nw) ? (rowm[wi] | colm[wi]) : 0u;
            const uint32_t hi = (wi + 1 < nw) ? (rowm[wi + 1] | colm[wi + 1]) : 0u;
            uint32_t rem = __funnelshift_r(lo, hi, q0 & 31);
            int nrem = f2 - q0;
            int8_t* bp = reinterpret_cast<int8_t*>(wptr + nfull_min);
            if (nrem >= 4) {
                wptr[nfull_min] = ((rem & 0xFu) * 0x00204081u) & 0x01010101u;
                rem >>= 4; nrem -= 4; bp += 4;
            }
            *(nrem > 0 ? bp : S.sink) = (int8_t)(rem & 1u);
            *(nrem > 1 ? bp + 1 : S.sink) = (int8_t)((rem >> 1) & 1u);
            *(nrem > 2 ? bp + 2 : S.sink) = (int8_t)((rem >> 2) & 1u);
        }
    }
    // ---- layer 2, obstacles: the reference writes block cells at their ABSOLUTE chip coordinates (:422-426) ---
    if (A_T == 0 && env_blocks != nullptr && on) {
        for (int b = 0; b < cfg.n_blocks; ++b) {
            const int bx = env_blocks[2 * b], by = env_blocks[2 * b + 1];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int ci = bx + (q >> 1), cj = by + (q & 1);
                if (ci < fov && cj < fov) rec[2 * f2 + ci * fov + cj] = 1;
            }
        }
    }
    // ---- layer 0: droplets inside the window (:408-413); layer 1: clipped goals of the other droplets
    //      with |dx| < fov/2 and |dy| < fov/2 (:416-420) --------------------------------------------------
    const int ox = x - hf, oy = y - hf;
    if constexpr (A_T == 10) {
        // Ten droplets, few of them inside any one window: the visible ones come out of the coordinate sets as a bit
        // mask (three VABSDIFF4 pairs instead of ten per-droplet tests) and only those are painted, in ascending
        // order.  The trip count is the largest number of visible droplets over the warp (4-5 on a 20x20 chip).
        static_assert(FOV_T & 1, "the window test below is the odd-fov one");
        uint32_t m = on ? cs.visible((uint32_t)x, (uint32_t)y, hf) : 0u;
        while (__any_sync(kFull, m != 0u)) {
            const bool act = m != 0u;
            const int j = act ? __ffs(m) - 1 : 0;
            m &= m - 1u;
            const uint32_t d = get(j);
            const int rx = (int)(d & 255u) - ox, ry = (int)((d >> 8) & 255u) - oy;
            *(act ? rec + rx * fov + ry : S.sink) = (int8_t)(j + 1);
            int cx = (int)((d >> 16) & 255u) - ox, cy = (int)(d >> 24) - oy;
            cx = min(max(cx, 0), fov - 1);
            cy = min(max(cy, 0), fov - 1);
            *((act && j != i) ? rec + f2 + cx * fov + cy : S.sink) = (int8_t)(j + 1);      // ascending j: later index overwrites
        }
    } else {
    const uint32_t vis_bias = (uint32_t)(0x7F - ((fov - 1) >> 1)) * 0x0101u;  // byte + bias sets bit 7 iff byte > (fov-1)/2
#pragma unroll
    for (int j = 0; j < A; ++j) {
        const uint32_t d = get(j);
        const uint32_t ad = __vabsdiffu4(me, d) & 0xFFFFu;            // |dx| | |dy|<<8
        const bool vis = ((ad + vis_bias) & 0x8080u) == 0u;           // 2|dx| < fov && 2|dy| < fov
        const int rx = (int)(d & 255u) - ox, ry = (int)((d >> 8) & 255u) - oy;
        const bool in0 = (fov & 1) ? vis : ((unsigned)rx < (unsigned)fov && (unsigned)ry < (unsigned)fov);
        // predicated-off stores are redirected to a sink byte instead of branching around them
        *((on && in0) ? rec + rx * fov + ry : S.sink) = (int8_t)(j + 1);
        int cx = (int)((d >> 16) & 255u) - ox, cy = (int)(d >> 24) - oy;
        cx = min(max(cx, 0), fov - 1);
        cy = min(max(cy, 0), fov - 1);
        *((on && vis && j != i) ? rec + f2 + cx * fov + cy : S.sink) = (int8_t)(j + 1);  // ascending j: later index overwrites
    }
    }
    // ---- direction bytes (:442-454) from the host-built table -------------------------------------------
    const int dxi = on ? gx - x + W - 1 : 0, dyi = on ? gy - y + Lc - 1 : 0;
    *(on ? rec + 3 * f2 : S.sink) = S.dirx[dxi];
    *(on ? rec + 3 * f2 + 1 : S.sink) = S.diry[dyi];
}

// ---- DMFBenv_v0_1.getOneObs (dmfb.py:727-835) of one agent: generic path only, one thread per agent ----------
// Layers: 0 droplets in the window; 1 own goal (projected onto the window for fewer than 10 droplets, :751-761);
// 2 goals of the visible others, nearest-to-goal first, each drawn where the ray droplet -> goal leaves the
// window and pushed to a free 4-neighbour when the cell is taken (:764-808); 3 obstacles at absolute
// coordinates + border (:811-831); then the NUMERATORS of the direction ((tar_y-y)/length, (tar_x-x)/width) (:833).
template <typename GetWord>
__device__ __noinline__ void paint_agent_v01(const dmfb_cfg_t& cfg, const TileLayout& L, const TileSmem& S, int agent_in_tile,
                                             int i, uint32_t me, bool on, GetWord get,
                                             const uint8_t* __restrict__ env_blocks)
{
    const int fov = cfg.fov, hf = fov >> 1, f2 = fov * fov, A = L.A, W = cfg.width, Lc = cfg.length;
    uint32_t w[DMFB_MAX_AGENTS];
    for (int j = 0; j < A; ++j) w[j] = get(j);              // every lane takes part in the shuffles
    if (!on) return;
    const int x = me & 255u, y = (me >> 8) & 255u, gx = (me >> 16) & 255u, gy = me >> 24;
    const int ox = x - hf, oy = y - hf;
    int8_t* rec = S.tile + agent_in_tile * L.D;
    uint32_t seen = 0;                                      // other droplets inside the window
    for (int j = 0; j < A; ++j) {
        const int rx = (int)(w[j] & 255u) - ox, ry = (int)((w[j] >> 8) & 255u) - oy;
        if ((unsigned)rx < (unsigned)fov && (unsigned)ry < (unsigned)fov) {
            rec[rx * fov + ry] = (int8_t)(j + 1);
            if (j != i) seen |= 1u << j;
        }
    }
    {
        int rx = gx - ox, ry = gy - oy;
        const bool inside = (unsigned)rx < (unsigned)fov && (unsigned)ry < (unsigned)fov;
        if (A < 10) {
            rx = min(max(rx, 0), fov - 1);
            ry = min(max(ry, 0), fov - 1);
        }
        if (A < 10 || inside) rec[f2 + rx * fov + ry] = (int8_t)(i + 1);
    }
    int8_t* l2 = rec + 2 * f2;
    while (seen) {
        // list.sort(key=distance) is stable: smallest Manhattan distance to the goal first, ties by index
        int best = 0, bd = 0x7FFFFFFF;
        for (uint32_t m = seen; m; m &= m - 1u) {
            const int j = __ffs(m) - 1;
            const uint32_t d = w[j];
            const int dist = abs((int)(d & 255u) - (int)((d >> 16) & 255u)) + abs((int)((d >> 8) & 255u) - (int)(d >> 24));
            if (dist < bd) { bd = dist; best = j; }
        }
        seen &= ~(1u << best);
        const uint32_t d = w[best];
        const int sx = (int)(d & 255u) - ox, sy = (int)((d >> 8) & 255u) - oy;
        const int dx = (int)((d >> 16) & 255u) - (int)(d & 255u), dy = (int)(d >> 24) - (int)((d >> 8) & 255u);
        const int boundx = dx >= 0 ? fov - 1 - sx : -sx;
        const int boundy = dy >= 0 ? fov - 1 - sy : -sy;
        int cdx, cdy;
        if (abs(dx) <= abs(boundx) && abs(dy) <= abs(boundy)) { cdx = dx; cdy = dy; }
        else if (dx == 0) { cdx = 0; cdy = boundy; }
        else if (dy == 0) { cdx = boundx; cdy = 0; }
        else {
            // python: dx / dy * boundy == (dx / dy) * boundy in float64; dy * boundx / dx == (dy * boundx) / dx
            const double qx = __dmul_rn(__ddiv_rn((double)dx, (double)dy), (double)boundy);
            const double qy = __ddiv_rn((double)(dy * boundx), (double)dx);
            cdx = dx >= 0 ? min(boundx, (int)ceil(qx)) : max(boundx, (int)floor(qx));
            cdy = dy >= 0 ? min(boundy, (int)ceil(qy)) : max(boundy, (int)floor(qy));
        }
        const int ci = sx + cdx, cj = sy + cdy;
        const int8_t v = (int8_t)(best + 1);
        if (l2[ci * fov + cj] == 0) l2[ci * fov + cj] = v;
        else if (ci == sx && cj == sy) {}
        else if (ci + 1 < fov && l2[(ci + 1) * fov + cj] == 0) l2[(ci + 1) * fov + cj] = v;
        else if (ci - 1 >= 0 && l2[(ci - 1) * fov + cj] == 0) l2[(ci - 1) * fov + cj] = v;
        else if (cj + 1 < fov && l2[ci * fov + cj + 1] == 0) l2[ci * fov + cj + 1] = v;
        else if (cj - 1 >= 0 && l2[ci * fov + cj - 1] == 0) l2[ci * fov + cj - 1] = v;
    }
    int8_t* l3 = rec + 3 * f2;
    if (env_blocks != nullptr) {
        for (int b = 0; b < cfg.n_blocks; ++b) {
            const int bx = env_blocks[2 * b], by = env_blocks[2 * b + 1];
            for (int q = 0; q < 4; ++q) {
                const int ci = bx + (q >> 1), cj = by + (q & 1);
                if (ci < fov && cj < fov) l3[ci * fov + cj] = 1;
            }
        }
    }
    {
        const int lb = hf - x, rb = hf - (W - 1 - x), ub = hf - y, db = hf - (Lc - 1 - y);
        int r_lo = 0, r_hi = 0, q_lo = 0, q_hi = 0;
        if (lb > 0) { r_hi = min(lb, fov); } else if (rb > 0) { r_lo = max(fov - rb, 0); r_hi = fov; }
        if (ub > 0) { q_hi = min(ub, fov); } else if (db > 0) { q_lo = max(fov - db, 0); q_hi = fov; }
        for (int b = r_lo * fov; b < r_hi * fov; ++b) l3[b] = 1;
        for (int r = 0; r < fov; ++r)
            for (int q = q_lo; q < q_hi; ++q) l3[r * fov + q] = 1;
    }
    rec[4 * f2] = (int8_t)(gy - y);
    rec[4 * f2 + 1] = (int8_t)(gx - x);
}

// Words of the optional degraded-cell bit map per env (dmfb_state_t.health_bits)
__host__ __device__ __forceinline__ int health_bit_words(const dmfb_cfg_t& cfg) { return (cfg.width * cfg.length + 31) >> 5; }

// Folds the usage log of env n into its counters (threads tid, tid + nthreads, ... of the caller cooperate).  The
// caller synchronises before anyone reads the counters and clears usage_log_len afterwards.  The log is read 8 entries
// per load where the env's slice is 16-byte aligned, and a thread issues all its loads before the first increment: a
// reset inside a step kernel sits on that launch's critical path, so what counts is the number of dependent DRAM round
// trips, not the bytes.
__device__ __forceinline__ void replay_usage_log(const dmfb_cfg_t& cfg, const dmfb_state_t& st, int64_t n, int tid,
                                                 int nthreads, int known_len = -1)
{
    const int A = cfg.n_agents, Lc = cfg.length;
    const int len = min(known_len >= 0 ? known_len : st.usage_log_len[n], st.usage_log_cap);
    const uint16_t* log = st.usage_log + (size_t)n * st.usage_log_cap * A;
    uint32_t* usage = st.usage + (size_t)n * cfg.width * Lc;
    const int total = len * A;
    int done = 0;
    if ((reinterpret_cast<uintptr_t>(log) & 15u) == 0) {
        const uint4* log4 = reinterpret_cast<const uint4*>(log);
        const int n4 = total >> 3;
        constexpr int kBatch = 4;
        for (int base = 0; base < n4; base += kBatch * nthreads) {
            uint4 v[kBatch];
#pragma unroll
            for (int b = 0; b < kBatch; ++b) {
                const int k = base + b * nthreads + tid;
                v[b] = k < n4 ? __ldcg(log4 + k) : make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
            }
#pragma unroll
            for (int b = 0; b < kBatch; ++b) {
                const uint32_t w[4] = {v[b].x, v[b].y, v[b].z, v[b].w};
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const uint32_t c = (w[q >> 1] >> (16 * (q & 1))) & 0xFFFFu;
                    if (c != 0xFFFFu) atomicAdd(usage + (c & 255u) * Lc + (c >> 8), 1u);
                }
            }
        }
        done = n4 << 3;
    }
    for (int k = done + tid; k < total; k += nthreads) {
        const uint32_t c = log[k];
        if (c != 0xFFFFu) atomicAdd(usage + (c & 255u) * Lc + (c >> 8), 1u);
    }
}

// updateHealth (dmfb.py:465-471) of env n: cells with usage > 50 get health *= degrade (one IEEE multiply) and
// usage = 0.  Threads tid, tid + nthreads, ... cooperate.  The counters are scanned four cells per load, a batch of
// loads in flight per thread (see replay_usage_log); cells over the threshold are rare.
template <int kBatch = 4>
__device__ __forceinline__ void update_health_env(const dmfb_cfg_t& cfg, const dmfb_state_t& st, int64_t n, int tid, int nthreads)
{
    const int cells = cfg.width * cfg.length;
    uint32_t* usage = st.usage + (size_t)n * cells;
    double* health = st.health ? st.health + (size_t)n * cells : nullptr;
    const double* degrade = st.degrade ? st.degrade + (size_t)n * cells : nullptr;
    uint32_t* bits = st.health_bits ? st.health_bits + (size_t)n * health_bit_words(cfg) : nullptr;
    auto hit = [&](int k) {
        if (health) {
            const double h = health[k] * (degrade ? degrade[k] : 1.0);
            health[k] = h;
            if (bits && h != 1.0) atomicOr(bits + (k >> 5), 1u << (k & 31));
        }
        usage[k] = 0;
    };
    int done = 0;
    if ((reinterpret_cast<uintptr_t>(usage) & 15u) == 0) {
        const uint4* u4 = reinterpret_cast<const uint4*>(usage);
        const int n4 = cells >> 2;
        for (int base = 0; base < n4; base += kBatch * nthreads) {
            uint4 v[kBatch];
#pragma unroll
            for (int b = 0; b < kBatch; ++b) {
                const int k = base + b * nthreads + tid;
                v[b] = k < n4 ? __ldcg(u4 + k) : make_uint4(0u, 0u, 0u, 0u);
            }
#pragma unroll
            for (int b = 0; b < kBatch; ++b) {
                const int k = 4 * (base + b * nthreads + tid);
                if (v[b].x > 50u) hit(k);
                if (v[b].y > 50u) hit(k + 1);
                if (v[b].z > 50u) hit(k + 2);
                if (v[b].w > 50u) hit(k + 3);
            }
        }
        done = n4 << 2;
    }
    for (int k = done + tid; k < cells; k += nthreads)
        if (__ldcg(usage + k) > 50u) hit(k);
}

// Log replay + updateHealth of env n in ONE pass over the counters, through a histogram in shared memory: `hist`
// (16-bit counts, two cells per word; all zero on entry and on exit) is the step kernel's not-yet-painted tile.  The
// env's log entries are counted with shared-memory atomics, then every thread adds the histogram to four counters per
// load, applies the threshold (usage > 50 -> health *= degrade, usage = 0; dmfb.py:465-471) and writes the four counters
// back where something changed.  Replaying the log with one global RED per entry instead - 2,000 per env on a 50x50
// chip with 10 droplets, 660 K scattered atomics per step of 64K staggered envs - cost 12 of C3's 53 us per step
// (gpurun_out/s3_c3diag2.txt: 41.3 us with updateHealth compiled out).  An episode adds at most one count per cell and
// step, so 16 bits are enough for any usage_log_cap <= 65,535.
__device__ __forceinline__ void update_health_env_hist(const dmfb_cfg_t& cfg, const dmfb_state_t& st, int64_t n, int tid,
                                                       int nthreads, uint32_t* hist, int log_len)
{
    const int A = cfg.n_agents, Lc = cfg.length, cells = cfg.width * cfg.length;
    const uint16_t* log = st.usage_log + (size_t)n * st.usage_log_cap * A;
    const int total = min(log_len, st.usage_log_cap) * A;
    auto count = [&](uint32_t c) {
        if (c != 0xFFFFu) {
            const uint32_t k = (c & 255u) * (uint32_t)Lc + (c >> 8);
            atomicAdd(hist + (k >> 1), 1u << (16u * (k & 1u)));
        }
    };
    auto count2 = [&](uint32_t w) {
        count(w & 0xFFFFu);
        count(w >> 16);
    };
    int done = 0;
    if ((reinterpret_cast<uintptr_t>(log) & 15u) == 0) {
        const uint4* log4 = reinterpret_cast<const uint4*>(log);
        const int n4 = total >> 3;
        for (int base = 0; base < n4; base += 2 * nthreads) {
            const int k0 = base + tid, k1 = k0 + nthreads;
            const uint4 none = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
            const uint4 v0 = k0 < n4 ? __ldcg(log4 + k0) : none, v1 = k1 < n4 ? __ldcg(log4 + k1) : none;
            count2(v0.x); count2(v0.y); count2(v0.z); count2(v0.w);
            count2(v1.x); count2(v1.y); count2(v1.z); count2(v1.w);
        }
        done = n4 << 3;
    }
    for (int k = done + tid; k < total; k += nthreads) count(log[k]);
    __syncthreads();

    uint32_t* usage = st.usage + (size_t)n * cells;
    double* health = st.health ? st.health + (size_t)n * cells : nullptr;
    const double* degrade = st.degrade ? st.degrade + (size_t)n * cells : nullptr;
    uint32_t* bits = st.health_bits ? st.health_bits + (size_t)n * health_bit_words(cfg) : nullptr;
    auto settle = [&](uint32_t sum, int k) -> uint32_t {     // one cell: new counter value
        if (sum <= 50u) return sum;
        if (health) {
            const double h = health[k] * (degrade ? degrade[k] : 1.0);
            health[k] = h;
            if (bits && h != 1.0) atomicOr(bits + (k >> 5), 1u << (k & 31));
        }
        return 0u;
    };
    int scanned = 0;
    if ((reinterpret_cast<uintptr_t>(usage) & 15u) == 0) {
        uint4* u4 = reinterpret_cast<uint4*>(usage);
        uint2* h2 = reinterpret_cast<uint2*>(hist);
        const int n4 = cells >> 2;
        for (int base = 0; base < n4; base += 2 * nthreads) {
            const int k0 = base + tid, k1 = k0 + nthreads;
            const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
            const uint4 a = k0 < n4 ? __ldcg(u4 + k0) : zero, b = k1 < n4 ? __ldcg(u4 + k1) : zero;
            auto fold4 = [&](int k4, const uint4& u) {
                if (k4 >= n4) return;
                const uint2 h = h2[k4];
                if ((h.x | h.y) == 0u && u.x <= 50u && u.y <= 50u && u.z <= 50u && u.w <= 50u) return;   // untouched
                h2[k4] = make_uint2(0u, 0u);
                const int k = 4 * k4;
                uint4 o;
                o.x = settle(u.x + (h.x & 0xFFFFu), k);
                o.y = settle(u.y + (h.x >> 16), k + 1);
                o.z = settle(u.z + (h.y & 0xFFFFu), k + 2);
                o.w = settle(u.w + (h.y >> 16), k + 3);
                u4[k4] = o;
            };
            fold4(k0, a);
            fold4(k1, b);
        }
        scanned = n4 << 2;
    }
    for (int k = scanned + tid; k < cells; k += nthreads) {
        const uint32_t h = (hist[k >> 1] >> (16 * (k & 1))) & 0xFFFFu;
        const uint32_t u = __ldcg(usage + k);
        if (h == 0u && u <= 50u) continue;
        usage[k] = settle(u + h, k);
    }
    __syncthreads();
    for (int k = scanned + tid; k < cells; k += nthreads) hist[k >> 1] = 0u;   // the few words of the tail cells
    if (tid == 0) st.usage_log_len[n] = 0;
    __syncthreads();
}

// updateHealth for the envs of the tile flagged kFlagNewTask; whole CTA cooperates.  FUSED: inside a step launch, where
// this is the tail of the CTA (and, for the last CTAs, of the launch): the log lengths come from shared memory
// (S.loglen) instead of a dependent global read, and the tile - zero-filled, not painted yet - serves as the histogram
// of update_health_env_hist when the chip fits (two cells per word).
template <bool FUSED = false>
__device__ __forceinline__ void update_health_flagged(const dmfb_cfg_t& cfg, const dmfb_state_t& st, const TileSmem& S,
                                                      int64_t n0, int e_valid, uint32_t tile_bytes = 0)
{
    const bool logged = st.usage_log != nullptr && st.usage_log_len != nullptr;
#ifndef DMFB_UPDATE_HEALTH_ATOMICS
    if (FUSED && logged && (uint32_t)((cfg.width * cfg.length + 1) / 2) * 4u <= tile_bytes && st.usage_log_cap <= 65535) {
        for (int e = 0; e < e_valid; ++e)
            if (S.flag[e] & kFlagNewTask)
                update_health_env_hist(cfg, st, n0 + e, (int)threadIdx.x, (int)blockDim.x,
                                       reinterpret_cast<uint32_t*>(S.tile), S.loglen[e]);
        return;
    }
#endif
    if (logged) {   // the counters are read below: fold the log in first
        for (int e = 0; e < e_valid; ++e)
            if (S.flag[e] & kFlagNewTask)
                replay_usage_log(cfg, st, n0 + e, (int)threadIdx.x, (int)blockDim.x, FUSED ? S.loglen[e] : -1);
        __syncthreads();
        for (int e = (int)threadIdx.x; e < e_valid; e += (int)blockDim.x)
            if (S.flag[e] & kFlagNewTask) st.usage_log_len[n0 + e] = 0;
    }
    for (int e = 0; e < e_valid; ++e)
        if (S.flag[e] & kFlagNewTask) update_health_env<4>(cfg, st, n0 + e, (int)threadIdx.x, (int)blockDim.x);
}

// ------------------------------------------------------------------------ step --
// Per-lane view of DMFBenv.step: what one lane (droplet g.i of env n) loads, computes and writes.
struct LaneIn {
    uint32_t d;         // packed droplet word
    int a;              // action
    double draw;        // injected move-success draw (DEG only)
    int sc_in, cum_in;  // leader lane only: step_count, cumulative constraints
    uint32_t episode;   // leader lane only
    int frozen_i;       // leader lane only
    uint32_t next_cur;  // leader lane only: task-prefetch cursor (dmfb_state_t.next_cursor)
    int log_len;        // leader lane only: entries in the env's usage log (DEG only)
};

struct LaneOut {
    uint32_t word;      // droplet word after the step (after the reset when the env was auto-reset)
    double r;           // float64 reward
    uint32_t done_mask;
    int sc_out, cum, constraints, success, term, log_len;
    bool frozen, do_reset, cursor_dirty;
    uint32_t cursor;    // leader lane: dmfb_state_t.next_cursor after the reset logic (run_ahead() continues from it)
    float team;
    uint32_t episode;
};

// One coalesced round trip for everything a tile needs from global memory.
template <bool DEG_T>
__device__ __forceinline__ LaneIn load_lane_inputs(const dmfb_state_t& st, const void* __restrict__ actions, int aes,
                                                   const double* __restrict__ u, uint32_t flags, int64_t n, size_t ja,
                                                   bool lane_on, bool leader)
{
    LaneIn in;
    in.d = 0; in.a = 0; in.draw = 0.0; in.sc_in = 0; in.cum_in = 0; in.episode = 0; in.frozen_i = 0; in.next_cur = 0; in.log_len = 0;
    if (lane_on) {
        in.d = reinterpret_cast<const uint32_t*>(st.drop)[ja];
        in.a = load_action(actions, aes, ja);
        if (DEG_T && st.health && u) in.draw = u[ja];
    }
    if (leader) {
        in.sc_in = st.step_count[n];
        in.cum_in = st.constraints[n];
        if (st.episode) in.episode = st.episode[n];
        if (flags & DMFB_STEP_FREEZE_TERM) in.frozen_i = st.terminated[n];
        if ((flags & DMFB_STEP_AUTO_RESET) && st.next_cursor) in.next_cur = st.next_cursor[n];
        // loaded here, with the other inputs, instead of where it is used: one dependent round trip less per step
        if (DEG_T && st.usage && st.usage_log_len != nullptr) in.log_len = st.usage_log_len[n];
    }
    return in;
}

// moveDroplets + step bookkeeping (+ fused auto-reset) for the droplet held by this lane.  Registers and
// warp shuffles only; every lane of the warp must call.
template <int G, int A_T, bool DEG_T>
__device__ __forceinline__ LaneOut dmfb_dynamics(const dmfb_cfg_t& cfg, const dmfb_state_t& st, const Group<G>& g, int A,
                                                 int64_t n, size_t ja, bool env_on, bool lane_on, LaneIn in,
                                                 const double* __restrict__ u, uint64_t seed, uint32_t flags,
                                                 int32_t* status_flag, CoordSets<A_T>& cs, const uint32_t* env_bits)
{
    const int W = cfg.width, Lc = cfg.length;
    const uint32_t all_mask = (A >= 32) ? 0xFFFFFFFFu : ((1u << A) - 1u);
    const uint32_t d = in.d;
    const int a = in.a;
    double prob = 1.0, draw = in.draw;
    const bool have_prob = DEG_T && st.health != nullptr;
    if (have_prob && lane_on) {  // getMoveProb (:361-363): the cell occupied at the start of the step
        const uint32_t cell = (d & 255u) * (uint32_t)Lc + ((d >> 8) & 255u);
        // a clear bit in the (L2-sized) degraded-cell map means health == 1.0 exactly: no gather from the big array
        bool degraded = true;
        if (env_bits) degraded = (env_bits[cell >> 5] >> (cell & 31u)) & 1u;
        if (degraded) prob = st.health[(size_t)n * W * Lc + cell];
    }
    const int sc_in = g.get(in.sc_in, 0);
    const int cum_in = g.get(in.cum_in, 0);
    const uint32_t episode = g.get(in.episode, 0);
    const bool frozen = g.get(in.frozen_i, 0) != 0;  // lock-step padding (rollout.py:131-141)
    const int sc = sc_in + 1;                         // dmfb.py:561

    // ---- moveOneDroplet for all droplets (:325-359) ------------------------------------------------
    const uint32_t goal = d >> 16;
    const uint32_t start_cell = d & 0xFFFFu;          // "past" position
    const int x = d & 255u, y = (d >> 8) & 255u, gx = goal & 255u, gy = goal >> 8;
    const int od = abs(x - gx) + abs(y - gy);         // Droplet.distance (:93-95)
    const bool pre_done = (od == 0);                  // getTaskStatus before the moves (:278)
    const bool stalled = cfg.stall && pre_done;       // reward 0, no move, no draw (:331-332)
    bool pass = true;                                 // random.random() <= prob (:335)
    if (have_prob && lane_on && !stalled) {
        if (u) {
            pass = draw <= prob;
        } else if (prob < 1.0) {
            // A draw in [0,1) never exceeds prob >= 1, and the stream is counter based: nothing to consume.
            // Otherwise u53(r.x, r.y) = m / 2^53 with the 53-bit integer m, and m / 2^53 <= prob  <=>
            // m <= floor(prob * 2^53); the scaling is exact, so the integer compare equals the float64 one.
            const uint4 r = env_random(seed, kStreamMove, cfg.env_base + n, episode, (uint32_t)sc, (uint32_t)g.i);
            const uint64_t m = ((uint64_t)(r.x >> 5) << 26) | (uint64_t)(r.y >> 6);
            pass = prob >= 0.0 && m <= (uint64_t)__double2ll_rd(prob * 9007199254740992.0);
        }
    }
    const bool tries = lane_on && !stalled && !frozen && pass;
    uint32_t cand = start_cell;
    if (tries) {                                      // Droplet.move (:103-124)
        int nx = x + (a == 1) - (a == 2), ny = y + (a == 4) - (a == 3);
        nx = min(max(nx, 0), W - 1);
        ny = min(max(ny, 0), Lc - 1);
        cand = (uint32_t)nx | ((uint32_t)ny << 8);
        if ((unsigned)a > 4u && status_flag) atomicOr(status_flag, 1);     // TypeError('action is illegal') (:115-116)
        if (A_T == 0 && cfg.n_blocks) {                                    // _isTouchingBlocks -> revert (:338-340)
            const uint8_t* bl = st.blocks + (size_t)n * cfg.n_blocks * 2;
            for (int b = 0; b < cfg.n_blocks; ++b)
                if ((unsigned)(nx - (int)bl[2 * b]) <= 1u && (unsigned)(ny - (int)bl[2 * b + 1]) <= 1u) cand = start_cell;
        }
    }
    // Sequential resolution (:279-283): droplet i moves only if its candidate cell is not occupied by any
    // other droplet at that moment (j < i already moved, j > i still at their old cell) (:341-343).
    uint32_t cur = lane_on ? start_cell : (0xFF00u | (uint32_t)g.lane);   // idle lanes: unique off-chip cells
#pragma unroll
    for (int i = 0; i < A; ++i) {
        const uint32_t ci = g.get(cand, i);
        const unsigned taken = g.ballot(g.i != i && cur == ci);
        if (g.i == i && taken == 0u) cur = ci;
    }
    const int nx = cur & 255u, ny = cur >> 8;
    const int nd = abs(nx - gx) + abs(ny - gy);
    double r;                                         // base reward (:345-354)
    if (stalled) r = 0.0;
    else if (nd == od && od == 0) r = -0.1;
    else if (nd == od && a == 0) r = -0.25;
    else if (nd < od) r = -0.1;
    else r = -0.4;

    // ---- comflic_static / comflic_dynamic (:254-271): final vs final, saved vs final ---------------
    int sta = 0, dyn = 0;
    if constexpr (A_T != 0) {
        // all droplets of the env at once: coordinates published in shared memory, four others per VABSDIFF4
        const uint32_t pad = 128u + (uint32_t)(cfg.fov >> 1);
        if (lane_on) {
            cs.put(0, 1, g.i, (uint32_t)nx, (uint32_t)ny, pad);
            cs.put(2, 3, g.i, (uint32_t)x, (uint32_t)y, pad);
        }
        __syncwarp();
        cs.load_current();
        uint32_t px[CoordSets<A_T>::kWords], py[CoordSets<A_T>::kWords];
#pragma unroll
        for (int w = 0; w < CoordSets<A_T>::kWords; ++w) { px[w] = cs.word(2, w); py[w] = cs.word(3, w); }
        // a droplet is always within one cell of itself and of its own past cell: the three "- 1"
        sta = CoordSets<A_T>::count_near((uint32_t)nx, (uint32_t)ny, cs.x, cs.y) - 1;              // cur_me ~ cur_j
        dyn = CoordSets<A_T>::count_near((uint32_t)x, (uint32_t)y, cs.x, cs.y) - 1                 // past_me ~ cur_j
              + CoordSets<A_T>::count_near((uint32_t)nx, (uint32_t)ny, px, py) - 1;                // cur_me ~ past_j
    } else {
        const uint32_t both = cur | (start_cell << 16);   // my (current, past) cells
#pragma unroll
        for (int j = 0; j < A; ++j) {
            const uint32_t bj = g.get(both, j);                 // (current_j, past_j)
            const uint32_t h1 = near_pair(both, dup_lo(bj));    // bit0: cur_me~cur_j, bit1: past_me~cur_j
            const uint32_t h2 = near_pair(both, dup_hi(bj));    // bit0: cur_me~past_j
            const uint32_t other = (j != g.i) ? 1u : 0u;
            sta += (int)(h1 & other);
            dyn += (int)((h1 >> 1) & other) + (int)(h2 & other);
        }
    }
    if (!lane_on) { sta = 0; dyn = 0; }
    const int constraints = g.sum(sta + dyn);                              // (:287)
    const bool post_done = (cur == goal);
    const uint32_t post_mask = g.ballot(lane_on && post_done);
    const bool all_done = (post_mask == all_mask);                         // np.all(getTaskStatus()) (:293)
    r = r - (double)(2 * sta);                                             // rewards - 2*sta - 2*dy, float64 (:288)
    r = r - (double)(2 * dyn);
    if (cfg.stall && pre_done) r = 0.0;                                    // (:289-292)
    if (all_done) {                                                        // (:293-296)
        r = r + 10.0;
        if (constraints == 0) r = r + 10.0;
    }
    if (frozen || !lane_on) r = 0.0;

    LaneOut o;
    o.team = g.sum((float)r) / (float)A;                                   // rollout.py:33 (f32 output, 1e-6 contract)
    o.r = r;
    o.frozen = frozen;
    o.episode = episode;
    // ---- DMFBenv.step bookkeeping (:572-586) --------------------------------------------------------
    o.cum = cum_in + (frozen ? 0 : constraints);
    o.sc_out = frozen ? sc_in : sc;
    o.constraints = frozen ? 0 : constraints;
    o.done_mask = post_mask;
    o.success = 0;
    o.cursor = 0u;
    o.cursor_dirty = false;
    o.log_len = 0;
    if (sc < cfg.max_step) o.success = (all_done && o.cum == 0) ? 1 : 0;
    else o.done_mask = all_mask;
    if (frozen) { o.done_mask = all_mask; o.success = 0; }
    o.term = (o.done_mask == all_mask) ? 1 : 0;

#ifdef DMFB_HEALTH_PREFETCH_NEXT
    // The next step's getMoveProb reads health at the cell the droplet stands on NOW: if that cell is degraded, ask L2
    // for its line, so that the gather - a dependent access of the next step's warp - does not go to DRAM
    if (have_prob && lane_on && env_bits && !frozen && !post_done) {
        const uint32_t c2 = (uint32_t)nx * (uint32_t)Lc + (uint32_t)ny;
        if ((env_bits[c2 >> 5] >> (c2 & 31u)) & 1u)
            asm volatile("prefetch.global.L2 [%0];" :: "l"(st.health + (size_t)n * W * Lc + c2));
    }
#endif
    if (DEG_T && (flags & DMFB_STEP_RECORD_USAGE) && st.usage) {   // addUsage (:459-463): droplets not done after the move
        const bool add = lane_on && !frozen && !post_done;
        bool logged = false;
        if (st.usage_log != nullptr && st.usage_log_len != nullptr) {
            const int len = g.get(in.log_len, 0);
            logged = len < st.usage_log_cap;                       // a full log falls back to direct increments
            o.log_len = len + (logged ? 1 : 0);
            if (logged && lane_on && !frozen) {
                st.usage_log[((size_t)n * st.usage_log_cap + len) * A + g.i] = add ? (uint16_t)(nx | (ny << 8)) : (uint16_t)0xFFFFu;
                if (g.i == 0) st.usage_log_len[n] = len + 1;
            }
        }
        if (add && !logged) atomicAdd(st.usage + ((size_t)n * W + nx) * Lc + ny, 1u);   // result unused -> RED.ADD
    } else if (DEG_T && (flags & DMFB_STEP_AUTO_RESET) && st.usage && st.usage_log != nullptr && st.usage_log_len != nullptr) {
        o.log_len = g.get(in.log_len, 0);                          // record=False: a fused reset still replays the log
    }

    // ---- fused auto-reset: DMFBenv.reset(new=False) (:589-597) for envs that just terminated -------
    o.word = (d & 0xFFFF0000u) | cur;
    o.do_reset = (flags & DMFB_STEP_AUTO_RESET) && o.term && !frozen && env_on;
    if (flags & DMFB_STEP_AUTO_RESET) {
        const bool prefetch = st.next_task != nullptr && st.next_cursor != nullptr;
        uint32_t cur = prefetch ? g.get(in.next_cur, 0) : 0u;     // search state of THIS env's next task
        const uint32_t cur_in = cur;
        bool need = o.do_reset;
        if (prefetch && need && (cur & kTaskReady)) {             // the search ran ahead and is finished: pick it up
            if (lane_on) o.word = st.next_task[ja];
            need = false;
        }
        // otherwise finish it now, from the first attempt nobody has examined yet
        o.word = generate_layout<G>(cfg, st, g, A, seed, cfg.env_base + n - g.idx, episode + 1u, need, cur & ~kTaskReady, o.word);
        if (o.do_reset) cur = 0u;                                 // the search for the episode after the new one starts over
        if (A_T == 0) generate_blocks<G>(cfg, st, g, seed, n, episode + 1u, o.do_reset, lane_on, o.word);
        if (o.do_reset) {
            o.sc_out = 0;
            o.cum = 0;
            if (lane_on && st.start) reinterpret_cast<uint16_t*>(st.start)[ja] = (uint16_t)(o.word & 0xFFFFu);
        }
        o.cursor = cur;
        o.cursor_dirty = cur != cur_in;
    }
    return o;
}

// Run-ahead task search of a fused auto-reset step (dmfb_state_t.next_task / next_cursor), called by every warp at the
// very END of the step kernel - after the tile's bulk store has been issued, so that the search overlaps the drain of
// the observation and no other warp waits for it at a barrier.  Per warp and step ONE env is served, the open search
// whose env is closest to its step limit, with up to prefetch_rounds() rounds of attempts (8x that when the env has
// at most 8 steps left: what remains of its search is then spread over its last steps instead of running inside the
// launch that resets it).  An env needs 1/p_accept attempts per episode (about 12 for 4 droplets on 10x10, 70 for 10
// on 20x20) and has the whole episode to find them.
template <int G>
__device__ __forceinline__ void run_ahead(const dmfb_cfg_t& cfg, const dmfb_state_t& st, const Group<G>& g, int A,
                                          uint64_t seed, int64_t n, size_t ja, bool env_on, bool lane_on, const LaneOut& o)
{
    if (st.next_task == nullptr || st.next_cursor == nullptr) return;
    uint32_t cur = o.cursor;
    const uint32_t per_round = (uint32_t)attempts_per_round<G>(A);
    const bool open = env_on && g.i == 0 && !(cur & kTaskReady) && cur < kMaxSamplerRounds * per_round;
    const uint32_t rem = (uint32_t)max(cfg.max_step - o.sc_out, 0);
    const uint32_t best = __reduce_min_sync(kFull, open ? ((rem << 5) | (uint32_t)g.lane) : 0xFFFFFFFFu);
    bool dirty = o.cursor_dirty;
    if (best != 0xFFFFFFFFu) {
        const uint32_t rounds = (uint32_t)prefetch_rounds(A) * ((best >> 5) <= 8u ? 8u : 1u);
        const int src = (int)(best & 31u), dst = src / G;
        const uint32_t epi = __shfl_sync(kFull, o.episode + 1u + (o.do_reset ? 1u : 0u), src);
        const uint32_t at = __shfl_sync(kFull, cur, src);
        const int64_t env = cfg.env_base + n - g.idx + dst;
        uint32_t task = 0, used = 0;
        bool hit = false;
        if (A == 10) {
            used = sample_rounds_warp<G, 10>(cfg, g, layout_stream(seed, env, epi), at, rounds, dst, task, hit);
        } else {
            for (; used < rounds && !hit; ++used) hit = sample_round<G>(cfg, g, A, seed, env, epi, at + used * per_round, dst, task);
        }
        if (g.idx == dst) {
            if (hit && lane_on) st.next_task[ja] = task;
            cur = hit ? kTaskReady : at + used * per_round;
            dirty = true;
        }
    }
    if (env_on && g.i == 0 && dirty) st.next_cursor[n] = cur;
}

// Coalesced write-back of the state and of the small per-step outputs (consecutive lanes -> consecutive
// agents / envs).
__device__ __forceinline__ void write_back_lane(const dmfb_state_t& st, const dmfb_out_t& out, const LaneOut& o, int64_t n,
                                                size_t ja, int i, bool lane_on, bool leader)
{
    if (lane_on) {
        if (!o.frozen) reinterpret_cast<uint32_t*>(st.drop)[ja] = o.word;
        if (out.reward) out.reward[ja] = (float)o.r;
        if (out.reward_f64) out.reward_f64[ja] = o.r;
        if (out.done) out.done[ja] = (uint8_t)((o.done_mask >> i) & 1u);
    }
    if (leader) {
        st.step_count[n] = o.sc_out;
        st.constraints[n] = o.cum;
        st.terminated[n] = (uint8_t)(o.term && !o.do_reset);
        if (o.do_reset && st.episode) st.episode[n] = o.episode + 1u;
        if (out.team_reward) out.team_reward[n] = o.team;
        if (out.constraints) out.constraints[n] = o.constraints;
        if (out.success) out.success[n] = (uint8_t)o.success;
        if (out.terminated) out.terminated[n] = (uint8_t)o.term;
        if (out.padded) out.padded[n] = (uint8_t)o.frozen;
    }
}

// avail mask of a tile: all ones; zeros for padded envs (rollout.py:22,138-139)
__device__ __forceinline__ void write_avail(const dmfb_cfg_t& cfg, const dmfb_out_t& out, int A, int64_t n0, int e_valid,
                                            int any_frozen, int tid, int nthreads, int agent, bool lane_on, bool frozen)
{
    if (!out.avail) return;
    const int per_env = A * cfg.n_actions;
    uint8_t* gav = out.avail + (size_t)n0 * per_env;
    const int nbytes = e_valid * per_env;
    if (!any_frozen && (nbytes & 15) == 0 && (reinterpret_cast<uintptr_t>(gav) & 15) == 0) {
        const uint4 ones = make_uint4(0x01010101u, 0x01010101u, 0x01010101u, 0x01010101u);
#pragma unroll 1
        for (int k = tid; k < (nbytes >> 4); k += nthreads) reinterpret_cast<uint4*>(gav)[k] = ones;
    } else if (lane_on) {
#pragma unroll 1
        for (int k = 0; k < cfg.n_actions; ++k) gav[agent * cfg.n_actions + k] = frozen ? 0 : 1;
    }
}

// A_T / E_T: compile-time droplet count and tile size for the shipped configs (0 = run-time values).
// DEG_T = false strips the degradation path (health / usage / draws) when the state has none.
// The degrading instances of the shipped tiles may use the 72 registers that their shared-memory residency leaves free
// (9 CTAs of 96 threads, 13 of 64): left alone, ptxas settled on 56 and spilled a dozen values on the hot path.
template <int FOV_T, int G, int A_T, int E_T, bool DEG_T>
__global__ void __launch_bounds__(E_T ? cta_threads(E_T, G) : kMaxThreads,
                                  (E_T && DEG_T) ? 65536 / (72 * cta_threads(E_T ? E_T : 1, G)) : 0)
dmfb_step_kernel(const __grid_constant__ dmfb_cfg_t cfg, const dmfb_state_t st, const void* __restrict__ actions,
                 int aes, const double* __restrict__ u, uint64_t seed, uint32_t flags, const dmfb_out_t out, int E_rt)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int E = E_T ? E_T : E_rt;
    const int A = A_T ? A_T : cfg.n_agents, W = cfg.width, Lc = cfg.length;
    const TileLayout L(E, A, FOV_T ? FOV_T : cfg.fov, W, Lc, FOV_T ? 3 * FOV_T * FOV_T + 2 : cfg.obs_dim);
    const TileSmem S(smem_raw, L);
    const int tid = (int)threadIdx.x;
    const Group<G> g(tid);
    const int64_t n0 = (int64_t)blockIdx.x * E;
    const int e_valid = (int)min((int64_t)E, (int64_t)st.n_envs - n0);
    const int e = g.env_in_tile(tid);               // env of this lane group inside the tile
    const int64_t n = n0 + e;
    const bool env_on = g.valid && e < e_valid;
    const bool lane_on = env_on && g.i < A;         // this lane holds droplet g.i of env n
    const bool leader = env_on && g.i == 0;
    const int agent = e * A + g.i;                  // agent index inside the tile
    const size_t ja = (size_t)n * A + g.i;          // index into [N,A] tensors

    // Programmatic dependent launch: let the next kernel of the stream (normally the next env step) start its
    // own prologue while this grid drains; everything that touches global memory comes after the wait, which
    // returns only once the preceding grid has completed and its writes are visible.
    asm volatile("griddepcontrol.launch_dependents;");
    load_tables(cfg, L, S, tid, (int)blockDim.x);
    if constexpr (FOV_T != 0 && A_T != 0 && E_T != 0)
        zero_tile_static<((E_T * A_T * (3 * FOV_T * FOV_T + 2) + 15) / 16) * 16, cta_threads(E_T, G)>(S.tile, tid);
    else
        zero_tile(L, S, tid, (int)blockDim.x);
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const LaneIn in = load_lane_inputs<DEG_T>(st, actions, aes, u, flags, n, ja, lane_on, leader);
#ifndef DMFB_NO_BITS_PREFETCH
    if (DEG_T && st.health_bits && st.health && lane_on) {
        // the droplet's bit lies somewhere in the env's map (a few 128-byte lines): ask for those lines together with
        // the inputs, so that the lookup - whose address needs the droplet word - finds them close by
        const int hb_bytes = health_bit_words(cfg) * 4;
        const char* hb = reinterpret_cast<const char*>(st.health_bits) + (size_t)n * hb_bytes;
        for (int k = g.i * 128; k < hb_bytes; k += A * 128) asm volatile("prefetch.global.L1 [%0];" :: "l"(hb + k));
    }
#endif
#ifndef DMFB_WHATIF_NO_RESET_PREFETCH
    if (DEG_T && (flags & DMFB_STEP_AUTO_RESET) && st.usage) {
        // An env on its last step before the limit is reset in this launch for certain: updateHealth will then scan its
        // counters (and replay its usage log), a chain of dependent DRAM round trips at the very end of the CTA - the
        // tail of the whole launch.  Ask L2 for those lines now, while the dynamics run.
        const int sc_now = g.get(in.sc_in, 0);
        if (env_on && sc_now + 1 >= cfg.max_step) {
            const char* ub = reinterpret_cast<const char*>(st.usage + (size_t)n * W * Lc);
            for (int k = g.i * 128; k < W * Lc * 4; k += A * 128) asm volatile("prefetch.global.L2 [%0];" :: "l"(ub + k));
            if (st.usage_log != nullptr) {
                const char* lb = reinterpret_cast<const char*>(st.usage_log + (size_t)n * st.usage_log_cap * A);
                for (int k = g.i * 128; k < st.usage_log_cap * A * 2; k += A * 128)
                    asm volatile("prefetch.global.L2 [%0];" :: "l"(lb + k));
            }
        }
    }
#endif

    CoordSets<A_T> cs(S.sets, env_on ? e : 0);
    // degraded-cell bit map of this lane's env.  (Staging the maps of a warp's envs in shared memory with one coalesced
    // load was measured and dropped: C3 47.3 -> 48.4 us fresh, 64.5 -> 71.9 us with every cell degraded.)
    const uint32_t* env_bits = nullptr;
    if (DEG_T && st.health_bits && st.health) env_bits = st.health_bits + (size_t)(env_on ? n : 0) * health_bit_words(cfg);
    const LaneOut o = dmfb_dynamics<G, A_T, DEG_T>(cfg, st, g, A, n, ja, env_on, lane_on, in, u, seed, flags, out.status, cs,
                                                   env_bits);
    if constexpr (A_T == 10) {
        // the paint reads the current positions from the sets: an env that was just reset has new ones
        if ((flags & DMFB_STEP_AUTO_RESET) && __any_sync(kFull, o.do_reset)) {
            __syncwarp();
            if (lane_on && o.do_reset)
                cs.put(0, 1, g.i, o.word & 255u, (o.word >> 8) & 255u, 128u + (uint32_t)(cfg.fov >> 1));
            __syncwarp();
            cs.load_current();
        }
    }
    write_back_lane(st, out, o, n, ja, g.i, lane_on, leader);
    if (leader) {
        S.flag[e] = (uint8_t)(o.do_reset ? kFlagNewTask : 0);
        if (DEG_T) S.loglen[e] = o.log_len;
    }
    const int any_frozen = __syncthreads_or(o.frozen && env_on);   // also the zero-fill / table barrier

    write_avail(cfg, out, A, n0, e_valid, any_frozen, tid, (int)blockDim.x, agent, lane_on, o.frozen);
    if (DEG_T && (flags & DMFB_STEP_AUTO_RESET) && st.usage) {
        // rare: only tiles in which an env was just reset scan its electrodes (updateHealth, dmfb.py:465-471)
#ifndef DMFB_WHATIF_NO_UPDATE_HEALTH
        if (__syncthreads_or(o.do_reset)) update_health_flagged<true>(cfg, st, S, n0, e_valid, L.tile_bytes);
#else
        if (leader && o.do_reset && st.usage_log_len) st.usage_log_len[n] = 0;   // timing experiment
#endif
    }

    const uint32_t word = o.word;
    const uint8_t* env_blocks = (A_T == 0 && cfg.n_blocks && env_on) ? st.blocks + (size_t)n * cfg.n_blocks * 2 : nullptr;
    bool v01 = false;
    if constexpr (FOV_T == 0 && A_T == 0) v01 = cfg.obs_version == DMFB_OBS_V01;   // generic instance only
    if (v01)
        paint_agent_v01(cfg, L, S, agent, g.i, word, lane_on && !o.frozen, [&](int j) { return g.get(word, j); }, env_blocks);
    else
        paint_agent<FOV_T, A_T>(cfg, L, S, agent, g.i, word, lane_on && !o.frozen, [&](int j) { return g.get(word, j); },
                                cs, env_blocks);
    const bool in_flight = store_tile_issue(out.obs + (size_t)n0 * A * L.D, S.tile, (uint32_t)(e_valid * A * L.D));
    if (flags & DMFB_STEP_AUTO_RESET) {
#ifdef DMFB_WHATIF_NO_RUNAHEAD
        flags |= kStepNoRunAhead;
#endif
        if (!(flags & kStepNoRunAhead)) run_ahead<G>(cfg, st, g, A, seed, n, ja, env_on, lane_on, o);
        else if (leader && o.cursor_dirty && st.next_cursor) st.next_cursor[n] = o.cursor;
    }
    if (in_flight) tma_store_wait_read_all();
}

// --------------------------------------------------------------------- reset --

// mode 0: reset (new task), mode 1: restart (back to start cells), mode 2: observe only
template <int FOV_T, int G>
__global__ void __launch_bounds__(kMaxThreads)
dmfb_reset_kernel(const __grid_constant__ dmfb_cfg_t cfg, const dmfb_state_t st, const uint8_t* __restrict__ mask,
                  int mode, int new_task, const uint8_t* __restrict__ layouts, const uint8_t* __restrict__ block_layouts,
                  const double* __restrict__ degrade_in, uint64_t seed, int8_t* __restrict__ obs, int E)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const TileLayout L(cfg, E);
    const TileSmem S(smem_raw, L);
    const int tid = (int)threadIdx.x;
    const Group<G> g(tid);
    const int A = L.A;
    const int64_t n0 = (int64_t)blockIdx.x * E;
    const int e_valid = (int)min((int64_t)E, (int64_t)st.n_envs - n0);
    const int e = g.env_in_tile(tid);
    const int64_t n = n0 + e;
    const bool env_on = g.valid && e < e_valid;
    const bool lane_on = env_on && g.i < A;
    const bool leader = env_on && g.i == 0;
    const size_t ja = (size_t)n * A + g.i;
    const int cells = cfg.width * cfg.length;

    int sel_i = 0;
    if (leader) sel_i = (mask == nullptr) || (mask[n] != 0);
    const int n_selected = __syncthreads_count(sel_i);
    if (n_selected == 0) return;  // nothing to do in this tile (the common case of a masked reset)
    const bool selected = g.get(sel_i, 0) != 0;

    load_tables(cfg, L, S, tid, (int)blockDim.x);
    if (obs) zero_tile(L, S, tid, (int)blockDim.x);

    uint32_t word = lane_on ? reinterpret_cast<const uint32_t*>(st.drop)[ja] : 0u;
    if (mode == 0) {
        uint32_t episode = 0;
        if (leader && st.episode) episode = st.episode[n] + 1u;
        episode = g.get(episode, 0);
        if (layouts) {
            if (lane_on && selected) word = reinterpret_cast<const uint32_t*>(layouts)[ja];
        } else {
            // a search that ran ahead (dmfb_state_t.next_task) is for exactly this episode: pick it up or resume it
            const bool prefetch = st.next_task != nullptr && st.next_cursor != nullptr;
            uint32_t cur = 0;
            if (prefetch && leader && selected) cur = st.next_cursor[n];
            cur = g.get(cur, 0);
            bool need = selected && env_on;
            if (need && (cur & kTaskReady)) {
                if (lane_on) word = st.next_task[ja];
                need = false;
            }
            word = generate_layout<G>(cfg, st, g, A, seed, cfg.env_base + n - g.idx, episode, need, cur & ~kTaskReady, word);
        }
        // the episode number moves on: whatever was prefetched for it is used up (or overridden by an injected task)
        if (leader && selected && st.next_cursor) st.next_cursor[n] = 0u;
        if (cfg.n_blocks) {   // refresh() regenerates the obstacles with every task (dmfb.py:174-177)
            if (block_layouts) {
                if (leader && selected)
                    for (int b = 0; b < 2 * cfg.n_blocks; ++b)
                        st.blocks[(size_t)n * cfg.n_blocks * 2 + b] = block_layouts[(size_t)n * cfg.n_blocks * 2 + b];
            } else {
                generate_blocks<G>(cfg, st, g, seed, n, episode, selected && env_on, lane_on, word);
            }
        }
        if (leader && selected && st.episode) st.episode[n] = episode;
        if (lane_on && selected && st.start) reinterpret_cast<uint16_t*>(st.start)[ja] = (uint16_t)(word & 0xFFFFu);
    } else if (mode == 1) {  // restart: droplets back to their start cells (dmfb.py:185-190)
        if (lane_on && selected) word = (word & 0xFFFF0000u) | reinterpret_cast<const uint16_t*>(st.start)[ja];
    }
    if (mode != 2 && selected) {
        if (lane_on) reinterpret_cast<uint32_t*>(st.drop)[ja] = word;
        if (leader) {
            st.step_count[n] = 0;
            st.constraints[n] = 0;
            st.terminated[n] = 0;
        }
    }
    if (leader) S.flag[e] = selected ? (kFlagSelected | kFlagNewTask) : 0;
    __syncthreads();

    // refresh(new) (dmfb.py:174-183): new -> health=1, usage=0, degrade redrawn; else updateHealth (:465-471)
    if (mode == 0 && (st.usage || st.health || st.degrade)) {
        if (new_task) {
            for (int ee = 0; ee < e_valid; ++ee) {
                if (!S.flag[ee]) continue;
                const int64_t nn = n0 + ee;
                uint32_t* usage = st.usage ? st.usage + (size_t)nn * cells : nullptr;
                double* health = st.health ? st.health + (size_t)nn * cells : nullptr;
                double* degrade = st.degrade ? st.degrade + (size_t)nn * cells : nullptr;
                const uint32_t episode = st.episode ? st.episode[nn] : 0u;
                if (threadIdx.x == 0 && st.usage_log_len) st.usage_log_len[nn] = 0;   // usage = 0: the log goes with it
                if (st.health_bits && health)
                    for (int k = threadIdx.x; k < health_bit_words(cfg); k += blockDim.x)
                        st.health_bits[(size_t)nn * health_bit_words(cfg) + k] = 0u;
                for (int k = threadIdx.x; k < cells; k += blockDim.x) {
                    if (usage) usage[k] = 0;
                    if (health) health[k] = 1.0;
                    if (degrade) {
                        double dg = 1.0;
                        if (degrade_in) {
                            dg = degrade_in[(size_t)nn * cells + k];
                        } else if (cfg.b_degrade) {  // _random_health_statue (dmfb.py:157-164)
                            const uint4 r = env_random(seed, kStreamDegrade, cfg.env_base + nn, episode, (uint32_t)k, 0u);
                            dg = __dadd_rn(__dmul_rn(u53(r.x, r.y), 0.4), 0.6);   // rand * 0.4 + 0.6 as two roundings, like NumPy (no FMA)
                            if (u53(r.z, r.w) < 1.0 - cfg.per_degrade) dg = 1.0;
                        }
                        degrade[k] = dg;
                    }
                }
            }
        } else if (st.usage) {
            update_health_flagged(cfg, st, S, n0, e_valid);
        }
    }
    if (obs == nullptr) return;
    const bool on = lane_on && selected;
    auto get = [&](int j) { return g.get(word, j); };
    const uint8_t* env_blocks = (cfg.n_blocks && env_on) ? st.blocks + (size_t)n * cfg.n_blocks * 2 : nullptr;
    bool v01 = false;
    if constexpr (FOV_T == 0) v01 = cfg.obs_version == DMFB_OBS_V01;
    if (v01) paint_agent_v01(cfg, L, S, e * A + g.i, g.i, word, on, get, env_blocks);
    else paint_agent<FOV_T, 0>(cfg, L, S, e * A + g.i, g.i, word, on, get, CoordSets<0>(nullptr, 0), env_blocks);
    int8_t* gobs = obs + (size_t)n0 * A * L.D;
    if (n_selected == e_valid) store_tile(gobs, S.tile, (uint32_t)(e_valid * A * L.D));
    else store_rows_masked(gobs, S.tile, e_valid, A * L.D, S.flag);
}

// --------------------------------------------------------------- flush usage --
__global__ void __launch_bounds__(128)
dmfb_flush_usage_kernel(const __grid_constant__ dmfb_cfg_t cfg, const dmfb_state_t st)
{
    const int64_t n = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);   // one warp per env
    if (n >= st.n_envs) return;
    replay_usage_log(cfg, st, n, (int)(threadIdx.x & 31), 32);
    __syncwarp();
    if ((threadIdx.x & 31) == 0) st.usage_log_len[n] = 0;
}

// --------------------------------------------------------------- task search --
// Run-ahead task search for 10 droplets as a kernel of its own, launched by dmfb_step right after a fused auto-reset
// step.  Whole-set rejection at 1.4 % acceptance (10 droplets, 20x20) costs ~30 % of the step kernel's instructions
// when it runs inside the step kernel, where the warps that search also delay their CTAs (barrier, tail of the
// launch: C2 37 -> 75 us with staggered resets); here every warp does the same bounded amount of work and the kernel
// overlaps the step kernels of the other sub-batches.
//
// Grid: a CTA of 8 warps looks after kSearchEnvs = 128 consecutive envs.  Thread t reads the cursor of env t (coalesced),
// the ~3 % of the envs whose next task is still unknown are compacted into a list in shared memory, and the warps
// serve that list round-robin, one env per warp and turn; most CTAs have at most one turn per warp, a CTA without an
// open search leaves after the one load.  (With one WARP per env - 8,192 CTAs per step of 64K envs that had next to
// nothing to do - the launch itself cost 11 us per step: 45.4 us with every search finished at once against 34.1 us
// with the same searches and the kernel not launched, gpurun_out/s3_diag.txt.  512 CTAs per step do not.)
//
// On chips of at most 30x30 cells a warp examines 32 CONSECUTIVE attempts at once, ONE PER LANE, against a private
// bit board in shared memory (column `lane` of board[row][32]: no bank conflicts, no synchronisation): row x+1 holds,
// at bit y+1, the cells within Chebyshev distance 1 of the points drawn so far, so a point is tested with one load and
// one shift and added with three read-modify-writes - exactly the reference's "min pairwise squared distance > 2" rule
// (dmfb.py:220), point by point.  The lowest accepted attempt of the 32 wins, which is the attempt the one-per-warp
// flavour (sample_rounds_warp: same counters, same cells) would have stopped at: the tasks are the same whoever
// searches.  ~30 instructions per attempt instead of ~55, and 32 attempts in the latency of 16.
constexpr int kBoardRows = 32;                                   // W + 2 rows, W <= 30
#ifndef DMFB_SEARCH_ENVS
#define DMFB_SEARCH_ENVS 64
#endif
constexpr int kSearchEnvs = DMFB_SEARCH_ENVS;                    // envs per CTA of the search kernel (128 / 64 / 32: 40.1 / 38.6 / 39.6 us)
constexpr int kSearchWarps = 8;
__device__ __forceinline__ uint32_t attempt_cell(uint64_t z, uint32_t W, uint32_t Lc)
{
    return __umulhi((uint32_t)z, W) | (__umulhi((uint32_t)(z >> 32), Lc) << 8);
}

// 32 * ceil(rounds / 32) attempts of env n from attempt `cur` on, one per lane.  `col` = this lane's board column.
__device__ __forceinline__ void search_env_lanewise(const dmfb_cfg_t& cfg, const dmfb_state_t& st, uint64_t base, int64_t n,
                                                    uint32_t cur, uint32_t rounds, uint32_t* col, int lane)
{
    constexpr uint64_t kPhi = 0x9E3779B97F4A7C15ull;
    constexpr int P = 20;
    const uint32_t W = (uint32_t)cfg.width, Lc = (uint32_t)cfg.length;
    const uint32_t batches = (rounds + 31u) / 32u;
    for (uint32_t b = 0; b < batches; ++b, cur += 32u) {
        const uint64_t ctr = base + (uint64_t)(cur + (uint32_t)lane) * (uint64_t)P * kPhi;
        for (uint32_t r = 0; r < W + 2u; ++r) col[r * 32u] = 0u;
        bool alive = true;
#pragma unroll 1
        for (int p = 0; p < P; p += 4) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint64_t z = mix64(ctr + (uint64_t)(p + q + 1) * kPhi);
                const uint32_t x = __umulhi((uint32_t)z, W), y = __umulhi((uint32_t)(z >> 32), Lc);
                if (alive) {
                    uint32_t* row = col + x * 32u;                // rows x, x+1, x+2 = cells x-1, x, x+1
                    const uint32_t mid = row[32];
                    if ((mid >> (y + 1u)) & 1u) {
                        alive = false;
                    } else {
                        const uint32_t m = 7u << y;               // cells y-1, y, y+1 at bits y .. y+2
                        row[0] |= m;
                        row[32] = mid | m;
                        row[64] |= m;
                    }
                }
            }
#ifdef DMFB_WHATIF_SEARCH_FREE
            alive = true;                                         // timing experiment: every search ends at once
            break;
#endif
            if (!__any_sync(kFull, alive)) break;
        }
        const unsigned ok = __ballot_sync(kFull, alive);
        if (ok) {
            const uint64_t cw = base + (uint64_t)(cur + (uint32_t)(__ffs(ok) - 1)) * (uint64_t)P * kPhi;
            if (lane < 10) {
                const uint32_t s = attempt_cell(mix64(cw + (uint64_t)(2 * lane + 1) * kPhi), W, Lc);
                const uint32_t t = attempt_cell(mix64(cw + (uint64_t)(2 * lane + 2) * kPhi), W, Lc);
                st.next_task[(size_t)n * 10 + lane] = s | (t << 16);
            }
            if (lane == 0) st.next_cursor[n] = kTaskReady;
            return;
        }
    }
    if (lane == 0) st.next_cursor[n] = cur;
}

__global__ void __launch_bounds__(kSearchWarps * 32)
dmfb_task_search_kernel(const __grid_constant__ dmfb_cfg_t cfg, const dmfb_state_t st, uint64_t seed, uint32_t rounds)
{
    __shared__ uint32_t board[kSearchWarps][kBoardRows][32];
    __shared__ uint32_t open_cur[kSearchEnvs];
    __shared__ uint8_t open_env[kSearchEnvs];
    __shared__ int n_open;
    // the next step kernel of the stream may start its prologue now; this kernel's own loads wait for the step kernel
    // in front of it (programmatic dependent launch on both sides)
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int tid = (int)threadIdx.x, warp = tid >> 5;
    const Group<32> g(tid);
    const int64_t n0 = (int64_t)blockIdx.x * kSearchEnvs;
    if (tid == 0) n_open = 0;
    __syncthreads();
    if (tid < kSearchEnvs) {                                      // warps 0-3: one env per thread
        uint32_t cur = kTaskReady;
        if (n0 + tid < st.n_envs) cur = st.next_cursor[n0 + tid];
        const bool open = !(cur & kTaskReady) && cur < kMaxSamplerRounds;
        const unsigned m = __ballot_sync(kFull, open);
        int at = 0;
        if (g.lane == 0 && m) at = atomicAdd(&n_open, __popc(m));
        at = __shfl_sync(kFull, at, 0) + __popc(m & ((1u << g.lane) - 1u));
        if (open) {
            open_env[at] = (uint8_t)tid;
            open_cur[at] = cur;
        }
    }
    __syncthreads();
    const int todo = n_open;
    for (int k = warp; k < todo; k += kSearchWarps) {
        const int64_t n = n0 + open_env[k];
        const uint32_t cur = open_cur[k];
        uint32_t epi = 0;
        if (g.lane == 0 && st.episode) epi = st.episode[n] + 1u;
        epi = __shfl_sync(kFull, epi, 0);
        const uint64_t base = layout_stream(seed, cfg.env_base + n, epi);
#ifndef DMFB_SEARCH_WARPWISE
        if (cfg.width <= kBoardRows - 2 && cfg.length <= 30) {
            search_env_lanewise(cfg, st, base, n, cur, rounds, &board[warp][0][g.lane], g.lane);
            continue;
        }
#endif
        uint32_t task = 0;
        bool hit = false;
        const uint32_t used = sample_rounds_warp<32, 10>(cfg, g, base, cur, rounds, 0, task, hit);
        if (hit && g.lane < 10) st.next_task[(size_t)n * 10 + g.lane] = task;
        if (g.lane == 0) st.next_cursor[n] = hit ? kTaskReady : cur + used;
    }
}

// ---------------------------------------------------------- health bit map --
// Rebuilds dmfb_state_t.health_bits from health (bit k of an env = health[k] != 1.0): one warp per env.
__global__ void __launch_bounds__(128)
dmfb_health_bits_kernel(const __grid_constant__ dmfb_cfg_t cfg, const dmfb_state_t st)
{
    const int64_t n = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (n >= st.n_envs) return;
    const int lane = threadIdx.x & 31, cells = cfg.width * cfg.length, nw = health_bit_words(cfg);
    const double* health = st.health + (size_t)n * cells;
    uint32_t* bits = st.health_bits + (size_t)n * nw;
    for (int w = 0; w < nw; ++w) {
        const int k = w * 32 + lane;
        const unsigned m = __ballot_sync(kFull, k < cells && health[k] != 1.0);
        if (lane == 0) bits[w] = m;
    }
}

// ----------------------------------------------------------------- get_state --
// getglobalobs (dmfb.py:368-392): (3,W,L) per env, int8.  Tile of E2 envs staged in smem.
__global__ void __launch_bounds__(128)
dmfb_global_state_kernel(const __grid_constant__ dmfb_cfg_t cfg, const dmfb_state_t st, int8_t* __restrict__ out,
                         int E2, uint32_t tile_bytes)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int8_t* tile = reinterpret_cast<int8_t*>(smem_raw);
    const int64_t n0 = (int64_t)blockIdx.x * E2;
    const int e_valid = (int)min((int64_t)E2, (int64_t)st.n_envs - n0);
    const int A = cfg.n_agents, W = cfg.width, Lc = cfg.length;
    const int per_env = 3 * W * Lc;
    uint4* t4 = reinterpret_cast<uint4*>(tile);
    for (int k = threadIdx.x; k < (int)(tile_bytes >> 4); k += blockDim.x) t4[k] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
    // one thread per env keeps the "later index overwrites" order of the reference loop (:383-387)
    for (int e = threadIdx.x; e < e_valid; e += blockDim.x) {
        const uint32_t* gdrop = reinterpret_cast<const uint32_t*>(st.drop) + (size_t)(n0 + e) * A;
        int8_t* g = tile + (size_t)e * per_env;
        for (int i = 0; i < A; ++i) {
            const uint32_t d = gdrop[i];
            g[(d & 255u) * Lc + ((d >> 8) & 255u)] = (int8_t)(i + 1);
            g[W * Lc + ((d >> 16) & 255u) * Lc + (d >> 24)] = (int8_t)(i + 1);
        }
        for (int b = 0; b < cfg.n_blocks; ++b) {   // add_blocks_In_gloabal_Obs (:376-381)
            const uint8_t* bl = st.blocks + ((size_t)(n0 + e) * cfg.n_blocks + b) * 2;
            for (int q = 0; q < 4; ++q) g[2 * W * Lc + (bl[0] + (q >> 1)) * Lc + bl[1] + (q & 1)] = 1;
        }
    }
    store_tile(out + (size_t)n0 * per_env, tile, (uint32_t)(e_valid * per_env));
}

int group_size_for(int n_agents)
{
    if (n_agents == 10) return 10;   // three envs per warp instead of two with 16-lane groups (C2 / C3)
    int g = 4;
    while (g < n_agents) g <<= 1;
    return g;
}

// Envs per tile: a multiple of the 16-byte alignment unit, at most kMaxThreads worth of lane groups, ~40 KB of smem.
int tile_envs_for(const dmfb_cfg_t& cfg, int G)
{
    const int row = cfg.n_agents * cfg.obs_dim;
    const int per_warp = 32 / G;
    int max_envs = (kMaxThreads / 32) * per_warp;
    // measured: C1 best with 16 envs per CTA (8: 11.5 us, 32: 11.3 us against 10.3 as 4 sub-batches), C2/C3 with 8
    const int cap = G <= 4 ? 16 : (G == 10 ? 8 : 8 * per_warp);
    if (cap >= 16 / gcd_int(16, row) && cap < max_envs) max_envs = cap;
    static const int forced = getenv("DMFB_TILE_ENVS") ? atoi(getenv("DMFB_TILE_ENVS")) : 0;   // tuning knob, read once
    if (forced > 0 && cta_threads(forced, G) <= kMaxThreads) return forced;
    return pick_tile_envs(row, 40 * 1024, max_envs);
}

template <typename K>
int set_smem(K kernel, uint32_t bytes)
{
    if (bytes > 48 * 1024)
        DMFB_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return DMFB_OK;
}

int check_common(const dmfb_cfg_t* cfg, const dmfb_state_t* st)
{
    if (cfg && st && st->n_envs == 0) return DMFB_OK;   // empty batch: nothing to check, the callers return early
    if (!cfg || !st || st->n_envs < 0 || !st->drop || !st->step_count || !st->constraints || !st->terminated) {
        snprintf(g_last_error, sizeof(g_last_error), "null cfg/state pointer");
        return DMFB_ERR_BAD_ARG;
    }
    if (cfg->n_blocks != 0 && !st->blocks) {
        snprintf(g_last_error, sizeof(g_last_error), "n_blocks > 0 needs state->blocks");
        return DMFB_ERR_BAD_ARG;
    }
    return DMFB_OK;
}

// Calls f.template operator()<FOV_T, G>() with the compile-time specialisation matching (fov, G).
template <typename F>
int dispatch(int fov, int G, F&& f)
{
#define DMFB_G_CASES(FOVT)                             \
    switch (G) {                                       \
    case 4: return f.template operator()<FOVT, 4>();   \
    case 8: return f.template operator()<FOVT, 8>();   \
    case 10: return f.template operator()<FOVT, 10>(); \
    case 16: return f.template operator()<FOVT, 16>(); \
    default: return f.template operator()<FOVT, 32>(); \
    }
    switch (fov) {
    case 5: DMFB_G_CASES(5)
    case 7: DMFB_G_CASES(7)
    case 9: DMFB_G_CASES(9)
    default: DMFB_G_CASES(0)
    }
#undef DMFB_G_CASES
}

struct StepLaunch {
    const dmfb_cfg_t* cfg; const dmfb_state_t* st; const void* actions; int aes; const double* u;
    uint64_t seed; uint32_t flags; const dmfb_out_t* out; cudaStream_t s; int E, grid; uint32_t smem;
    template <int FOVT, int G, int AT, int ET, bool DEG>
    int go() const {
        int rc = set_smem(dmfb_step_kernel<FOVT, G, AT, ET, DEG>, smem);
        if (rc) return rc;
        cudaLaunchConfig_t lc{};
        lc.gridDim = dim3((unsigned)grid);
        lc.blockDim = dim3((unsigned)cta_threads(E, G));
        lc.dynamicSmemBytes = smem;
        lc.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        static const bool no_pdl = getenv("DMFB_NO_PDL") != nullptr;
        lc.attrs = attr;
        lc.numAttrs = no_pdl ? 0 : 1;
        DMFB_CUDA_TRY(cudaLaunchKernelEx(&lc, dmfb_step_kernel<FOVT, G, AT, ET, DEG>, *cfg, *st, actions, aes, u, seed,
                                         flags, *out, E));
        return DMFB_OK;
    }
    template <int FOVT, int G>
    int operator()() const {
        // fully specialised instances for the shipped benchmark configs (BASELINE.json C1, C2, C3)
        const bool deg = st->health != nullptr || st->usage != nullptr;
        if (cfg->n_blocks != 0) return go<FOVT, G, 0, 0, true>();
        if constexpr (FOVT == 9 && G == 4) {
            if (cfg->n_agents == 4 && E == 32) return deg ? go<9, 4, 4, 32, true>() : go<9, 4, 4, 32, false>();
            if (cfg->n_agents == 4 && E == 16) return deg ? go<9, 4, 4, 16, true>() : go<9, 4, 4, 16, false>();
        }
        if constexpr (FOVT == 9 && G == 10) {
            if (E == 24) return deg ? go<9, 10, 10, 24, true>() : go<9, 10, 10, 24, false>();
            if (E == 16) return deg ? go<9, 10, 10, 16, true>() : go<9, 10, 10, 16, false>();
            if (E == 8) return deg ? go<9, 10, 10, 8, true>() : go<9, 10, 10, 8, false>();
        }
        return go<FOVT, G, 0, 0, true>();
    }
};

struct ResetLaunch {
    const dmfb_cfg_t* cfg; const dmfb_state_t* st; const uint8_t* mask; int mode, new_task; const uint8_t* layouts;
    const uint8_t* block_layouts; const double* degrade; uint64_t seed; int8_t* obs; cudaStream_t s; int E, grid;
    uint32_t smem;
    template <int FOVT, int G>
    int operator()() const {
        int rc = set_smem(dmfb_reset_kernel<FOVT, G>, smem);
        if (rc) return rc;
        dmfb_reset_kernel<FOVT, G><<<grid, cta_threads(E, G), smem, s>>>(*cfg, *st, mask, mode, new_task, layouts, block_layouts,
                                                               degrade, seed, obs, E);
        return DMFB_OK;
    }
};

int launch_reset(const dmfb_cfg_t* cfg, const dmfb_state_t* state, const uint8_t* mask, int mode, int new_task,
                 const uint8_t* layouts, const uint8_t* block_layouts, const double* degrade, uint64_t seed, int8_t* obs,
                 void* stream)
{
    int rc = check_common(cfg, state);
    if (rc) return rc;
    if (state->n_envs == 0) return DMFB_OK;
    const int G = group_size_for(cfg->n_agents);
    const int E = tile_envs_for(*cfg, G);
    const TileLayout L(*cfg, E);
    ResetLaunch job{cfg, state, mask, mode, new_task, layouts, block_layouts, degrade, seed, obs,
                    static_cast<cudaStream_t>(stream), E, (state->n_envs + E - 1) / E, L.total};
    rc = dispatch(cfg->obs_version == DMFB_OBS_V01 ? 0 : cfg->fov, G, job);   // v0_1: generic instance only
    if (rc) return rc;
    g_launches.fetch_add(1);
    DMFB_CUDA_TRY(cudaGetLastError());
    return DMFB_OK;
}

}  // namespace
}  // namespace dmfb

using namespace dmfb;

extern "C" {

int dmfb_abi_version(void) { return DMFB_ABI_VERSION; }
const char* dmfb_last_cuda_error(void) { return g_last_error; }
uint64_t dmfb_launch_count(void) { return g_launches.load(); }

int dmfb_cfg_init(dmfb_cfg_t* cfg, int width, int length, int n_agents, int n_blocks, int fov, int stall,
                  int b_degrade, double per_degrade)
{
    if (!cfg) return DMFB_ERR_BAD_ARG;
    memset(cfg, 0, sizeof(*cfg));
    if (width < 5 || length < 5) return DMFB_ERR_CHIP_TOO_SMALL;              // dmfb.py:489
    if (n_agents <= 0) return DMFB_ERR_BAD_ARG;                               // dmfb.py:490
    if (fov > (width < length ? width : length)) return DMFB_ERR_FOV_TOO_LARGE;  // dmfb.py:139-140
    if (n_agents > (int)((width + 1) * (length + 1) / 9)) return DMFB_ERR_TOO_MANY_DROPLETS;  // dmfb.py:144-146
    if (width > DMFB_MAX_DIM || length > DMFB_MAX_DIM || n_agents > DMFB_MAX_AGENTS || fov < 1 || fov > DMFB_MAX_FOV)
        return DMFB_ERR_BAD_ARG;
    const int hf = fov / 2;
    if (hf == 10) return DMFB_ERR_DIV_ZERO;
    if (n_blocks < 0 || n_blocks > DMFB_MAX_BLOCKS) return DMFB_ERR_BAD_ARG;
    if ((double)n_blocks * 4.0 / (double)(width * length) > 0.2) n_blocks = 0;   // 'Too many required modules' (dmfb.py:232-234)
    cfg->width = width; cfg->length = length; cfg->n_agents = n_agents; cfg->n_blocks = n_blocks; cfg->fov = fov;
    cfg->stall = stall ? 1 : 0; cfg->b_degrade = b_degrade ? 1 : 0; cfg->per_degrade = per_degrade;
    cfg->max_step = 2 * (width + length);
    cfg->n_actions = 5;
    cfg->obs_dim = 3 * fov * fov + 2;
    cfg->obs_version = DMFB_OBS_BASE;
    cfg->l2_words = (fov * fov + 31) / 32;
    cfg->env_base = 0;
    // direction table (dmfb.py:442-454): float64 division + round-half-even, as python's round()
    for (int axis = 0; axis < 2; ++axis) {
        const int dim = axis == 0 ? width : length;
        int8_t* tab = axis == 0 ? cfg->dir_x : cfg->dir_y;
        const double scale = (double)(dim - hf) / (double)(10 - hf);
        for (int d = -(dim - 1); d <= dim - 1; ++d) {
            int v = d;
            if (d > hf) v = (int)__builtin_rint((double)(d - hf) / scale) + hf;
            else if (d < -hf) v = (int)__builtin_rint((double)(d + hf) / scale) - hf;
            tab[d + dim - 1] = (int8_t)v;
        }
    }
    // boundary bit patterns (dmfb.py:428-439)
    for (int c = 0; c <= 2 * hf; ++c) {
        for (int xq = 0; xq < fov; ++xq)
            for (int yq = 0; yq < fov; ++yq) {
                const int q = xq * fov + yq;
                const bool row_on = c == 0 ? false : (c <= hf ? xq < c : xq >= fov - (c - hf));
                const bool col_on = c == 0 ? false : (c <= hf ? yq < c : yq >= fov - (c - hf));
                if (row_on) cfg->l2_row[c][q >> 5] |= 1u << (q & 31);
                if (col_on) cfg->l2_col[c][q >> 5] |= 1u << (q & 31);
            }
    }
    return DMFB_OK;
}

int dmfb_flush_usage(const dmfb_cfg_t* cfg, const dmfb_state_t* state, void* stream)
{
    int rc = check_common(cfg, state);
    if (rc) return rc;
    if (state->n_envs == 0 || !state->usage || !state->usage_log || !state->usage_log_len) return DMFB_OK;
    dmfb_flush_usage_kernel<<<(state->n_envs + 3) / 4, 128, 0, static_cast<cudaStream_t>(stream)>>>(*cfg, *state);
    g_launches.fetch_add(1);
    DMFB_CUDA_TRY(cudaGetLastError());
    return DMFB_OK;
}

int dmfb_sync_health_bits(const dmfb_cfg_t* cfg, const dmfb_state_t* state, void* stream)
{
    int rc = check_common(cfg, state);
    if (rc) return rc;
    if (state->n_envs == 0 || !state->health || !state->health_bits) return DMFB_OK;
    dmfb_health_bits_kernel<<<(state->n_envs + 3) / 4, 128, 0, static_cast<cudaStream_t>(stream)>>>(*cfg, *state);
    g_launches.fetch_add(1);
    DMFB_CUDA_TRY(cudaGetLastError());
    return DMFB_OK;
}

int dmfb_cfg_set_obs_version(dmfb_cfg_t* cfg, int obs_version)
{
    if (!cfg || cfg->fov < 1 || (obs_version != DMFB_OBS_BASE && obs_version != DMFB_OBS_V01)) return DMFB_ERR_BAD_ARG;
    cfg->obs_version = obs_version;
    cfg->obs_dim = (obs_version == DMFB_OBS_V01 ? 4 : 3) * cfg->fov * cfg->fov + 2;
    return DMFB_OK;
}

int dmfb_step(const dmfb_cfg_t* cfg, const dmfb_state_t* state, const void* actions, int action_elem_size,
              const double* u_inject, uint64_t seed, uint32_t flags, const dmfb_out_t* out, void* stream)
{
    int rc = check_common(cfg, state);
    if (rc) return rc;
    if (state->n_envs == 0) return DMFB_OK;
    if (!actions || !out || !out->obs || (action_elem_size != 1 && action_elem_size != 4 && action_elem_size != 8)) {
        snprintf(g_last_error, sizeof(g_last_error), "dmfb_step: bad actions/out");
        return DMFB_ERR_BAD_ARG;
    }
    const int G = group_size_for(cfg->n_agents);
    const int E = tile_envs_for(*cfg, G);
    const TileLayout L(*cfg, E);
    // 10 droplets with fused auto-reset: the run-ahead task search is a second, small kernel (dmfb_task_search_kernel)
    static const bool search_in_step = getenv("DMFB_SEARCH_IN_STEP") != nullptr;   // experiment knob
    // ... where whole-set rejection is expensive: the expected number of conflicting pairs among the 2A points,
    // C(2A,2) * 9 / (W*L), is 4.3 on a 20x20 chip (1.4 % of the attempts accepted) and 0.7 on 50x50 (50 %), where the
    // few attempts cost less inside the step kernel than a second launch does
    const double conflicts = 0.5 * (2.0 * cfg->n_agents) * (2.0 * cfg->n_agents - 1.0) * 9.0 / ((double)cfg->width * cfg->length);
    const bool search_kernel = (flags & DMFB_STEP_AUTO_RESET) && cfg->n_agents == 10 && conflicts > 2.8 && state->next_task &&
                               state->next_cursor && !search_in_step;
    const bool skip_search_once = (flags & DMFB_STEP_SKIP_TASK_SEARCH) != 0u;
    const uint32_t search_share = ((flags >> 8) & 0xFFu) ? ((flags >> 8) & 0xFFu) : 1u;
    flags &= ~(kStepNoRunAhead | DMFB_STEP_SKIP_TASK_SEARCH | DMFB_STEP_SEARCH_SHARE(0xFF));
    if (search_kernel) flags |= kStepNoRunAhead;
    StepLaunch job{cfg, state, actions, action_elem_size, u_inject, seed, flags, out, static_cast<cudaStream_t>(stream),
                   E, (state->n_envs + E - 1) / E, L.total};
    rc = dispatch(cfg->obs_version == DMFB_OBS_V01 ? 0 : cfg->fov, G, job);   // v0_1: generic instance only
    if (rc) return rc;
    g_launches.fetch_add(1);
    DMFB_CUDA_TRY(cudaGetLastError());
    static const bool skip_search = getenv("DMFB_WHATIF_SKIP_SEARCH") != nullptr;   // timing experiment
    if (search_kernel && !skip_search && !skip_search_once) {
        // 32 attempts per open env and step (one per lane): a search of 70 attempts on average is over ~3 steps after
        // the reset; a caller that launches the search every P-th step only asks for P times as many
        static const int rounds_knob = getenv("DMFB_SEARCH_ROUNDS") ? atoi(getenv("DMFB_SEARCH_ROUNDS")) : 0;   // tuning knob
        static const bool no_pdl = getenv("DMFB_NO_PDL") != nullptr;
        cudaLaunchConfig_t lc{};
        lc.gridDim = dim3((unsigned)((state->n_envs + kSearchEnvs - 1) / kSearchEnvs));
        lc.blockDim = dim3(kSearchWarps * 32);
        lc.stream = static_cast<cudaStream_t>(stream);
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        lc.attrs = attr;
        lc.numAttrs = no_pdl ? 0 : 1;
        DMFB_CUDA_TRY(cudaLaunchKernelEx(&lc, dmfb_task_search_kernel, *cfg, *state, seed,
                                         (uint32_t)(rounds_knob > 0 ? rounds_knob : 32) * search_share));
        g_launches.fetch_add(1);
    }
    return DMFB_OK;  // DMFB_STEP_AUTO_RESET is fused into the step kernel
}

int dmfb_reset(const dmfb_cfg_t* cfg, const dmfb_state_t* state, const uint8_t* mask, int new_task,
               const uint8_t* layouts, const uint8_t* block_layouts, const double* degrade, uint64_t seed, int8_t* obs,
               void* stream)
{
    return launch_reset(cfg, state, mask, 0, new_task, layouts, block_layouts, degrade, seed, obs, stream);
}

int dmfb_restart(const dmfb_cfg_t* cfg, const dmfb_state_t* state, const uint8_t* mask, int8_t* obs, void* stream)
{
    if (state && !state->start) {
        snprintf(g_last_error, sizeof(g_last_error), "dmfb_restart needs state->start");
        return DMFB_ERR_BAD_ARG;
    }
    return launch_reset(cfg, state, mask, 1, 0, nullptr, nullptr, nullptr, 0, obs, stream);
}

int dmfb_observe(const dmfb_cfg_t* cfg, const dmfb_state_t* state, int8_t* obs, void* stream)
{
    if (state && state->n_envs == 0) return DMFB_OK;
    if (!obs) return DMFB_ERR_BAD_ARG;
    return launch_reset(cfg, state, nullptr, 2, 0, nullptr, nullptr, nullptr, 0, obs, stream);
}

int dmfb_global_state(const dmfb_cfg_t* cfg, const dmfb_state_t* state, int8_t* out, void* stream)
{
    int rc = check_common(cfg, state);
    if (rc) return rc;
    if (state->n_envs == 0) return DMFB_OK;
    if (!out) return DMFB_ERR_BAD_ARG;
    const int per_env = 3 * cfg->width * cfg->length;
    const int E2 = pick_tile_envs(per_env, 32 * 1024, 64);
    const uint32_t tile_bytes = ((uint32_t)(E2 * per_env) + 15u) & ~15u;
    rc = set_smem(dmfb_global_state_kernel, tile_bytes);
    if (rc) return rc;
    const int grid = (state->n_envs + E2 - 1) / E2;
    dmfb_global_state_kernel<<<grid, 128, tile_bytes, static_cast<cudaStream_t>(stream)>>>(*cfg, *state, out, E2,
                                                                                            tile_bytes);
    g_launches.fetch_add(1);
    DMFB_CUDA_TRY(cudaGetLastError());
    return DMFB_OK;
}

}  // extern "C"
