// meda_kernels.cu — batched MEDA (micro-electrode-dot-array) environment step for sm_100a (B200).
//
// Replaces, for N independent chips at once, the reference call tree
//   MEDAEnv.step (env/MEDA/meda.py:513-539) -> RoutingTaskManager.moveDroplets (:241-259)
//   -> moveOneDroplet (:261-292) / getMoveProb (:302-309) / Droplet.move (:106-138) -> calPunish (:321-330)
//   -> getObs (:607-611) -> getOneObs (:613-674, or MEDAEnv_v0_1 :788-844 / MEDAEnv_v0_2 :850-897) -> addUsage (:591-598)
// plus MEDAEnv.reset (:541-550) with refresh/addTask (:161-185) and updateHealth (:600-605).
//
// Design: WARP-AUTONOMOUS.  A warp owns EW consecutive envs (EW chosen so that their observation span is a
// multiple of 16 bytes -> one TMA bulk store per warp) and runs the whole step for them without any CTA
// barrier: lane = droplet for the dynamics (droplets of a MEDA chip move independently, the reference has no
// collision prevention; the pairwise punish counts and env scalars are exchanged with shuffles), then
// lane = (env, agent, layer) for the observation: a 5x5 footprint seen through the window is a RECTANGLE
// (also when clipped onto the window), which one lane fills with predicated byte stores at immediate offsets;
// footprints of one layer are filled in the reference's order by the same lane, so "later index overwrites"
// holds by program order.  The CTA is only the unit that shares the shared-memory carve-up.
#include "common.cuh"

namespace dmfb {
namespace {

constexpr int kThreads = 256;
constexpr int kRad = 2;          // RoutingTaskManager.r (meda.py:150)

// Shared memory of the reset kernel: one CTA per tile of E envs.
struct MedaLayout {
    int E, A, D;
    uint32_t tile_bytes, off_word, off_flag, total;
    __host__ __device__ MedaLayout(const meda_cfg_t& c, int E_) {
        E = E_; A = c.n_agents; D = c.obs_dim;
        tile_bytes = ((uint32_t)(E * A * D) + 15u) & ~15u;
        uint32_t o = tile_bytes;
        off_word = o; o += (uint32_t)(E * A) * 4u;     // packed droplet words
        off_flag = o; o += ((uint32_t)E + 3u) & ~3u;   // per env flags
        total = (o + 15u) & ~15u;
    }
};

constexpr uint8_t kEnvSelected = 1;   // reset: env selected / step: env live (paint its rows)

struct MedaSmem {
    int8_t* tile;
    uint32_t* word;
    uint8_t* flag;
    __device__ MedaSmem(unsigned char* base, const MedaLayout& L) {
        tile = reinterpret_cast<int8_t*>(base);
        word = reinterpret_cast<uint32_t*>(base + L.off_word);
        flag = reinterpret_cast<uint8_t*>(base + L.off_flag);
    }
};

// Droplet.move (meda.py:106-138): step 3 on the axes, 2 on the diagonals, pushed back on chip
// (x against `length`, y against `width`).  Actions outside 0..8 fall through every branch like in the reference.
__device__ __forceinline__ void meda_move(int& xc, int& yc, int a, int width, int length)
{
    if (a == 8) return;
    const int dx = (a == 1) * 3 - (a == 3) * 3 + ((a == 4) | (a == 5)) * 2 - ((a == 6) | (a == 7)) * 2;
    const int dy = (a == 2) * 3 - (a == 0) * 3 + ((a == 5) | (a == 6)) * 2 - ((a == 4) | (a == 7)) * 2;
    xc += dx;
    yc += dy;
    if (xc + kRad >= length) xc = length - 1 - kRad; else if (xc - kRad < 0) xc = kRad;
    if (yc + kRad >= width) yc = width - 1 - kRad; else if (yc - kRad < 0) yc = kRad;
}

// The part of the 5x5 footprint centred on (X, Y) that shows in the window with origin (ox, oy), as the
// rectangle [xa, xb] x [ya, yb] of window cells.  Inside-only painting drops the cells outside the window;
// clipped painting (np.clip of the cell coordinates, meda.py:664-667,874-878) moves them onto the border, which
// is again a rectangle because clipping is monotone.  Empty iff xa > xb or ya > yb (never when clipped).
struct FootRect { int xa, xb, ya, yb; };

__device__ __forceinline__ FootRect foot_rect(int fov, int ox, int oy, int X, int Y, bool clip)
{
    const int nx0 = X - kRad - ox, ny0 = Y - kRad - oy;
    FootRect r;
    r.xa = max(nx0, 0); r.xb = min(nx0 + 2 * kRad, fov - 1);
    r.ya = max(ny0, 0); r.yb = min(ny0 + 2 * kRad, fov - 1);
    if (clip) {
        r.xa = min(r.xa, fov - 1); r.xb = max(r.xb, 0);
        r.ya = min(r.ya, fov - 1); r.yb = max(r.yb, 0);
    }
    return r;
}

// One lane fills the rectangle (at most 5x5) with `val`: 25 predicated byte stores at immediate offsets.
__device__ __forceinline__ void fill_rect(uint32_t layer, int fov, const FootRect& r, int val, bool active)
{
    const int w = r.xb - r.xa + 1, h = active ? r.yb - r.ya + 1 : 0;
    const uint32_t p = layer + (uint32_t)(r.ya * fov + r.xa);
#pragma unroll
    for (int dy = 0; dy <= 2 * kRad; ++dy) {
#pragma unroll
        for (int dx = 0; dx <= 2 * kRad; ++dx)
            if (dy < h && dx < w) sts_u8(p + (uint32_t)(dy * fov + dx), (uint32_t)val);
    }
}

// getOneObs for the agents of `n_env` consecutive envs, painted by ONE WARP into `tile` ([n_env][A][D], zeroed).
// words: packed droplets [n_env][A] (shared memory), flags[e] & kEnvSelected selects the envs to paint.
//   base  (MEDAEnv.getOneObs, meda.py:613-674): lane = (env, agent, layer 0..3): own droplet / own goal / the
//         other droplets ascending / ALL other goals ascending, clipped; the layer-0 lane adds dir_VEC (:672).
//   v0_2  (MEDAEnv_v0_2.getOneObs, meda.py:850-897): lane = (env, agent, layer 0..1): all droplets ascending /
//         goals of the OBSERVED others (droplets with a cell in the window) in python-set order, clipped; a second
//         pass, lane = (env, agent, rows | columns), writes the border layer and the direction bytes.
template <int VER, int A_T, int FOV_T>
__device__ __forceinline__ void meda_paint_warp(const meda_cfg_t& cfg, const uint32_t* words, const uint8_t* flags,
                                                int8_t* tile_ptr, int n_env, const uint8_t* __restrict__ set_order)
{
    const uint32_t tile = smem_u32(tile_ptr);
    const int A = A_T ? A_T : cfg.n_agents, fov = FOV_T ? FOV_T : cfg.fov, f2 = fov * fov, hf = fov >> 1;
    const int D = cfg.obs_dim;
    const int lane = threadIdx.x & 31;
    constexpr int PL = VER == MEDA_OBS_V02 ? 2 : 4;
    const int n_items = n_env * A * PL;
    for (int p0 = 0; p0 < n_items; p0 += 32) {
        const int p = p0 + lane;
        const bool valid = p < n_items;
        const int l = p & (PL - 1);
        const int g = valid ? p / PL : 0;
        const int e = g / A, i = g - e * A;
        const bool live = valid && (flags[e] & kEnvSelected);
        const uint32_t* w = words + e * A;
        const uint32_t me = w[i];
        const int cx = me & 255u, cy = (me >> 8) & 255u, gx = (me >> 16) & 255u, gy = me >> 24;
        const int ox = cx - hf, oy = cy - hf;
        const uint32_t rec = tile + (uint32_t)(g * D);
        const uint32_t layer = rec + (uint32_t)(l * f2);
        if (VER == MEDA_OBS_BASE) {
            const bool own = l < 2, goal = l & 1, clip = l == 3;
            const int n_iter = A > 1 ? A - 1 : 1;
            for (int k = 0; k < n_iter; ++k) {
                const int j = own ? i : k + (k >= i);                  // k-th other droplet, ascending
                const bool act = live && (own ? k == 0 : j < A);
                const uint32_t d = w[min(j, A - 1)];
                const int X = goal ? (d >> 16) & 255u : d & 255u, Y = goal ? d >> 24 : (d >> 8) & 255u;
                fill_rect(layer, fov, foot_rect(fov, ox, oy, X, Y, clip), j + 1, act);
            }
            if (live && l == 0) {
                sts_u8(rec + 4 * f2, (uint32_t)(gx - cx));                // dir_VEC (:672)
                sts_u8(rec + 4 * f2 + 1, (uint32_t)(gy - cy));
            }
        } else {
            // role 0: all droplets ascending; 1: own goal (v0_1 only, meda.py:808-816); 2: goals of the observed
            // others in python-set order, clipped; 3: idle lane.  v0_1 layers = roles; v0_2 has no own-goal layer.
            const int role = (VER == MEDA_OBS_V02) ? 2 * l : l;
            uint32_t observed = 0;                                     // droplets with a cell inside the window (:858-866)
            for (int j = 0; j < A; ++j) {
                const int dx = (int)(w[j] & 255u) - cx, dy = (int)((w[j] >> 8) & 255u) - cy;
                observed |= (uint32_t)(abs(dx) <= hf + kRad && abs(dy) <= hf + kRad) << j;
            }
            const uint8_t* order = (set_order && role == 2) ? set_order + (size_t)observed * A : nullptr;
            for (int k = 0; k < A; ++k) {
                int j = order ? (int)order[k] : k;                     // 0xFF terminates the set
                if (role == 1) j = i;
                const bool act = live && (role == 0 || (role == 1 && k == 0) ||
                                          (role == 2 && j < A && ((observed >> j) & 1u) && j != i));
                const uint32_t d = w[min(j, A - 1)];
                const int X = role ? (d >> 16) & 255u : d & 255u, Y = role ? d >> 24 : (d >> 8) & 255u;
                fill_rect(layer, fov, foot_rect(fov, ox, oy, X, Y, role == 2), j + 1, act);
            }
        }
    }
    if (VER != MEDA_OBS_BASE) {
        // border layer (v0_2 layer 2, :880-891; v0_1 layer 3, :826-839): x-derived bounds on the ROW axis with `width`,
        // y-derived on the column axis with `length`
        constexpr int kBorder = VER == MEDA_OBS_V02 ? 2 : 3;
        const int n_b = n_env * A * 2;
        for (int p0 = 0; p0 < n_b; p0 += 32) {
            const int p = p0 + lane;
            const bool valid = p < n_b;
            const int kind = p & 1;
            const int g = valid ? p >> 1 : 0;
            const int e = g / A, i = g - e * A;
            const bool live = valid && (flags[e] & kEnvSelected);
            const uint32_t me = words[e * A + i];
            const int cx = me & 255u, cy = (me >> 8) & 255u, gx = (me >> 16) & 255u, gy = me >> 24;
            const uint32_t rec = tile + (uint32_t)(g * D);
            const uint32_t lay = rec + kBorder * f2;
            if (kind == 0) {
                const int lb = hf - cx, rb = hf - (cfg.width - 1 - cx);
                int r_lo = 0, r_hi = 0;                                // rows [lo, hi) set to 1
                if (lb > 0) { r_lo = 0; r_hi = min(lb, fov); } else if (rb > 0) { r_lo = max(fov - rb, 0); r_hi = fov; }
                int b = r_lo * fov;                                    // whole rows are one contiguous byte range
                const int b1 = live ? r_hi * fov : b;
                for (; b < b1 && ((lay + b) & 3u); ++b) sts_u8(lay + b, 1u);
                for (; b + 4 <= b1; b += 4) sts_u32(lay + b, 0x01010101u);
                for (; b < b1; ++b) sts_u8(lay + b, 1u);
            } else {
                const int ub = hf - cy, db = hf - (cfg.length - 1 - cy);
                int q_lo = 0, q_hi = 0;                                // columns [lo, hi) set to 1
                if (ub > 0) { q_lo = 0; q_hi = min(ub, fov); } else if (db > 0) { q_lo = max(fov - db, 0); q_hi = fov; }
                const int nq = live ? q_hi - q_lo : 0;
                if (nq > 0) {
                    for (int r = 0; r < fov; ++r) {
                        const uint32_t q = lay + (uint32_t)(r * fov + q_lo);
#pragma unroll
                        for (int k = 0; k < 8; ++k)
                            if (k < nq) sts_u8(q + k, 1u);
                        for (int k = 8; k < nq; ++k) sts_u8(q + k, 1u);
                    }
                }
                if (live && VER == MEDA_OBS_V02) {                     // direction vector (:895)
                    sts_u8(rec + 3 * f2, (uint32_t)cfg.dir_y[gy - cy + cfg.width - 1]);
                    sts_u8(rec + 3 * f2 + 1, (uint32_t)cfg.dir_x[gx - cx + cfg.length - 1]);
                }
                if (live && VER == MEDA_OBS_V01) {                     // (:840) numerators of (dy / width, dx / length)
                    sts_u8(rec + 4 * f2, (uint32_t)(gy - cy));
                    sts_u8(rec + 4 * f2 + 1, (uint32_t)(gx - cx));
                }
            }
        }
    }
}

// Warp-level counterpart of store_tile(): the calling warp's finished tile -> global memory.  `tile` has the same
// 16-byte phase as `gdst` (the kernel places it so), hence everything between the first and the last 16-byte boundary
// of the destination leaves as ONE TMA bulk store issued by lane 0, whatever the alignment of the destination; the
// < 16 head and tail bytes are ordinary stores.  Returns true when a bulk store is in flight: lane 0 must
// tma_store_wait_read_all() before the tile is written again or the CTA exits.
template <bool PHASED>
__device__ __forceinline__ bool store_tile_warp(int8_t* __restrict__ gdst, const int8_t* tile, uint32_t nbytes)
{
    const uint32_t lane = threadIdx.x & 31u;
    // PHASED = false: the host saw that every warp's destination is 16-byte aligned (no head)
    const uint32_t head = PHASED ? min((16u - (uint32_t)(reinterpret_cast<uintptr_t>(gdst) & 15u)) & 15u, nbytes) : 0u;
    const uint32_t bulk = (nbytes - head) & ~15u;
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0 && bulk) {
        tma_store_1d(gdst + head, tile + head, bulk);
        tma_store_commit();
    }
    if (PHASED && lane < head) gdst[lane] = tile[lane];
    for (uint32_t b = head + bulk + lane; b < nbytes; b += 32u) gdst[b] = tile[b];
    return bulk != 0;
}

// Folds the usage log of env n into its counters: every logged droplet centre stands for its 5x5 footprint
// (addUsage, meda.py:591-598).  Threads tid, tid + nthreads, ... cooperate; the caller synchronises and clears the length.
__device__ __forceinline__ void meda_replay_usage_log(const meda_cfg_t& cfg, const meda_state_t& st, int64_t n, int tid,
                                                      int nthreads)
{
    const int A = cfg.n_agents, Lc = cfg.length;
    const int len = min(st.usage_log_len[n], st.usage_log_cap);
    const uint16_t* log = st.usage_log + (size_t)n * st.usage_log_cap * A;
    uint32_t* usage = st.usage + (size_t)n * cfg.width * Lc;
    for (int k = tid; k < len * A; k += nthreads) {
        const uint32_t c = log[k];
        if (c == 0xFFFFu) continue;
        uint32_t* p = usage + ((int)(c >> 8) - kRad) * Lc + ((int)(c & 255u) - kRad);
        for (int dy = 0; dy <= 2 * kRad; ++dy)
            for (int dx = 0; dx <= 2 * kRad; ++dx) atomicAdd(p + dy * Lc + dx, 1u);
    }
}

// Words of the optional degraded-cell bit map per env (meda_state_t.health_bits): bit y*length + x
__host__ __device__ __forceinline__ int meda_bit_words(const meda_cfg_t& cfg) { return (cfg.width * cfg.length + 31) >> 5; }

// updateHealth (meda.py:600-605) of env n: cells with usage > 50 get health *= degrade and usage = 0.  Threads tid,
// tid + nthreads, ... cooperate; four cells per load and a batch of loads in flight per thread, because inside a
// step launch (auto-reset) the dependent DRAM round trips of this scan are on the critical path.
__device__ __forceinline__ void meda_update_health_env(const meda_cfg_t& cfg, const meda_state_t& st, int64_t n, int tid,
                                                       int nthreads)
{
    const int cells = cfg.width * cfg.length;
    uint32_t* usage = st.usage + (size_t)n * cells;
    double* health = st.health + (size_t)n * cells;
    const double* degrade = st.degrade ? st.degrade + (size_t)n * cells : nullptr;
    uint32_t* bits = st.health_bits ? st.health_bits + (size_t)n * meda_bit_words(cfg) : nullptr;
    auto hit = [&](int k) {
        const double h = health[k] * (degrade ? degrade[k] : 1.0);
        health[k] = h;
        if (bits && h != 1.0) atomicOr(bits + (k >> 5), 1u << (k & 31));
        usage[k] = 0;
    };
    int done = 0;
    if ((reinterpret_cast<uintptr_t>(usage) & 15u) == 0) {
        const uint4* u4 = reinterpret_cast<const uint4*>(usage);
        const int n4 = cells >> 2;
        constexpr int kBatch = 4;
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

// updateHealth for the flagged envs of the tile
__device__ __forceinline__ void meda_update_health(const meda_cfg_t& cfg, const meda_state_t& st, const MedaSmem& S,
                                                   int64_t n0, int e_valid)
{
    if (!cfg.b_degrade || !st.usage || !st.health) return;
    for (int e = 0; e < e_valid; ++e)
        if (S.flag[e] & kEnvSelected) meda_update_health_env(cfg, st, n0 + e, (int)threadIdx.x, (int)blockDim.x);
}

// refresh/addTask/_genLegalDroplet (meda.py:161-185,213-233): centres uniform in [r, dim-r-1]; a droplet
// (destination) is redrawn while its centre is closer than 1.5*(2+2+2) = 9 to an earlier droplet (destination);
// the destination is also redrawn while it overlaps its own droplet.  One thread per env (sequential by nature).
// Droplets that cannot be placed make the reference loop for ever (meda.py:213-233); here the generator gives up after
// kMaxSamplerRounds redraws of one droplet, restores `words` from `prev` (the env's current layout) and returns false.
__device__ bool meda_generate_tasks(const meda_cfg_t& cfg, uint64_t seed, int64_t env, uint32_t episode, uint32_t* words,
                                    const uint32_t* prev)
{
    const int A = cfg.n_agents, W = cfg.width, Lc = cfg.length;
    uint64_t state = seed ^ (0x9E3779B97F4A7C15ull * (uint64_t)(kStreamLayout + 1));
    state += (uint64_t)env * 0xD1342543DE82EF95ull + ((uint64_t)episode << 32) * 0xDA942042E4DD58B5ull;
    state = mix64(state);
    for (int i = 0; i < A; ++i) {
        uint32_t sx, sy, tx, ty;
        for (uint32_t rounds = 0;; ++rounds) {
            if (rounds >= kMaxSamplerRounds) {                    // droplets that cannot be placed
                for (int j = 0; j < A; ++j) words[j] = prev[j];
                return false;
            }
            const uint64_t z = mix64(state += 0x9E3779B97F4A7C15ull);
            sy = kRad + __umulhi((uint32_t)z, (uint32_t)(W - 2 * kRad));
            sx = kRad + __umulhi((uint32_t)(z >> 32), (uint32_t)(Lc - 2 * kRad));
            bool ok = true;
            for (int j = 0; j < i; ++j) {
                const int dx = (int)sx - (int)(words[j] & 255u), dy = (int)sy - (int)((words[j] >> 8) & 255u);
                if (dx * dx + dy * dy < 81) { ok = false; break; }
            }
            if (ok) break;
        }
        for (uint32_t rounds = 0;; ++rounds) {
            if (rounds >= kMaxSamplerRounds) {
                for (int j = 0; j < A; ++j) words[j] = prev[j];
                return false;
            }
            const uint64_t z = mix64(state += 0x9E3779B97F4A7C15ull);
            ty = kRad + __umulhi((uint32_t)z, (uint32_t)(W - 2 * kRad));
            tx = kRad + __umulhi((uint32_t)(z >> 32), (uint32_t)(Lc - 2 * kRad));
            bool ok = true;
            for (int j = 0; j < i; ++j) {
                const int dx = (int)tx - (int)((words[j] >> 16) & 255u), dy = (int)ty - (int)(words[j] >> 24);
                if (dx * dx + dy * dy < 81) { ok = false; break; }
            }
            if (!ok) continue;
            if (abs((int)tx - (int)sx) <= 2 * kRad && abs((int)ty - (int)sy) <= 2 * kRad) continue;  // isDropletOverlap
            break;
        }
        words[i] = sx | (sy << 8) | (tx << 16) | (ty << 24);
    }
    return true;
}

// DMFB_STEP_AUTO_RESET inside the step launch: MEDAEnv.reset() (meda.py:541-550) of env n by the warp that just stepped
// it.  Lane `src` draws the task (sequential by nature) into `words` (the env's slice of the warp's shared droplet words,
// which the paint reads next) and resets the env's scalars; with degradation all 32 lanes then replay the usage log and
// run updateHealth.  All lanes call with uniform n / src.
__device__ __forceinline__ void meda_reset_env_warp(const meda_cfg_t& cfg, const meda_state_t& st, int64_t n, uint64_t seed,
                                                    uint32_t* words, int src)
{
    const int lane = threadIdx.x & 31, A = cfg.n_agents;
    if (lane == src) {
        const uint32_t episode = st.episode ? st.episode[n] + 1u : 0u;
        if (st.episode) st.episode[n] = episode;
        uint32_t* gdrop = reinterpret_cast<uint32_t*>(st.drop) + (size_t)n * A;   // holds the layout after the step
        if (!meda_generate_tasks(cfg, seed, cfg.env_base + n, episode, words, gdrop) && st.gen_status)
            atomicOr(st.gen_status, DMFB_STATUS_SAMPLER_GAVE_UP);
        for (int i = 0; i < A; ++i) {
            gdrop[i] = words[i];
            st.status[(size_t)n * A + i] = 0;
            if (st.start) reinterpret_cast<uint16_t*>(st.start)[(size_t)n * A + i] = (uint16_t)(words[i] & 0xFFFFu);
        }
        st.step_count[n] = 0;
        st.fails[n] = 0;
        st.terminated[n] = 0;
    }
    if (st.usage && st.usage_log != nullptr && st.usage_log_len != nullptr) {   // m_usage is about to be read
        meda_replay_usage_log(cfg, st, n, lane, 32);
        __syncwarp();
        if (lane == 0) st.usage_log_len[n] = 0;
    }
    if (cfg.b_degrade && st.usage && st.health) meda_update_health_env(cfg, st, n, lane, 32);   // (meda.py:600-605)
    __syncwarp();
}

// Shared memory of the step kernel: per warp a tile of EW envs' observations, their packed droplet words and
// per-env flags.
struct StepLayout {
    int EW, WPC;
    uint32_t warp_tile, off_word, off_flag, total;
    __host__ __device__ StepLayout(const meda_cfg_t& c, int EW_, int WPC_) {
        EW = EW_; WPC = WPC_;
        warp_tile = (((uint32_t)(EW * c.n_agents * c.obs_dim) + 15u) & ~15u) + 16u;   // + room for the store's phase
        uint32_t o = warp_tile * (uint32_t)WPC;
        off_word = o; o += (uint32_t)(WPC * EW * c.n_agents) * 4u;
        off_flag = o; o += ((uint32_t)(WPC * EW) + 3u) & ~3u;
        total = (o + 15u) & ~15u;
    }
};

// Per-lane (= per-droplet) inputs of one step, loaded one group ahead of their use.
struct DropIn {
    uint32_t d, status;
    int a, fails0, sc0, log_len;
    bool frozen;
};

__device__ __forceinline__ DropIn meda_load_inputs(const meda_state_t& st, const void* __restrict__ actions, int aes,
                                                   uint32_t flags, int64_t n0, int EW, int A, int cfg_cells)
{
    const int lane = threadIdx.x & 31;
    DropIn in;
    in.d = 0; in.status = 1; in.a = 8; in.fails0 = 0; in.sc0 = 0; in.log_len = 0; in.frozen = false;
    const int ev = (int)min((int64_t)EW, (int64_t)st.n_envs - n0);
    if (lane < ev * A) {
        const int e = lane / A;
        const int64_t n = n0 + e;
        const size_t ja = (size_t)n0 * A + lane;
        in.d = reinterpret_cast<const uint32_t*>(st.drop)[ja];
        in.status = st.status[ja];
        in.a = load_action(actions, aes, ja);
        in.fails0 = st.fails[n];
        in.sc0 = st.step_count[n];
        in.frozen = (flags & DMFB_STEP_FREEZE_TERM) && st.terminated[n];
        if (st.usage_log_len) in.log_len = st.usage_log_len[n];
    }
    if (st.health_bits != nullptr && st.health != nullptr) {
        // the bits under a droplet lie somewhere in the env's degraded-cell map (a few 128-byte lines): ask for those
        // lines together with the inputs, so that the lookups - whose addresses need the droplet word - find them in L1
        const int hb_bytes = ((cfg_cells + 31) >> 5) * 4;
        const char* hb = reinterpret_cast<const char*>(st.health_bits) + (size_t)n0 * hb_bytes;
        for (int k = lane * 128; k < ev * hb_bytes; k += 32 * 128) asm volatile("prefetch.global.L1 [%0];" :: "l"(hb + k));
    }
    return in;
}

// MEDAEnv.step (meda.py:513-539).  Warp w of the grid takes group w of EW consecutive envs (EW * A <= 32).
// (A persistent variant - grid capped at the resident CTAs, every warp looping over groups with its next inputs
// prefetched and the TMA drain overlapped - was measured slower, 115 vs 100 us, and removed.)
template <int VER, int A_T, int FOV_T, bool PHASED>
__global__ void __launch_bounds__(kThreads)
meda_step_kernel(const __grid_constant__ meda_cfg_t cfg, const meda_state_t st, const void* __restrict__ actions, int aes,
                 const double* __restrict__ u, uint64_t seed, uint32_t flags, const uint8_t* __restrict__ set_order,
                 const meda_out_t out, int EW, int n_groups)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int A = A_T ? A_T : cfg.n_agents, W = cfg.width, Lc = cfg.length, D = cfg.obs_dim;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpc = blockDim.x >> 5;
    const StepLayout L(cfg, EW, wpc);
    int8_t* const region = reinterpret_cast<int8_t*>(smem_raw + (size_t)warp * L.warp_tile);
    uint32_t* const s_word = reinterpret_cast<uint32_t*>(smem_raw + L.off_word) + warp * EW * A;
    uint8_t* const s_flag = smem_raw + L.off_flag + warp * EW;
    const int cells = W * Lc;

    const int grp = blockIdx.x * wpc + warp;
    if (grp >= n_groups) return;                              // no CTA-wide barrier below
    {
        // loads first (lane = droplet), so that their latency overlaps the zero fill of the tile.  (Programmatic
        // dependent launch - zero fill before griddepcontrol.wait, loads after it - was measured and dropped: the
        // exposed load latency costs more than the overlapped prologue saves, 63.1 -> 64.4 us as 4 sub-batches.)
        const DropIn in = meda_load_inputs(st, actions, aes, flags, (int64_t)grp * EW, EW, A, cells);
        {
            uint4* t4 = reinterpret_cast<uint4*>(region);
            const uint4 z = make_uint4(0u, 0u, 0u, 0u);
            for (int k = lane; k < (int)(L.warp_tile >> 4); k += 32) t4[k] = z;
        }
        const int64_t n0 = (int64_t)grp * EW;
        // the tile starts at the 16-byte phase of its destination, so that the middle of it can leave by TMA
        int8_t* const gobs = out.obs + (size_t)n0 * A * D;
        int8_t* const tile = PHASED ? region + (reinterpret_cast<uintptr_t>(gobs) & 15u) : region;
        const int ev = (int)min((int64_t)EW, (int64_t)st.n_envs - n0);
        const bool mine = lane < ev * A;
        const int e = mine ? lane / A : 0, i = lane - e * A;
        const int eb = e * A;                                 // first lane of this droplet's env
        const int64_t n = n0 + e;
        const size_t ja = (size_t)n * A + i;
        const uint32_t d = in.d;
        uint32_t status = in.status;
        const int a = in.a, fails0 = in.fails0, sc0 = in.sc0;
        const bool frozen = in.frozen;

        // ---- moveOneDroplet (meda.py:261-292): droplets are independent --------------------------------------
        int xc = d & 255u, yc = (d >> 8) & 255u;
        const int gx = (d >> 16) & 255u, gy = d >> 24;
        uint32_t code = 0;                                    // 0: 0.0, 1: -0.2, 2: -0.08, 3: -0.4
        if (mine && !frozen && !status) {                     // sticky status: reward 0, nothing moves (:248-249)
            const int old2 = (xc - gx) * (xc - gx) + (yc - gy) * (yc - gy);
            if (old2 < 16) {                                  // distance < r_i + r_goal = 4: snap onto the goal (:272-277)
                xc = gx; yc = gy; status = 1;
            } else {
                bool move = true;
                if (st.health) {                              // getMoveProb (:302-309): sequential float64 mean of 25 cells
                    const double* h = st.health + (size_t)n * cells;
                    // the (L2-sized) degraded-cell bit map first: five clear bits in each of the five rows mean 25
                    // cells of exactly 1.0, whose sequential sum is 25.0 and whose mean is 1.0 - no gather
                    uint32_t any_degraded = 1u;
                    if (st.health_bits) {
                        const int nw = meda_bit_words(cfg);
                        const uint32_t* hb = st.health_bits + (size_t)n * nw;
                        any_degraded = 0u;
#pragma unroll
                        for (int dy = -kRad; dy <= kRad; ++dy) {
                            const int k0 = (yc + dy) * Lc + xc - kRad, w0 = k0 >> 5;
                            any_degraded |= __funnelshift_r(hb[w0], hb[min(w0 + 1, nw - 1)], k0 & 31) & 31u;
                        }
                    }
                    double prob = 0.0;
                    if (any_degraded) {
                        for (int y = yc - kRad; y <= yc + kRad; ++y)
                            for (int x = xc - kRad; x <= xc + kRad; ++x) prob += h[y * Lc + x];
                        prob = prob / 25.0;
                    } else {
                        prob = 1.0;
                    }
                    double draw;
                    if (u) draw = u[ja];
                    else {
                        const uint32_t episode = st.episode ? st.episode[n] : 0u;
                        const uint4 r = env_random(seed, kStreamMove, cfg.env_base + n, episode, (uint32_t)sc0 + 1u,
                                                   (uint32_t)i);
                        draw = u53(r.x, r.y);
                    }
                    move = draw <= prob;                      // random.random() <= prob (:280)
                }
                if (move) meda_move(xc, yc, a, W, Lc);
                const int new2 = (xc - gx) * (xc - gx) + (yc - gy) * (yc - gy);
                // (:283-290) all comparisons of the float distances are exact on the integer squares
                code = (new2 < 16) ? 0u : (new2 == old2 && a == 8) ? 1u : (new2 < old2) ? 2u : 3u;
            }
        }
        const uint32_t word = (d & 0xFFFF0000u) | (uint32_t)xc | ((uint32_t)yc << 8);

        // ---- calPunish (:321-330): pairs closer than 1.5 * (r_i + r_j) = 6, exchanged by shuffles -------------
        int my_pun = 0, all = 1;
        for (int j = 0; j < A; ++j) {
            const uint32_t wj = __shfl_sync(0xFFFFFFFFu, word, (eb + j) & 31);
            const uint32_t sj = __shfl_sync(0xFFFFFFFFu, status, (eb + j) & 31);
            const int dx = xc - (int)(wj & 255u), dy = yc - (int)((wj >> 8) & 255u);
            my_pun += (j != i) & (dx * dx + dy * dy < 36);
            all &= (int)(sj & 1u);
        }
        int total = 0;
        for (int j = 0; j < A; ++j) total += __shfl_sync(0xFFFFFFFFu, my_pun, (eb + j) & 31);

        // ---- MEDAEnv.step bookkeeping (:521-538) -------------------------------------------------------------
        if (frozen) total = 0;
        const int fails = fails0 + total;                     // the reference keeps -0.6 * this count (:521)
        const int sc = sc0 + (frozen ? 0 : 1);
        double r = code == 0 ? 0.0 : code == 1 ? -0.2 : code == 2 ? -0.08 : -0.4;
        if (my_pun) {                                         // punish[i] -= 0.6, pun times; rewards[i] += punish[i]
            double pn = 0.0;
            for (int k = 0; k < my_pun; ++k) pn -= 0.6;
            r = r + pn;
        }
        if (all) {                                            // (:522-525)
            r = r + 3.0;
            if (fails == 0) r = r + 3.0;
        }
        if (frozen) r = 0.0;
        const bool in_time = !frozen && sc < cfg.max_step;    // (:529-537)
        const uint32_t done = in_time ? (status & 1u) : 1u;
        const float rf = (float)r;
        float team = 0.f;
        for (int j = 0; j < A; ++j) team += __shfl_sync(0xFFFFFFFFu, rf, (eb + j) & 31);

        if (mine) {
            if (out.reward) out.reward[ja] = rf;
            if (out.reward_f64) out.reward_f64[ja] = r;
            if (out.done) out.done[ja] = (uint8_t)done;
            if (!frozen) {
                reinterpret_cast<uint32_t*>(st.drop)[ja] = word;
                st.status[ja] = (uint8_t)(status & 1u);
            }
            if (out.avail) {
                uint8_t* av = out.avail + ja * cfg.n_actions;
                for (int k = 0; k < cfg.n_actions; ++k) av[k] = frozen ? 0 : 1;
            }
            if (i == 0) {
                const int term = in_time ? all : 1;
                if (out.constraints) out.constraints[n] = total;
                if (out.success) out.success[n] = (uint8_t)((in_time && all && fails == 0) ? 1 : 0);
                if (out.terminated) out.terminated[n] = (uint8_t)term;
                if (out.padded) out.padded[n] = (uint8_t)frozen;
                if (out.team_reward) out.team_reward[n] = team / (float)A;
                if (!frozen) {
                    st.terminated[n] = (uint8_t)term;
                    st.fails[n] = fails;
                    st.step_count[n] = sc;
                    if (term && (flags & DMFB_STEP_AUTO_RESET) && st.reset_list != nullptr)
                        st.reset_list[atomicAdd(st.reset_count, 1)] = (int32_t)n;
                }
                s_flag[e] = frozen ? 0 : kEnvSelected;
            }
            s_word[lane] = word;
        }

        // ---- addUsage (:591-598): footprints of one env may overlap -> RED.ADD per cell; one instruction per
        //      droplet, lane = footprint cell, so that the 25 cells coalesce into the few sectors they share ------
        if (st.usage) {
            bool use = mine && in_time && !done;              // only while step_count < max_step, agents not done
            if (st.usage_log != nullptr && st.usage_log_len != nullptr && mine && !frozen && in.log_len < st.usage_log_cap) {
                // log the droplet instead (2 bytes); a reset or meda_flush_usage replays the footprint
                st.usage_log[((size_t)n * st.usage_log_cap + in.log_len) * A + i] =
                    use ? (uint16_t)(xc | (yc << 8)) : (uint16_t)0xFFFFu;
                if (i == 0) st.usage_log_len[n] = in.log_len + 1;
                use = false;
            }
            const int cell_off = (lane / 5 - kRad) * Lc + (lane % 5 - kRad);
            for (uint32_t m = __ballot_sync(0xFFFFFFFFu, use); m; m &= m - 1u) {
                const int src = __ffs(m) - 1;
                const uint32_t w = __shfl_sync(0xFFFFFFFFu, word, src);
                const int env_off = __shfl_sync(0xFFFFFFFFu, e, src) * cells;
                if (lane < 25)
                    atomicAdd(st.usage + (size_t)n0 * cells + env_off + (int)((w >> 8) & 255u) * Lc + (int)(w & 255u) + cell_off, 1u);
            }
        }
        __syncwarp();
        if ((flags & DMFB_STEP_AUTO_RESET) && st.reset_list == nullptr) {
            // fused auto-reset: the envs of this warp that just terminated start their next episode here; their rows
            // then show its first observation (reward / done / info stay those of the finished step)
            const int term_env = (mine && i == 0 && !frozen && (in_time ? all : 1)) ? 1 : 0;
            for (uint32_t m = __ballot_sync(0xFFFFFFFFu, term_env); m; m &= m - 1u) {
                const int src = __ffs(m) - 1;                 // first lane of the env
                meda_reset_env_warp(cfg, st, n0 + src / A, seed, s_word + src, src);
            }
        }
        meda_paint_warp<VER, A_T, FOV_T>(cfg, s_word, s_flag, tile, ev, set_order);
        if (store_tile_warp<PHASED>(gobs, tile, (uint32_t)(ev * A * D)) && lane == 0)
            tma_store_wait_read_all();                        // shared memory must outlive the bulk read
    }
}

// mode 0: reset, mode 1: restart (droplets back to their start cells, meda.py:170-173,552-561), mode 2: observe only
__global__ void __launch_bounds__(kThreads)
meda_reset_kernel(const __grid_constant__ meda_cfg_t cfg, const meda_state_t st, const uint8_t* __restrict__ mask, int mode,
                  int new_chip, const uint8_t* __restrict__ layouts, const double* __restrict__ degrade_in, uint64_t seed,
                  const uint8_t* __restrict__ set_order, int8_t* __restrict__ obs, int E)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const MedaLayout L(cfg, E);
    const MedaSmem S(smem_raw, L);
    const int A = L.A;
    const int64_t n0 = (int64_t)blockIdx.x * E;
    const int e_valid = (int)min((int64_t)E, (int64_t)st.n_envs - n0);
    const int cells = cfg.width * cfg.length;

    int sel = 0;
    if ((int)threadIdx.x < e_valid) sel = (mask == nullptr) || (mask[n0 + threadIdx.x] != 0);
    const int n_selected = __syncthreads_count(sel);
    if (n_selected == 0) return;
    if (obs != nullptr) {
        uint4* t4 = reinterpret_cast<uint4*>(S.tile);
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
        for (int k = threadIdx.x; k < (int)(L.tile_bytes >> 4); k += blockDim.x) t4[k] = z;
    }
    for (int e = threadIdx.x; e < e_valid; e += blockDim.x) {   // e_valid <= E <= blockDim.x
        const int64_t n = n0 + e;
        uint32_t* words = S.word + e * A;
        uint32_t* gdrop = reinterpret_cast<uint32_t*>(st.drop) + (size_t)n * A;
        S.flag[e] = sel ? kEnvSelected : 0;
        if (sel && mode == 0) {
            const uint32_t episode = st.episode ? st.episode[n] + 1u : 0u;
            if (st.episode) st.episode[n] = episode;
            if (layouts) {
                const uint32_t* lay = reinterpret_cast<const uint32_t*>(layouts) + (size_t)n * A;
                for (int i = 0; i < A; ++i) words[i] = lay[i];
            } else {
                if (!meda_generate_tasks(cfg, seed, cfg.env_base + n, episode, words, gdrop) && st.gen_status)
                    atomicOr(st.gen_status, DMFB_STATUS_SAMPLER_GAVE_UP);
            }
            for (int i = 0; i < A; ++i) {
                gdrop[i] = words[i];
                st.status[(size_t)n * A + i] = 0;
                if (st.start) reinterpret_cast<uint16_t*>(st.start)[(size_t)n * A + i] = (uint16_t)(words[i] & 0xFFFFu);
            }
            st.step_count[n] = 0;
            st.fails[n] = 0;
            st.terminated[n] = 0;
        } else if (sel && mode == 1) {
            for (int i = 0; i < A; ++i) {
                words[i] = (gdrop[i] & 0xFFFF0000u) | reinterpret_cast<const uint16_t*>(st.start)[(size_t)n * A + i];
                gdrop[i] = words[i];
                st.status[(size_t)n * A + i] = 0;
            }
            st.step_count[n] = 0;      // `fails` is NOT cleared by the reference's restart (meda.py:552-561)
            st.terminated[n] = 0;
        } else {
            for (int i = 0; i < A; ++i) words[i] = gdrop[i];
        }
    }
    __syncthreads();
    if (mode == 0) {
        if (new_chip) {  // MEDAEnv.__init__ (meda.py:494-504): health 1, usage 0, degradation factors drawn
            for (int e = 0; e < e_valid; ++e) {
                if (!S.flag[e]) continue;
                const int64_t n = n0 + e;
                const uint32_t episode = st.episode ? st.episode[n] : 0u;
                if (threadIdx.x == 0 && st.usage_log_len) st.usage_log_len[n] = 0;   // usage = 0: the log goes with it
                if (st.health_bits && st.health)
                    for (int k = threadIdx.x; k < meda_bit_words(cfg); k += blockDim.x)
                        st.health_bits[(size_t)n * meda_bit_words(cfg) + k] = 0u;
                for (int k = threadIdx.x; k < cells; k += blockDim.x) {
                    if (st.usage) st.usage[(size_t)n * cells + k] = 0;
                    if (st.health) st.health[(size_t)n * cells + k] = 1.0;
                    if (st.degrade) {
                        double dg = 1.0;
                        if (degrade_in) dg = degrade_in[(size_t)n * cells + k];
                        else if (cfg.b_degrade) {
                            const uint4 r = env_random(seed, kStreamDegrade, cfg.env_base + n, episode, (uint32_t)k, 0u);
                            dg = __dadd_rn(__dmul_rn(u53(r.x, r.y), 0.4), 0.6);   // rand * 0.4 + 0.6 as two roundings, like NumPy (no FMA)
                            if (u53(r.z, r.w) < 1.0 - cfg.per_degrade) dg = 1.0;
                        }
                        st.degrade[(size_t)n * cells + k] = dg;
                    }
                }
            }
        } else {
            if (st.usage && st.usage_log != nullptr && st.usage_log_len != nullptr) {   // m_usage is about to be read
                for (int e = 0; e < e_valid; ++e)
                    if (S.flag[e]) meda_replay_usage_log(cfg, st, n0 + e, (int)threadIdx.x, (int)blockDim.x);
                __syncthreads();
                for (int e = (int)threadIdx.x; e < e_valid; e += (int)blockDim.x)
                    if (S.flag[e]) st.usage_log_len[n0 + e] = 0;
            }
            meda_update_health(cfg, st, S, n0, e_valid);  // after the observation in the reference (:547-548); obs does not read health
        }
    }
    if (obs == nullptr) return;
    {
        // every warp paints a contiguous share of the tile's envs
        const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
        const int share = (e_valid + nwarps - 1) / nwarps, e0 = warp * share;
        const int ne = min(share, e_valid - e0);
        if (ne > 0) {
            int8_t* t = S.tile + (size_t)e0 * A * L.D;
            if (cfg.obs_version == MEDA_OBS_V02)
                meda_paint_warp<MEDA_OBS_V02, 0, 0>(cfg, S.word + e0 * A, S.flag + e0, t, ne, set_order);
            else if (cfg.obs_version == MEDA_OBS_V01)
                meda_paint_warp<MEDA_OBS_V01, 0, 0>(cfg, S.word + e0 * A, S.flag + e0, t, ne, set_order);
            else
                meda_paint_warp<MEDA_OBS_BASE, 0, 0>(cfg, S.word + e0 * A, S.flag + e0, t, ne, set_order);
        }
    }
    int8_t* gobs = obs + (size_t)n0 * A * L.D;
    if (n_selected == e_valid) store_tile(gobs, S.tile, (uint32_t)(e_valid * A * L.D));
    else store_rows_masked(gobs, S.tile, e_valid, A * L.D, S.flag);
}

__global__ void __launch_bounds__(128)
meda_flush_usage_kernel(const __grid_constant__ meda_cfg_t cfg, const meda_state_t st)
{
    const int64_t n = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);   // one warp per env
    if (n >= st.n_envs) return;
    meda_replay_usage_log(cfg, st, n, (int)(threadIdx.x & 31), 32);
    __syncwarp();
    if ((threadIdx.x & 31) == 0) st.usage_log_len[n] = 0;
}

// Rebuilds meda_state_t.health_bits from health (bit k of an env = health[k] != 1.0): one warp per env.
__global__ void __launch_bounds__(128)
meda_health_bits_kernel(const __grid_constant__ meda_cfg_t cfg, const meda_state_t st)
{
    const int64_t n = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (n >= st.n_envs) return;
    const int lane = threadIdx.x & 31, cells = cfg.width * cfg.length, nw = meda_bit_words(cfg);
    const double* health = st.health + (size_t)n * cells;
    uint32_t* bits = st.health_bits + (size_t)n * nw;
    for (int w = 0; w < nw; ++w) {
        const int k = w * 32 + lane;
        const unsigned m = __ballot_sync(0xFFFFFFFFu, k < cells && health[k] != 1.0);
        if (lane == 0) bits[w] = m;
    }
}

// DMFB_STEP_AUTO_RESET, second half: MEDAEnv.reset() (meda.py:541-550) for exactly the envs the step just appended to
// st.reset_list.  One small CTA per env, a few hundred envs per step, instead of a masked sweep over the whole batch:
// thread 0 draws the task, all threads replay the usage log and run updateHealth, warp 0 paints the first
// observation of the new episode and stores it with a (phase-matched) TMA bulk store.  The last CTA to finish empties
// the list.
template <int VER>
__global__ void __launch_bounds__(128)
meda_reset_list_kernel(const __grid_constant__ meda_cfg_t cfg, const meda_state_t st, uint64_t seed,
                       const uint8_t* __restrict__ set_order, int8_t* __restrict__ obs)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int A = cfg.n_agents, D = cfg.obs_dim;
    const int tid = (int)threadIdx.x, nthreads = (int)blockDim.x;
    const uint32_t region_bytes = (((uint32_t)(A * D) + 15u) & ~15u) + 16u;
    int8_t* const region = reinterpret_cast<int8_t*>(smem_raw);
    uint32_t* const s_word = reinterpret_cast<uint32_t*>(smem_raw + region_bytes);
    uint8_t* const s_flag = smem_raw + region_bytes + (size_t)A * 4;
    const int count = st.reset_count[0];
    for (int k = blockIdx.x; k < count; k += gridDim.x) {
        const int64_t n = st.reset_list[k];
        for (int q = tid; q < (int)(region_bytes >> 4); q += nthreads) reinterpret_cast<uint4*>(region)[q] = make_uint4(0u, 0u, 0u, 0u);
        if (tid == 0) {
            const uint32_t episode = st.episode ? st.episode[n] + 1u : 0u;
            if (st.episode) st.episode[n] = episode;
            uint32_t* gdrop = reinterpret_cast<uint32_t*>(st.drop) + (size_t)n * A;
            if (!meda_generate_tasks(cfg, seed, cfg.env_base + n, episode, s_word, gdrop) && st.gen_status)
                atomicOr(st.gen_status, DMFB_STATUS_SAMPLER_GAVE_UP);
            for (int i = 0; i < A; ++i) {
                gdrop[i] = s_word[i];
                st.status[(size_t)n * A + i] = 0;
                if (st.start) reinterpret_cast<uint16_t*>(st.start)[(size_t)n * A + i] = (uint16_t)(s_word[i] & 0xFFFFu);
            }
            st.step_count[n] = 0;
            st.fails[n] = 0;
            st.terminated[n] = 0;
            s_flag[0] = kEnvSelected;
        }
        if (st.usage && st.usage_log != nullptr && st.usage_log_len != nullptr) {   // m_usage is about to be read
            meda_replay_usage_log(cfg, st, n, tid, nthreads);
            __syncthreads();
            if (tid == 0) st.usage_log_len[n] = 0;
        }
        if (cfg.b_degrade && st.usage && st.health) meda_update_health_env(cfg, st, n, tid, nthreads);   // (meda.py:600-605)
        __syncthreads();                                          // task words and the zeroed tile before the paint
        if (tid < 32) {
            int8_t* const gobs = obs + (size_t)n * A * D;
            int8_t* const tile = region + (reinterpret_cast<uintptr_t>(gobs) & 15u);
            meda_paint_warp<VER, 0, 0>(cfg, s_word, s_flag, tile, 1, set_order);
            if (store_tile_warp<true>(gobs, tile, (uint32_t)(A * D)) && tid == 0) tma_store_wait_read_all();
        }
        __syncthreads();                                          // the tile is free again
    }
    // the last CTA out empties the list for the next step
    if (tid == 0) {
        __threadfence();
        if (atomicAdd(st.reset_count + 1, 1) == (int)gridDim.x - 1) {
            st.reset_count[0] = 0;
            st.reset_count[1] = 0;
        }
    }
}

int meda_launch_reset_list(const meda_cfg_t* cfg, const meda_state_t* st, uint64_t seed, const uint8_t* set_order, int8_t* obs,
                           void* stream)
{
    const uint32_t region = ((((uint32_t)(cfg->n_agents * cfg->obs_dim) + 15u) & ~15u) + 16u);
    const uint32_t smem = region + cfg->n_agents * 4 + 16;
    static thread_local int sms = 0;
    if (!sms) {
        int dev = 0;
        DMFB_CUDA_TRY(cudaGetDevice(&dev));
        DMFB_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    }
    int grid = st->n_envs;                                        // never more CTAs than envs
    if (grid > 8 * sms) grid = 8 * sms;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
#define MEDA_RESET_LIST(V)                                                                                         \
    {                                                                                                              \
        static thread_local uint32_t smem_set = 0;                                                                 \
        if (smem > 48 * 1024 && smem > smem_set) {                                                                 \
            DMFB_CUDA_TRY(cudaFuncSetAttribute(meda_reset_list_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            smem_set = smem;                                                                                       \
        }                                                                                                          \
        meda_reset_list_kernel<V><<<grid, 128, smem, s>>>(*cfg, *st, seed, set_order, obs);                        \
    }
    if (cfg->obs_version == MEDA_OBS_V02) MEDA_RESET_LIST(MEDA_OBS_V02)
    else if (cfg->obs_version == MEDA_OBS_V01) MEDA_RESET_LIST(MEDA_OBS_V01)
    else MEDA_RESET_LIST(MEDA_OBS_BASE)
#undef MEDA_RESET_LIST
    g_launches.fetch_add(1);
    DMFB_CUDA_TRY(cudaGetLastError());
    return DMFB_OK;
}

int meda_tile_envs(const meda_cfg_t& cfg)
{
    const int row = cfg.n_agents * cfg.obs_dim;
    return pick_tile_envs(row, 48 * 1024, 64);
}

int meda_check(const meda_cfg_t* cfg, const meda_state_t* st)
{
    if (cfg && st && st->n_envs == 0) return DMFB_OK;   // empty batch
    if (!cfg || !st || st->n_envs < 0 || !st->drop || !st->status || !st->step_count || !st->fails || !st->terminated) {
        snprintf(g_last_error, sizeof(g_last_error), "meda: null cfg/state pointer");
        return DMFB_ERR_BAD_ARG;
    }
    return DMFB_OK;
}

int meda_launch_reset(const meda_cfg_t* cfg, const meda_state_t* st, const uint8_t* mask, int mode, int new_chip,
                      const uint8_t* layouts, const double* degrade, uint64_t seed, const uint8_t* set_order, int8_t* obs,
                      void* stream)
{
    int rc = meda_check(cfg, st);
    if (rc) return rc;
    if (st->n_envs == 0) return DMFB_OK;
    const int E = meda_tile_envs(*cfg);
    const MedaLayout L(*cfg, E);
    if (L.total > 48 * 1024)
        DMFB_CUDA_TRY(cudaFuncSetAttribute(meda_reset_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
    const int grid = (st->n_envs + E - 1) / E;
    meda_reset_kernel<<<grid, kThreads, L.total, static_cast<cudaStream_t>(stream)>>>(*cfg, *st, mask, mode, new_chip,
                                                                                       layouts, degrade, seed, set_order,
                                                                                       obs, E);
    g_launches.fetch_add(1);
    DMFB_CUDA_TRY(cudaGetLastError());
    return DMFB_OK;
}

// Envs per warp of the step kernel (EW * A <= 32 droplet lanes).  Any count works with the phase-matched store; the
// default keeps the smallest count whose observation span is a multiple of 16 bytes, which measured best.
int meda_warp_envs(const meda_cfg_t& cfg)
{
    const int A = cfg.n_agents;
    static const int forced = getenv("MEDA_WARP_ENVS") ? atoi(getenv("MEDA_WARP_ENVS")) : 0;   // tuning knob
    if (forced >= 1 && forced * A <= 32) return forced;
    const int unit = 16 / gcd_int(16, A * cfg.obs_dim);
    if (unit * A <= 32) return unit;
    return 32 / A;
}

template <int VER, int A_T, int FOV_T>
int meda_launch_step_t(const meda_cfg_t* cfg, const meda_state_t* st, const void* actions, int aes, const double* u,
                       uint64_t seed, uint32_t flags, const uint8_t* set_order, const meda_out_t* out, void* stream)
{
    const int EW = meda_warp_envs(*cfg);
    int wpc = 1;   // small CTAs pack the shared memory of an SM best (measured: 1 warp < 2 < 4 < 8; base obs 66.3 -> 63.5 us)
    static const int forced_wpc = getenv("MEDA_WARPS_PER_CTA") ? atoi(getenv("MEDA_WARPS_PER_CTA")) : 0;   // tuning knob
    if (forced_wpc >= 1 && forced_wpc <= kThreads / 32) wpc = forced_wpc;
    while (wpc > 1 && StepLayout(*cfg, EW, wpc).total > 200u * 1024u) wpc >>= 1;
    const StepLayout L(*cfg, EW, wpc);
    // every warp's destination is 16-byte aligned when the tensor is and a group's span is a multiple of 16 bytes
    const bool phased = (reinterpret_cast<uintptr_t>(out->obs) & 15u) != 0 || (EW * cfg->n_agents * cfg->obs_dim) % 16 != 0;
    auto kern = phased ? meda_step_kernel<VER, A_T, FOV_T, true> : meda_step_kernel<VER, A_T, FOV_T, false>;
    static thread_local uint32_t smem_set[2] = {0, 0};        // per kernel instance (phased or not)
    if (L.total > 48 * 1024 && L.total > smem_set[phased]) {
        DMFB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
        smem_set[phased] = L.total;
    }
    const int n_groups = (st->n_envs + EW - 1) / EW;          // one group of EW envs per warp
    const int grid = (n_groups + wpc - 1) / wpc;
    kern<<<grid, wpc * 32, L.total, static_cast<cudaStream_t>(stream)>>>(*cfg, *st, actions, aes, u, seed, flags, set_order,
                                                                         *out, EW, n_groups);
    g_launches.fetch_add(1);
    DMFB_CUDA_TRY(cudaGetLastError());
    return DMFB_OK;
}

int meda_launch_step(const meda_cfg_t* cfg, const meda_state_t* st, const void* actions, int aes, const double* u,
                     uint64_t seed, uint32_t flags, const uint8_t* set_order, const meda_out_t* out, void* stream)
{
#define MEDA_STEP_ARGS cfg, st, actions, aes, u, seed, flags, set_order, out, stream
    const bool c4 = cfg->n_agents == 4 && cfg->fov == 19;
    switch (cfg->obs_version) {
    case MEDA_OBS_V02:
        return c4 ? meda_launch_step_t<MEDA_OBS_V02, 4, 19>(MEDA_STEP_ARGS) : meda_launch_step_t<MEDA_OBS_V02, 0, 0>(MEDA_STEP_ARGS);
    case MEDA_OBS_V01:
        return c4 ? meda_launch_step_t<MEDA_OBS_V01, 4, 19>(MEDA_STEP_ARGS) : meda_launch_step_t<MEDA_OBS_V01, 0, 0>(MEDA_STEP_ARGS);
    default:
        return c4 ? meda_launch_step_t<MEDA_OBS_BASE, 4, 19>(MEDA_STEP_ARGS) : meda_launch_step_t<MEDA_OBS_BASE, 0, 0>(MEDA_STEP_ARGS);
    }
#undef MEDA_STEP_ARGS
}

}  // namespace
}  // namespace dmfb

using namespace dmfb;

extern "C" {

int meda_cfg_init(meda_cfg_t* cfg, int width, int length, int n_agents, int fov, int b_degrade, double per_degrade,
                  int obs_version)
{
    if (!cfg) return DMFB_ERR_BAD_ARG;
    memset(cfg, 0, sizeof(*cfg));
    if (width <= 0 || length <= 0 || n_agents <= 0) return DMFB_ERR_BAD_ARG;                 // meda.py:472-473
    if (n_agents > (width / 15) * (length / 15)) return DMFB_ERR_TOO_MANY_DROPLETS;        // n_limit, meda.py:151-154
    if (width > DMFB_MAX_DIM || length > DMFB_MAX_DIM || n_agents > DMFB_MAX_AGENTS || fov < 5 || fov > 2 * DMFB_MAX_FOV)
        return DMFB_ERR_BAD_ARG;
    if (obs_version != MEDA_OBS_BASE && obs_version != MEDA_OBS_V01 && obs_version != MEDA_OBS_V02) return DMFB_ERR_BAD_ARG;
    cfg->width = width; cfg->length = length; cfg->n_agents = n_agents; cfg->fov = fov;
    cfg->b_degrade = b_degrade ? 1 : 0; cfg->per_degrade = per_degrade; cfg->obs_version = obs_version;
    cfg->max_step = width + length;                                                        // meda.py:492
    cfg->n_actions = 9;
    cfg->obs_dim = (obs_version == MEDA_OBS_V02 ? 3 : 4) * fov * fov + 2;
    cfg->radius = kRad;
    // v0_2 direction table (meda.py:895): round(d / (dim / 30)), python round == rint on the float64 quotient
    for (int d = -(width - 1); d <= width - 1; ++d)
        cfg->dir_y[d + width - 1] = (int8_t)(int)__builtin_rint((double)d / ((double)width / 30.0));
    for (int d = -(length - 1); d <= length - 1; ++d)
        cfg->dir_x[d + length - 1] = (int8_t)(int)__builtin_rint((double)d / ((double)length / 30.0));
    return DMFB_OK;
}

int meda_step(const meda_cfg_t* cfg, const meda_state_t* state, const void* actions, int action_elem_size,
              const double* u_inject, uint64_t seed, uint32_t flags, const uint8_t* set_order, const meda_out_t* out,
              void* stream)
{
    int rc = meda_check(cfg, state);
    if (rc) return rc;
    if (state->n_envs == 0) return DMFB_OK;
    if (!actions || !out || !out->obs || (action_elem_size != 1 && action_elem_size != 4 && action_elem_size != 8)) {
        snprintf(g_last_error, sizeof(g_last_error), "meda_step: bad actions/out");
        return DMFB_ERR_BAD_ARG;
    }
    if (cfg->obs_version != MEDA_OBS_BASE && cfg->n_agents > 8 && !set_order) {
        snprintf(g_last_error, sizeof(g_last_error), "meda_step: v0_1 / v0_2 obs with more than 8 agents needs set_order");
        return DMFB_ERR_BAD_ARG;
    }
    if (state->n_envs == 0) return DMFB_OK;
    rc = meda_launch_step(cfg, state, actions, action_elem_size, u_inject, seed, flags, set_order, out, stream);
    if (rc) return rc;
    // DMFB_STEP_AUTO_RESET: fused into the step kernel, unless the caller supplied the reset list: then the step only
    // lists the envs that terminated and a second, small kernel resets exactly those
    if ((flags & DMFB_STEP_AUTO_RESET) && state->reset_list && state->reset_count)
        return meda_launch_reset_list(cfg, state, seed, set_order, out->obs, stream);
    return DMFB_OK;
}

int meda_reset(const meda_cfg_t* cfg, const meda_state_t* state, const uint8_t* mask, int new_chip, const uint8_t* layouts,
               const double* degrade, uint64_t seed, const uint8_t* set_order, int8_t* obs, void* stream)
{
    return meda_launch_reset(cfg, state, mask, 0, new_chip, layouts, degrade, seed, set_order, obs, stream);
}

int meda_restart(const meda_cfg_t* cfg, const meda_state_t* state, const uint8_t* mask, const uint8_t* set_order,
                 int8_t* obs, void* stream)
{
    if (state && state->n_envs != 0 && !state->start) {
        snprintf(g_last_error, sizeof(g_last_error), "meda_restart needs state->start");
        return DMFB_ERR_BAD_ARG;
    }
    return meda_launch_reset(cfg, state, mask, 1, 0, nullptr, nullptr, 0, set_order, obs, stream);
}

int meda_observe(const meda_cfg_t* cfg, const meda_state_t* state, const uint8_t* set_order, int8_t* obs, void* stream)
{
    if (state && state->n_envs == 0) return DMFB_OK;
    if (!obs) return DMFB_ERR_BAD_ARG;
    return meda_launch_reset(cfg, state, nullptr, 2, 0, nullptr, nullptr, 0, set_order, obs, stream);
}

int meda_flush_usage(const meda_cfg_t* cfg, const meda_state_t* state, void* stream)
{
    int rc = meda_check(cfg, state);
    if (rc) return rc;
    if (state->n_envs == 0 || !state->usage || !state->usage_log || !state->usage_log_len) return DMFB_OK;
    meda_flush_usage_kernel<<<(state->n_envs + 3) / 4, 128, 0, static_cast<cudaStream_t>(stream)>>>(*cfg, *state);
    g_launches.fetch_add(1);
    DMFB_CUDA_TRY(cudaGetLastError());
    return DMFB_OK;
}

int meda_sync_health_bits(const meda_cfg_t* cfg, const meda_state_t* state, void* stream)
{
    int rc = meda_check(cfg, state);
    if (rc) return rc;
    if (state->n_envs == 0 || !state->health || !state->health_bits) return DMFB_OK;
    meda_health_bits_kernel<<<(state->n_envs + 3) / 4, 128, 0, static_cast<cudaStream_t>(stream)>>>(*cfg, *state);
    g_launches.fetch_add(1);
    DMFB_CUDA_TRY(cudaGetLastError());
    return DMFB_OK;
}

// Iteration order of a CPython set holding the small ints of `mask_bits`, inserted in ascending order
// (setobject.c: 8-slot table that grows to the next power of two > 4*used once fill*5 >= mask*3; hash(i) = i;
// 9 linear probes only where i+9 <= mask, then i = (5i + 1 + perturb) & mask).  MEDAEnv_v0_2 iterates such a
// set (meda.py:862-872).  out[k] = k-th element, 0xFF-terminated, `n_max` entries.
int meda_set_order(uint32_t mask_bits, int n_max, uint8_t* out)
{
    if (!out || n_max < 1 || n_max > 32) return DMFB_ERR_BAD_ARG;
    int table[256];
    int size = 8, used = 0;
    for (int k = 0; k < 256; ++k) table[k] = -1;
    auto insert = [&](int v) {
        const size_t m = (size_t)size - 1;
        size_t perturb = (size_t)v, i = (size_t)v & m;
        for (;;) {
            const size_t lim = (i + 9 <= m) ? 9 : 0;
            for (size_t p = 0; p <= lim; ++p)
                if (table[i + p] < 0) { table[i + p] = v; return; }
            perturb >>= 5;
            i = (i * 5 + 1 + perturb) & m;
        }
    };
    for (int v = 0; v < n_max; ++v) {
        if (!((mask_bits >> v) & 1u)) continue;
        insert(v);
        ++used;
        if ((size_t)used * 5 >= ((size_t)size - 1) * 3) {
            int old[256];
            const int oldsize = size;
            memcpy(old, table, sizeof(old));
            int newsize = 8;
            while (newsize <= used * 4) newsize <<= 1;
            size = newsize;
            for (int k = 0; k < 256; ++k) table[k] = -1;
            for (int k = 0; k < oldsize; ++k)
                if (old[k] >= 0) insert(old[k]);
        }
    }
    int n = 0;
    for (int k = 0; k < size && n < n_max; ++k)
        if (table[k] >= 0) out[n++] = (uint8_t)table[k];
    while (n < n_max) out[n++] = 0xFF;
    return DMFB_OK;
}

}  // extern "C"
