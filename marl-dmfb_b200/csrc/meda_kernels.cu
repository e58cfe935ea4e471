// meda_kernels.cu — batched MEDA (micro-electrode-dot-array) environment step for sm_100a (B200).
//
// Replaces, for N independent chips at once, the reference call tree
//   MEDAEnv.step (env/MEDA/meda.py:513-539) -> RoutingTaskManager.moveDroplets (:241-259)
//   -> moveOneDroplet (:261-292) / getMoveProb (:302-309) / Droplet.move (:106-138) -> calPunish (:321-330)
//   -> getObs (:607-611) -> getOneObs (:613-674, or MEDAEnv_v0_2 :850-897) -> addUsage (:591-598)
// plus MEDAEnv.reset (:541-550) with refresh/addTask (:161-185) and updateHealth (:600-605).
//
// Design: one CTA (8 warps) per tile of E envs, E chosen so that the tile's observation span is a multiple
// of 16 bytes (one TMA bulk store per tile).  Droplets of a MEDA chip move independently (the reference has
// no collision prevention), so the dynamics are one thread per droplet; the pairwise punish counts and the
// per-env bookkeeping are one thread per env; the observation of an agent (4 or 3 layers of fov x fov cells
// filled from 5x5 footprints) is painted by ONE WARP PER AGENT, lane = footprint cell, layers in the
// reference's write order so that "later index overwrites" is preserved.
#include "common.cuh"

namespace dmfb {
namespace {

constexpr int kThreads = 256;
constexpr int kRad = 2;          // RoutingTaskManager.r (meda.py:150)
constexpr int kFootCells = 25;   // (2r+1)^2

struct MedaLayout {
    int E, A, D;
    uint32_t tile_bytes, off_word, off_misc, off_rew, off_envi, off_done, off_flag, off_dirx, off_diry, total;
    __host__ __device__ MedaLayout(const meda_cfg_t& c, int E_) {
        E = E_; A = c.n_agents; D = c.obs_dim;
        tile_bytes = ((uint32_t)(E * A * D) + 15u) & ~15u;
        uint32_t o = tile_bytes;
        off_word = o; o += (uint32_t)(E * A) * 4u;     // packed droplet words after the moves
        off_misc = o; o += (uint32_t)(E * A) * 4u;     // status | code<<8 per droplet
        off_rew = o; o += (uint32_t)(E * A) * 4u;      // float rewards (for the team mean)
        off_envi = o; o += (uint32_t)E * 8u;           // fails, step_count per env
        off_done = o; o += ((uint32_t)(E * A) + 3u) & ~3u;
        off_flag = o; o += ((uint32_t)E + 3u) & ~3u;   // per env flags
        off_dirx = o; o += ((uint32_t)(2 * c.length) + 3u) & ~3u;
        off_diry = o; o += ((uint32_t)(2 * c.width) + 3u) & ~3u;
        total = (o + 15u) & ~15u;
    }
};

constexpr uint8_t kEnvSelected = 1;   // reset: env selected / step: env live (paint its rows)
constexpr uint8_t kEnvUsage = 2;      // step: addUsage applies (step_count < max_step)
constexpr uint8_t kEnvFrozen = 4;     // padded step

struct MedaSmem {
    int8_t* tile;
    uint32_t* word;
    uint32_t* misc;
    float* rew;
    int32_t* envi;
    uint8_t* done;
    uint8_t* flag;
    int8_t* dirx;   // indexed d + length-1
    int8_t* diry;   // indexed d + width-1
    __device__ MedaSmem(unsigned char* base, const MedaLayout& L) {
        tile = reinterpret_cast<int8_t*>(base);
        word = reinterpret_cast<uint32_t*>(base + L.off_word);
        misc = reinterpret_cast<uint32_t*>(base + L.off_misc);
        rew = reinterpret_cast<float*>(base + L.off_rew);
        envi = reinterpret_cast<int32_t*>(base + L.off_envi);
        done = reinterpret_cast<uint8_t*>(base + L.off_done);
        flag = reinterpret_cast<uint8_t*>(base + L.off_flag);
        dirx = reinterpret_cast<int8_t*>(base + L.off_dirx);
        diry = reinterpret_cast<int8_t*>(base + L.off_diry);
    }
};

__device__ __forceinline__ void meda_prologue(const meda_cfg_t& cfg, const MedaLayout& L, const MedaSmem& S, bool zero)
{
    if (zero) {
        uint4* t4 = reinterpret_cast<uint4*>(S.tile);
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
        for (int k = threadIdx.x; k < (int)(L.tile_bytes >> 4); k += blockDim.x) t4[k] = z;
    }
    for (int k = threadIdx.x; k < 2 * cfg.length; k += blockDim.x) S.dirx[k] = cfg.dir_x[k];
    for (int k = threadIdx.x; k < 2 * cfg.width; k += blockDim.x) S.diry[k] = cfg.dir_y[k];
}

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

// One footprint pass of a warp: lane l < 25 owns cell (X-2 + l%5, Y-2 + l/5) and writes `val` into `layer`
// at its window position, either only when inside the window or clipped onto it.  Returns (warp-uniform)
// whether any cell fell inside the window.
__device__ __forceinline__ bool foot_pass(int8_t* layer, int fov, int ox, int oy, int X, int Y, int val, bool clip,
                                          bool enabled)
{
    const int lane = threadIdx.x & 31;
    const int lx = lane % 5, ly = lane / 5;
    int nx = X - kRad + lx - ox, ny = Y - kRad + ly - oy;
    const bool cell = enabled && lane < kFootCells;
    const bool inside = (unsigned)nx < (unsigned)fov && (unsigned)ny < (unsigned)fov;
    if (clip) {
        nx = min(max(nx, 0), fov - 1);
        ny = min(max(ny, 0), fov - 1);
    }
    if (cell && (inside || clip)) layer[ny * fov + nx] = (int8_t)val;
    return __any_sync(0xFFFFFFFFu, cell && inside);
}

// getOneObs of agent i of the env whose A packed words start at `words`, painted by one warp into `rec`.
__device__ void meda_paint_agent(const meda_cfg_t& cfg, const MedaSmem& S, const uint32_t* words, int i, int8_t* rec,
                                 const uint8_t* __restrict__ set_order)
{
    const int fov = cfg.fov, f2 = fov * fov, hf = fov >> 1, A = cfg.n_agents;
    const int lane = threadIdx.x & 31;
    const uint32_t me = words[i];
    const int cx = me & 255u, cy = (me >> 8) & 255u, gx = (me >> 16) & 255u, gy = me >> 24;
    const int ox = cx - hf, oy = cy - hf;
    if (cfg.obs_version == MEDA_OBS_BASE) {
        // MEDAEnv.getOneObs (meda.py:613-674)
        foot_pass(rec, fov, ox, oy, cx, cy, i + 1, false, true);                 // layer 0: own droplet
        foot_pass(rec + f2, fov, ox, oy, gx, gy, i + 1, false, true);            // layer 1: own goal
        for (int j = 0; j < A; ++j) {                                            // layer 2: other droplets, ascending
            const uint32_t d = words[j];
            foot_pass(rec + 2 * f2, fov, ox, oy, d & 255u, (d >> 8) & 255u, j + 1, false, j != i);
            __syncwarp();
        }
        for (int j = 0; j < A; ++j) {                                            // layer 3: ALL other goals, clipped
            const uint32_t d = words[j];
            foot_pass(rec + 3 * f2, fov, ox, oy, (d >> 16) & 255u, d >> 24, j + 1, true, j != i);
            __syncwarp();
        }
        if (lane == 0) {
            rec[4 * f2] = (int8_t)(gx - cx);                                     // dir_VEC (:672)
            rec[4 * f2 + 1] = (int8_t)(gy - cy);
        }
    } else {
        // MEDAEnv_v0_2.getOneObs (meda.py:850-897)
        uint32_t observed = 0;
        for (int j = 0; j < A; ++j) {                                            // layer 0: all droplets in the window
            const uint32_t d = words[j];
            if (foot_pass(rec, fov, ox, oy, d & 255u, (d >> 8) & 255u, j + 1, false, true)) observed |= 1u << j;
            __syncwarp();
        }
        // layer 1: goals of the observed others, clipped, in the iteration order of the python set (:871-878)
        const uint8_t* order = set_order ? set_order + (size_t)observed * A : nullptr;
        for (int k = 0; k < A; ++k) {
            const int j = order ? (int)order[k] : k;
            if (j >= A) break;                                                    // 0xFF terminator
            const bool take = ((observed >> j) & 1u) && j != i;
            const uint32_t d = words[j];
            foot_pass(rec + f2, fov, ox, oy, (d >> 16) & 255u, d >> 24, j + 1, true, take);
            __syncwarp();
        }
        // layer 2 (:880-891): x-derived bounds on the ROW axis with `width`, y-derived on the column axis with `length`
        const int lb = hf - cx, rb = hf - (cfg.width - 1 - cx);
        const int ub = hf - cy, db = hf - (cfg.length - 1 - cy);
        int r_lo = 0, r_hi = 0, q_lo = 0, q_hi = 0;   // [lo, hi) of rows / cols set to 1
        if (lb > 0) { r_lo = 0; r_hi = min(lb, fov); } else if (rb > 0) { r_lo = max(fov - rb, 0); r_hi = fov; }
        if (ub > 0) { q_lo = 0; q_hi = min(ub, fov); } else if (db > 0) { q_lo = max(fov - db, 0); q_hi = fov; }
        // whole rows are one contiguous byte range; the column band is written row by row, one lane per column
        {
            // bytes [b0, b1) of the record; unaligned head / tail by bytes, the middle as 4-byte words
            int8_t* const lay = rec + 2 * f2;
            const int b0 = r_lo * fov, b1 = r_hi * fov;
            if (b1 > b0) {
                const int mis = (int)(reinterpret_cast<uintptr_t>(lay + b0) & 3u);
                const int head = min((4 - mis) & 3, b1 - b0);
                const int nwords = (b1 - b0 - head) >> 2;
                const int tail0 = b0 + head + 4 * nwords;
                if (lane < head) lay[b0 + lane] = 1;
                uint32_t* w4 = reinterpret_cast<uint32_t*>(lay + b0 + head);
                for (int k = lane; k < nwords; k += 32) w4[k] = 0x01010101u;
                if (lane < b1 - tail0) lay[tail0 + lane] = 1;
            }
        }
        if (q_hi > q_lo) {
            const int nq = q_hi - q_lo;
            for (int q = lane; q < nq; q += 32)
                for (int r = 0; r < fov; ++r) rec[2 * f2 + r * fov + q_lo + q] = 1;
        }
        if (lane == 0) {                                                          // direction vector (:895)
            rec[3 * f2] = S.diry[gy - cy + cfg.width - 1];
            rec[3 * f2 + 1] = S.dirx[gx - cx + cfg.length - 1];
        }
    }
}

__device__ __forceinline__ void meda_paint_tile(const meda_cfg_t& cfg, const MedaLayout& L, const MedaSmem& S,
                                                int e_valid, uint8_t need_flag, const uint8_t* __restrict__ set_order)
{
    const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const uint32_t inv_a = 0xFFFFFFFFu / (uint32_t)L.A + 1u;                  // g / A == umulhi(g, inv_a) for g < 2^16
    for (int g = warp; g < e_valid * L.A; g += nwarps) {
        const int e = (int)__umulhi((uint32_t)g, inv_a), i = g - e * L.A;
        if (!(S.flag[e] & need_flag) || (S.flag[e] & kEnvFrozen)) continue;   // warp-uniform
        meda_paint_agent(cfg, S, S.word + e * L.A, i, S.tile + (size_t)g * L.D, set_order);
    }
}

// updateHealth (meda.py:600-605) for the flagged envs of the tile
__device__ __forceinline__ void meda_update_health(const meda_cfg_t& cfg, const meda_state_t& st, const MedaSmem& S,
                                                   int64_t n0, int e_valid)
{
    if (!cfg.b_degrade || !st.usage || !st.health) return;
    const int cells = cfg.width * cfg.length;
    for (int e = 0; e < e_valid; ++e) {
        if (!(S.flag[e] & kEnvSelected)) continue;
        uint32_t* usage = st.usage + (size_t)(n0 + e) * cells;
        double* health = st.health + (size_t)(n0 + e) * cells;
        const double* degrade = st.degrade ? st.degrade + (size_t)(n0 + e) * cells : nullptr;
        for (int k = threadIdx.x; k < cells; k += blockDim.x)
            if (usage[k] > 50) {
                health[k] = health[k] * (degrade ? degrade[k] : 1.0);
                usage[k] = 0;
            }
    }
}

__global__ void __launch_bounds__(kThreads)
meda_step_kernel(const __grid_constant__ meda_cfg_t cfg, const meda_state_t st, const void* __restrict__ actions, int aes,
                 const double* __restrict__ u, uint64_t seed, uint32_t flags, const uint8_t* __restrict__ set_order,
                 const meda_out_t out, int E)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const MedaLayout L(cfg, E);
    const MedaSmem S(smem_raw, L);
    const int A = L.A, W = cfg.width, Lc = cfg.length;
    const int64_t n0 = (int64_t)blockIdx.x * E;
    const int e_valid = (int)min((int64_t)E, (int64_t)st.n_envs - n0);
    const int cells = W * Lc;

    meda_prologue(cfg, L, S, true);
    for (int e = threadIdx.x; e < e_valid; e += blockDim.x)
        S.flag[e] = ((flags & DMFB_STEP_FREEZE_TERM) && st.terminated[n0 + e]) ? kEnvFrozen : kEnvSelected;
    __syncthreads();

    // ---- moveOneDroplet (meda.py:261-292): one thread per droplet, droplets are independent -----------
    for (int t = threadIdx.x; t < e_valid * A; t += blockDim.x) {
        const int e = t / A, i = t - e * A;
        const int64_t n = n0 + e;
        const size_t ja = (size_t)n * A + i;
        const uint32_t d = reinterpret_cast<const uint32_t*>(st.drop)[ja];
        int xc = d & 255u, yc = (d >> 8) & 255u;
        const int gx = (d >> 16) & 255u, gy = d >> 24;
        uint32_t status = st.status[ja];
        uint32_t code = 0;                                   // 0: 0.0, 1: -0.2, 2: -0.08, 3: -0.4
        if (!(S.flag[e] & kEnvFrozen) && !status) {          // sticky status: reward 0, nothing moves (:248-249)
            const int old2 = (xc - gx) * (xc - gx) + (yc - gy) * (yc - gy);
            if (old2 < 16) {                                 // distance < r_i + r_goal = 4: snap onto the goal (:272-277)
                xc = gx; yc = gy; status = 1;
            } else {
                const int a = load_action(actions, aes, ja);
                bool move = true;
                if (st.health) {                             // getMoveProb (:302-309): sequential float64 mean of 25 cells
                    const double* h = st.health + (size_t)n * cells;
                    double prob = 0.0;
                    for (int y = yc - kRad; y <= yc + kRad; ++y)
                        for (int x = xc - kRad; x <= xc + kRad; ++x) prob += h[y * Lc + x];
                    prob = prob / 25.0;
                    double draw;
                    if (u) draw = u[ja];
                    else {
                        const uint32_t episode = st.episode ? st.episode[n] : 0u;
                        const uint4 r = env_random(seed, kStreamMove, cfg.env_base + n, episode,
                                                   (uint32_t)st.step_count[n] + 1u, (uint32_t)i);
                        draw = u53(r.x, r.y);
                    }
                    move = draw <= prob;                     // random.random() <= prob (:280)
                }
                if (move) meda_move(xc, yc, a, W, Lc);
                const int new2 = (xc - gx) * (xc - gx) + (yc - gy) * (yc - gy);
                // (:283-290) all comparisons of the float distances are exact on the integer squares
                code = (new2 < 16) ? 0u : (new2 == old2 && a == 8) ? 1u : (new2 < old2) ? 2u : 3u;
            }
        }
        S.word[t] = (d & 0xFFFF0000u) | (uint32_t)xc | ((uint32_t)yc << 8);
        S.misc[t] = status | (code << 8);
    }
    __syncthreads();

    // ---- calPunish (:321-330) + MEDAEnv.step bookkeeping (:521-538): one thread per droplet; every thread
    //      recomputes the (cheap, A^2) pairwise counts of its env, the env scalars are written by droplet 0 ------
    for (int t = threadIdx.x; t < e_valid * A; t += blockDim.x) {
        const int e = t / A, i = t - e * A;
        const int64_t n = n0 + e;
        const size_t ja = (size_t)n * A + i;
        const uint32_t* words = S.word + e * A;
        const uint32_t* misc = S.misc + e * A;
        const bool frozen = S.flag[e] & kEnvFrozen;
        int total = 0, all = 1, my_pun = 0;
        for (int a = 0; a < A; ++a) {
            const int xa = words[a] & 255u, ya = (words[a] >> 8) & 255u;
            int pun = 0;
            for (int j = 0; j < A; ++j) {
                const int dx = xa - (int)(words[j] & 255u), dy = ya - (int)((words[j] >> 8) & 255u);
                pun += (j != a) & (dx * dx + dy * dy < 36);   // centre distance < 1.5 * (r_i + r_j) = 6
            }
            total += pun;
            all &= (int)(misc[a] & 1u);
            if (a == i) my_pun = pun;
        }
        if (frozen) total = 0;
        const int fails = st.fails[n] + total;                // the reference keeps -0.6 * this count (:521)
        const int sc = st.step_count[n] + (frozen ? 0 : 1);
        const uint32_t m = misc[i];
        const uint32_t code = (m >> 8) & 3u;
        double r = code == 0 ? 0.0 : code == 1 ? -0.2 : code == 2 ? -0.08 : -0.4;
        if (my_pun) {                                         // punish[i] -= 0.6, pun times; rewards[i] += punish[i]
            double p = 0.0;
            for (int k = 0; k < my_pun; ++k) p -= 0.6;
            r = r + p;
        }
        if (all) {                                            // (:522-525)
            r = r + 3.0;
            if (fails == 0) r = r + 3.0;
        }
        if (frozen) r = 0.0;
        const bool in_time = !frozen && sc < cfg.max_step;    // (:529-537)
        const uint32_t done = in_time ? (m & 1u) : 1u;
        if (out.reward) out.reward[ja] = (float)r;
        if (out.reward_f64) out.reward_f64[ja] = r;
        if (out.done) out.done[ja] = (uint8_t)done;
        if (!frozen) {
            reinterpret_cast<uint32_t*>(st.drop)[ja] = words[i];
            st.status[ja] = (uint8_t)(m & 1u);
        }
        S.rew[t] = (float)r;
        S.done[t] = (uint8_t)done;
        if (i == 0) {
            const int term = in_time ? all : 1;
            S.flag[e] = (uint8_t)(S.flag[e] | (in_time ? kEnvUsage : 0));
            S.envi[2 * e] = fails;
            S.envi[2 * e + 1] = sc;
            if (out.constraints) out.constraints[n] = total;
            if (out.success) out.success[n] = (uint8_t)((in_time && all && fails == 0) ? 1 : 0);
            if (out.terminated) out.terminated[n] = (uint8_t)term;
            if (out.padded) out.padded[n] = (uint8_t)frozen;
            if (!frozen) st.terminated[n] = (uint8_t)term;
        }
        if (out.avail) {
            uint8_t* av = out.avail + ja * cfg.n_actions;
            for (int k = 0; k < cfg.n_actions; ++k) av[k] = frozen ? 0 : 1;
        }
    }
    __syncthreads();
    // env counters are read by every droplet thread above, so they are written only after the barrier
    for (int e = threadIdx.x; e < e_valid; e += blockDim.x) {
        const int64_t n = n0 + e;
        if (!(S.flag[e] & kEnvFrozen)) {
            st.fails[n] = S.envi[2 * e];
            st.step_count[n] = S.envi[2 * e + 1];
        }
        if (out.team_reward) {
            float sum = 0.f;
            for (int i = 0; i < A; ++i) sum += S.rew[e * A + i];
            out.team_reward[n] = sum / (float)A;
        }
    }

    // ---- addUsage (:591-598): footprints of one env may overlap -> RED.ADD per cell, one warp per droplet ----
    if (st.usage) {
        const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5, lane = threadIdx.x & 31;
        const uint32_t inv_a = 0xFFFFFFFFu / (uint32_t)A + 1u;
        for (int g = warp; g < e_valid * A; g += nwarps) {
            const int e = (int)__umulhi((uint32_t)g, inv_a);
            if (!(S.flag[e] & kEnvUsage) || (S.flag[e] & kEnvFrozen)) continue;
            const uint32_t w = S.word[g];
            if (!S.done[g] && lane < kFootCells) {
                const int x = (int)(w & 255u) - kRad + lane % 5, y = (int)((w >> 8) & 255u) - kRad + lane / 5;
                atomicAdd(st.usage + (size_t)(n0 + e) * cells + y * Lc + x, 1u);
            }
        }
    }
    meda_paint_tile(cfg, L, S, e_valid, kEnvSelected, set_order);
    store_tile(out.obs + (size_t)n0 * A * L.D, S.tile, (uint32_t)(e_valid * A * L.D));
}

// refresh/addTask/_genLegalDroplet (meda.py:161-185,213-233): centres uniform in [r, dim-r-1]; a droplet
// (destination) is redrawn while its centre is closer than 1.5*(2+2+2) = 9 to an earlier droplet (destination);
// the destination is also redrawn while it overlaps its own droplet.  One thread per env (sequential by nature).
__device__ void meda_generate_tasks(const meda_cfg_t& cfg, uint64_t seed, int64_t env, uint32_t episode, uint32_t* words)
{
    const int A = cfg.n_agents, W = cfg.width, Lc = cfg.length;
    uint64_t state = seed ^ (0x9E3779B97F4A7C15ull * (uint64_t)(kStreamLayout + 1));
    state += (uint64_t)env * 0xD1342543DE82EF95ull + ((uint64_t)episode << 32) * 0xDA942042E4DD58B5ull;
    state = mix64(state);
    for (int i = 0; i < A; ++i) {
        uint32_t sx, sy, tx, ty;
        for (;;) {
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
        for (;;) {
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
    meda_prologue(cfg, L, S, obs != nullptr);
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
                meda_generate_tasks(cfg, seed, cfg.env_base + n, episode, words);
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
                for (int k = threadIdx.x; k < cells; k += blockDim.x) {
                    if (st.usage) st.usage[(size_t)n * cells + k] = 0;
                    if (st.health) st.health[(size_t)n * cells + k] = 1.0;
                    if (st.degrade) {
                        double dg = 1.0;
                        if (degrade_in) dg = degrade_in[(size_t)n * cells + k];
                        else if (cfg.b_degrade) {
                            const uint4 r = env_random(seed, kStreamDegrade, cfg.env_base + n, episode, (uint32_t)k, 0u);
                            dg = u53(r.x, r.y) * 0.4 + 0.6;
                            if (u53(r.z, r.w) < 1.0 - cfg.per_degrade) dg = 1.0;
                        }
                        st.degrade[(size_t)n * cells + k] = dg;
                    }
                }
            }
        } else {
            meda_update_health(cfg, st, S, n0, e_valid);  // after the observation in the reference (:547-548); obs does not read health
        }
    }
    if (obs == nullptr) return;
    meda_paint_tile(cfg, L, S, e_valid, kEnvSelected, set_order);
    int8_t* gobs = obs + (size_t)n0 * A * L.D;
    if (n_selected == e_valid) store_tile(gobs, S.tile, (uint32_t)(e_valid * A * L.D));
    else store_rows_masked(gobs, S.tile, e_valid, A * L.D, S.flag);
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
    if (obs_version != MEDA_OBS_BASE && obs_version != MEDA_OBS_V02) return DMFB_ERR_BAD_ARG;
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
    if (cfg->obs_version == MEDA_OBS_V02 && cfg->n_agents > 8 && !set_order) {
        snprintf(g_last_error, sizeof(g_last_error), "meda_step: v0_2 obs with more than 8 agents needs set_order");
        return DMFB_ERR_BAD_ARG;
    }
    if (state->n_envs == 0) return DMFB_OK;
    const int E = meda_tile_envs(*cfg);
    const MedaLayout L(*cfg, E);
    if (L.total > 48 * 1024)
        DMFB_CUDA_TRY(cudaFuncSetAttribute(meda_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
    const int grid = (state->n_envs + E - 1) / E;
    meda_step_kernel<<<grid, kThreads, L.total, static_cast<cudaStream_t>(stream)>>>(*cfg, *state, actions, action_elem_size,
                                                                                      u_inject, seed, flags, set_order, *out, E);
    g_launches.fetch_add(1);
    DMFB_CUDA_TRY(cudaGetLastError());
    if (flags & DMFB_STEP_AUTO_RESET)
        return meda_launch_reset(cfg, state, state->terminated, 0, 0, nullptr, nullptr, seed, set_order, out->obs, stream);
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
