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
//  * the tile is staged in shared memory: 16-byte zero fill, boundary layer expanded from per-agent
//    bit masks (4 output bytes per multiply), sparse byte scatter for droplet ids / clipped goals /
//    direction bytes, then ONE TMA bulk store (cp.async.bulk.global.shared::cta) per tile;
//  * dynamics (sequential, order dependent moves) run one thread per env on packed (x,y,gx,gy)
//    words held in shared memory; all small outputs are staged and written coalesced;
//  * HBM-bound integer/byte work: no tensor cores.
#include "common.cuh"

namespace dmfb {

thread_local char g_last_error[256] = "";
std::atomic<uint64_t> g_launches{0};

namespace {

constexpr int kThreads = 128;

// Shared-memory carve-up of one tile, computed identically on host and device.
struct TileLayout {
    int E, A, SA, D, nw, ncodes;
    uint32_t tile_bytes;
    uint32_t off_drop, off_past, off_rew, off_flag, off_donemask, off_l2row, off_l2col, off_dirx, off_diry;
    uint32_t off_f64, off_i32, off_act, total;
    __host__ __device__ TileLayout(const dmfb_cfg_t& c, int E_) {
        E = E_; A = c.n_agents; SA = A | 1; D = c.obs_dim; nw = c.l2_words; ncodes = 2 * (c.fov / 2) + 1;
        tile_bytes = ((uint32_t)(E * A * D) + 15u) & ~15u;
        uint32_t o = tile_bytes;
        off_f64 = o; o += (uint32_t)(E * A) * 8u * 2u;   // staged draws + health probabilities (float64)
        off_drop = o; o += (uint32_t)(E * SA) * 4u;
        off_past = o; o += (uint32_t)(E * SA) * 4u;
        off_rew = o; o += (uint32_t)(E * SA) * 4u;
        off_donemask = o; o += (uint32_t)E * 4u;
        off_i32 = o; o += (uint32_t)E * 4u * 8u;          // per-env scalars, see EnvScalar
        off_l2row = o; o += (uint32_t)(ncodes * nw) * 4u;
        off_l2col = o; o += (uint32_t)(ncodes * nw) * 4u;
        off_act = o; o += ((uint32_t)(E * A) + 3u) & ~3u;
        off_flag = o; o += ((uint32_t)E + 3u) & ~3u;
        off_dirx = o; o += ((uint32_t)(2 * c.width) + 3u) & ~3u;
        off_diry = o; o += ((uint32_t)(2 * c.length) + 3u) & ~3u;
        total = (o + 15u) & ~15u;
    }
};

// per-env scalars staged in smem: [slot][E] int32
enum EnvScalar { kStepIn = 0, kCumIn, kEpisode, kStepOut, kCumOut, kCons, kMisc /* success | term<<8 | padded<<16 | reset<<24 */, kTeam, kNumScalars };
static_assert(kNumScalars == 8, "TileLayout reserves 8 scalar slots");

// S.flag values
constexpr uint8_t kFlagFrozen = 1;    // padded step: zero observation
constexpr uint8_t kFlagSkip = 2;      // masked reset: env not selected
constexpr uint8_t kFlagNewTask = 4;   // auto-reset: a new task was generated, updateHealth still to run

struct TileSmem {
    int8_t* tile;
    double* draw;        // [E*A]
    double* prob;        // [E*A]
    uint32_t* drop;      // [E][SA] packed x | y<<8 | gx<<16 | gy<<24
    uint32_t* past;      // [E][SA] past x | y<<8 | sta<<16 | dyn<<24
    float* rew;          // [E][SA]
    uint32_t* donemask;  // [E]
    int32_t* sc;         // [kNumScalars][E]
    uint32_t* l2row;     // [ncodes][nw]
    uint32_t* l2col;
    int8_t* act;         // [E*A]
    uint8_t* flag;       // [E]
    int8_t* dirx;
    int8_t* diry;
    __device__ TileSmem(unsigned char* base, const TileLayout& L) {
        tile = reinterpret_cast<int8_t*>(base);
        draw = reinterpret_cast<double*>(base + L.off_f64);
        prob = draw + L.E * L.A;
        drop = reinterpret_cast<uint32_t*>(base + L.off_drop);
        past = reinterpret_cast<uint32_t*>(base + L.off_past);
        rew = reinterpret_cast<float*>(base + L.off_rew);
        donemask = reinterpret_cast<uint32_t*>(base + L.off_donemask);
        sc = reinterpret_cast<int32_t*>(base + L.off_i32);
        l2row = reinterpret_cast<uint32_t*>(base + L.off_l2row);
        l2col = reinterpret_cast<uint32_t*>(base + L.off_l2col);
        act = reinterpret_cast<int8_t*>(base + L.off_act);
        flag = reinterpret_cast<uint8_t*>(base + L.off_flag);
        dirx = reinterpret_cast<int8_t*>(base + L.off_dirx);
        diry = reinterpret_cast<int8_t*>(base + L.off_diry);
    }
};

__device__ __forceinline__ void load_tables(const dmfb_cfg_t& cfg, const TileLayout& L, const TileSmem& S)
{
    const int nt = L.ncodes * L.nw;
    for (int k = threadIdx.x; k < nt; k += blockDim.x) {
        const int c = k / L.nw, j = k - c * L.nw;
        S.l2row[k] = cfg.l2_row[c][j];
        S.l2col[k] = cfg.l2_col[c][j];
    }
    for (int k = threadIdx.x; k < 2 * cfg.width; k += blockDim.x) S.dirx[k] = cfg.dir_x[k];
    for (int k = threadIdx.x; k < 2 * cfg.length; k += blockDim.x) S.diry[k] = cfg.dir_y[k];
}

__device__ __forceinline__ void zero_tile(const TileLayout& L, const TileSmem& S)
{
    uint4* t4 = reinterpret_cast<uint4*>(S.tile);
    const int n16 = (int)(L.tile_bytes >> 4);
    for (int k = threadIdx.x; k < n16; k += blockDim.x) t4[k] = make_uint4(0u, 0u, 0u, 0u);
}

__device__ __forceinline__ int near1(uint32_t a, uint32_t b)
{
    // |ax-bx| <= 1 && |ay-by| <= 1  <=>  Euclid < 2 on the integer grid (dmfb.py:258,268)
    const int dx = (int)(a & 255u) - (int)(b & 255u);
    const int dy = (int)((a >> 8) & 255u) - (int)((b >> 8) & 255u);
    return (int)((unsigned)(dx + 1) <= 2u) & (int)((unsigned)(dy + 1) <= 2u);
}

// getObs() for the live envs of a tile (dmfb.py:395-457, 614-626): one thread per agent.
// The tile must already be zero filled and S.drop / S.flag valid (barrier before the call).
template <int FOV_T>
__device__ __forceinline__ void paint_agents(const dmfb_cfg_t& cfg, const TileLayout& L, const TileSmem& S,
                                             int e_valid)
{
    const int fov = FOV_T ? FOV_T : cfg.fov;
    const int hf = fov >> 1, f2 = fov * fov;
    const int A = L.A, SA = L.SA, D = L.D;
    const int nw = FOV_T ? (FOV_T * FOV_T + 31) / 32 : L.nw;
    const int W = cfg.width, Lc = cfg.length;
    for (int g = threadIdx.x; g < e_valid * A; g += blockDim.x) {
        const int e = g / A, i = g - e * A;
        if (S.flag[e] & (kFlagFrozen | kFlagSkip)) continue;
        const uint32_t* drop = S.drop + e * SA;
        const uint32_t me = drop[i];
        const int x = me & 255u, y = (me >> 8) & 255u, gx = (me >> 16) & 255u, gy = me >> 24;
        int8_t* rec = S.tile + (size_t)g * D;

        // ---- layer 2: off-chip boundary (dmfb.py:428-439) expanded from bit masks -------------
        {
            const int lb = hf - x, rb = hf - (W - 1 - x);
            const int ub = hf - y, db = hf - (Lc - 1 - y);
            const int rc = lb > 0 ? lb : (rb > 0 ? hf + rb : 0);
            const int cc = ub > 0 ? ub : (db > 0 ? hf + db : 0);
            const uint32_t* rowm = S.l2row + rc * nw;
            const uint32_t* colm = S.l2col + cc * nw;
            const int base = g * D + 2 * f2;      // byte offset of the layer inside the tile
            const int s = base & 3;               // misalignment of the layer start
            uint32_t* wptr = reinterpret_cast<uint32_t*>(S.tile + (base & ~3));
            const int nbits = f2 + s;
            const int nfull = nbits >> 2;         // whole 4-byte words
            const int ntail = nbits & 3;          // trailing bytes handled one by one (next agent's bytes follow)
            if (rc | cc) {
                uint32_t prev = 0;
                int k = 0;
#pragma unroll
                for (int j = 0; j <= nw; ++j) {
                    const uint32_t m = (j < nw) ? (rowm[j] | colm[j]) : 0u;
                    const uint32_t sh = __funnelshift_l(prev, m, s);  // bits of the mask shifted up by s
                    prev = m;
#pragma unroll
                    for (int t = 0; t < 8; ++t, ++k) {
                        const uint32_t nib = (sh >> (4 * t)) & 0xFu;
                        if (k < nfull) {
                            // spread 4 bits to 4 bytes: bit b -> byte b (no carries: 16 distinct partial products)
                            wptr[k] = (nib * 0x00204081u) & 0x01010101u;
                        } else if (k == nfull) {
                            int8_t* bp = reinterpret_cast<int8_t*>(wptr + k);
                            for (int b = 0; b < ntail; ++b) bp[b] = (int8_t)((nib >> b) & 1u);
                        }
                    }
                }
            }
        }
        // ---- layers 0 / 1: droplets in the window, clipped goals of visible others (:408-420) ---
        const int ox = x - hf, oy = y - hf;
        for (int j = 0; j < A; ++j) {
            const uint32_t d = drop[j];
            const int jx = d & 255u, jy = (d >> 8) & 255u;
            const int rx = jx - ox, ry = jy - oy;
            if ((unsigned)rx < (unsigned)fov && (unsigned)ry < (unsigned)fov) rec[rx * fov + ry] = (int8_t)(j + 1);
            if (j != i && 2 * abs(jx - x) < fov && 2 * abs(jy - y) < fov) {
                int cx = (int)((d >> 16) & 255u) - ox, cy = (int)(d >> 24) - oy;
                cx = min(max(cx, 0), fov - 1);
                cy = min(max(cy, 0), fov - 1);
                rec[f2 + cx * fov + cy] = (int8_t)(j + 1);  // ascending j: later index overwrites
            }
        }
        // ---- direction bytes (:442-454) from the host-built table ------------------------------
        rec[3 * f2] = S.dirx[gx - x + W - 1];
        rec[3 * f2 + 1] = S.diry[gy - y + Lc - 1];
    }
}

// _Generate_Start_End (dmfb.py:207-226): 2A uniform cells, whole set redrawn until every pairwise squared
// distance is > 2.  Drawing point by point and restarting at the first conflict accepts exactly the same
// sets with the same probabilities.
__device__ __noinline__ void generate_layout(const dmfb_cfg_t& cfg, uint64_t seed, int64_t env, uint32_t episode,
                                             uint32_t* drop)
{
    const int A = cfg.n_agents, m = 2 * A;
    uint8_t px[2 * DMFB_MAX_AGENTS], py[2 * DMFB_MAX_AGENTS];
    uint32_t attempt = 0;
    for (;;) {
        bool ok = true;
        for (int k = 0; k < m && ok; k += 2) {
            const uint4 r = env_random(seed, kStreamLayout, env, episode, attempt, (uint32_t)k);
            const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
            for (int q = 0; q < 2 && k + q < m && ok; ++q) {
                const int x = (int)__umulhi(rr[2 * q], (uint32_t)cfg.width);
                const int y = (int)__umulhi(rr[2 * q + 1], (uint32_t)cfg.length);
                for (int j = 0; j < k + q; ++j) {
                    const int dx = x - px[j], dy = y - py[j];
                    if (dx * dx + dy * dy <= 2) { ok = false; break; }
                }
                px[k + q] = (uint8_t)x;
                py[k + q] = (uint8_t)y;
            }
        }
        if (ok) break;
        ++attempt;
    }
    for (int i = 0; i < A; ++i)
        drop[i] = (uint32_t)px[i] | ((uint32_t)py[i] << 8) | ((uint32_t)px[A + i] << 16) | ((uint32_t)py[A + i] << 24);
}

// One env of DMFBenv.step; executed by one thread on inputs already staged in shared memory.
// Everything it produces goes back to shared memory; the CTA writes it out coalesced afterwards.
__device__ __forceinline__ void step_one_env(const dmfb_cfg_t& cfg, const dmfb_state_t& st, const TileLayout& L,
                                             const TileSmem& S, int e, int64_t n, bool have_prob, bool have_draw,
                                             uint64_t seed, uint32_t flags, const dmfb_out_t& out)
{
    const int A = L.A, SA = L.SA, E = L.E, W = cfg.width, Lc = cfg.length;
    uint32_t* drop = S.drop + e * SA;
    uint32_t* past = S.past + e * SA;
    float* rew = S.rew + e * SA;
    const uint32_t all_mask = (A >= 32) ? 0xFFFFFFFFu : ((1u << A) - 1u);

    if (S.flag[e] & kFlagFrozen) {
        // lock-step padding (rollout.py:131-141): zero obs / reward / avail, terminated = padded = 1
        S.donemask[e] = all_mask;
        for (int i = 0; i < A; ++i) {
            rew[i] = 0.f;
            if (out.reward_f64) out.reward_f64[(size_t)n * A + i] = 0.0;
        }
        S.sc[kStepOut * E + e] = S.sc[kStepIn * E + e];
        S.sc[kCumOut * E + e] = S.sc[kCumIn * E + e];
        S.sc[kCons * E + e] = 0;
        S.sc[kMisc * E + e] = (1 << 8) | (1 << 16);
        S.sc[kTeam * E + e] = __float_as_int(0.f);
        return;
    }

    const int sc = S.sc[kStepIn * E + e] + 1;                    // dmfb.py:561
    const uint32_t episode = (uint32_t)S.sc[kEpisode * E + e];
    uint32_t pre_done = 0;                                        // getTaskStatus before the moves (:278)
    uint64_t base_code = 0;                                       // 2 bits per droplet: 0 -> 0.0, 1 -> -0.1, 2 -> -0.25, 3 -> -0.4
    bool illegal = false;

    for (int i = 0; i < A; ++i) {                                 // moveOneDroplet, sequential (:279-283, 325-359)
        const uint32_t d = drop[i];
        const int x = d & 255u, y = (d >> 8) & 255u, gx = (d >> 16) & 255u, gy = d >> 24;
        past[i] = d & 0xFFFFu;
        const int od = abs(x - gx) + abs(y - gy);
        if (od == 0) pre_done |= 1u << i;
        uint32_t code;
        if (cfg.stall && od == 0) {
            code = 0;                                             // reward 0, no move, no draw (:331-332)
        } else {
            const int a = S.act[e * A + i];
            bool move = true;
            if (have_prob) {
                const double prob = S.prob[e * A + i];            // getMoveProb (:361-363), cell = position at step start
                double draw;
                if (have_draw) {
                    draw = S.draw[e * A + i];
                } else {
                    const uint4 r = env_random(seed, kStreamMove, cfg.env_base + n, episode, (uint32_t)sc, (uint32_t)i);
                    draw = u53(r.x, r.y);
                }
                move = draw <= prob;                              // random.random() <= prob (:335)
            }
            int nx = x, ny = y;
            if (move) {
                if ((unsigned)a > 4u) illegal = true;             // TypeError('action is illegal') (:115-116)
                nx = x + (a == 1) - (a == 2);                     // Droplet.move (:103-124)
                ny = y + (a == 4) - (a == 3);
                nx = min(max(nx, 0), W - 1);
                ny = min(max(ny, 0), Lc - 1);
                const uint32_t cand = (uint32_t)nx | ((uint32_t)ny << 8);
                bool hit = false;                                 // _isinvalidaction (:310-323): cell taken?
                for (int j = 0; j < A; ++j) hit |= (j != i) & ((drop[j] & 0xFFFFu) == cand);
                if (hit) { nx = x; ny = y; }
            }
            const int nd = abs(nx - gx) + abs(ny - gy);
            code = (nd == od && od == 0) ? 1u : (nd == od && a == 0) ? 2u : (nd < od) ? 1u : 3u;  // (:345-354)
            drop[i] = (d & 0xFFFF0000u) | (uint32_t)nx | ((uint32_t)ny << 8);
        }
        base_code |= (uint64_t)code << (2 * i);
    }

    // comflic_static / comflic_dynamic (:254-271) on the final and the saved positions
    int constraints = 0;
    uint32_t post_done = 0;
    for (int k = 0; k < A; ++k) {
        const uint32_t ck = drop[k], pk = past[k];
        int sta = 0, dyn = 0;
        for (int j = 0; j < A; ++j) {
            if (j == k) continue;
            const uint32_t cj = drop[j], pj = past[j];
            sta += near1(ck, cj);
            dyn += near1(pk, cj) + near1(pj, ck);
        }
        constraints += sta + dyn;
        past[k] = (pk & 0xFFFFu) | ((uint32_t)sta << 16) | ((uint32_t)dyn << 24);
        if ((ck & 0xFFFFu) == (ck >> 16)) post_done |= 1u << k;
    }
    const bool all_done = (post_done == all_mask);               // np.all(getTaskStatus()) after the moves (:293)

    double sum = 0.0;
    for (int k = 0; k < A; ++k) {
        const uint32_t code = (uint32_t)(base_code >> (2 * k)) & 3u;
        double r = code == 0 ? 0.0 : code == 1 ? -0.1 : code == 2 ? -0.25 : -0.4;
        const uint32_t pk = past[k];
        r = r - (double)(2 * (int)((pk >> 16) & 255u));           // rewards - 2*sta - 2*dy in float64 (:288)
        r = r - (double)(2 * (int)(pk >> 24));
        if (cfg.stall && ((pre_done >> k) & 1u)) r = 0.0;         // (:289-292)
        if (all_done) {                                           // (:293-296)
            r = r + 10.0;
            if (constraints == 0) r = r + 10.0;
        }
        rew[k] = (float)r;
        if (out.reward_f64) out.reward_f64[(size_t)n * A + k] = r;
        sum += r;
    }

    if ((flags & DMFB_STEP_RECORD_USAGE) && st.usage) {           // addUsage (:459-463)
        uint16_t* usage = st.usage + (size_t)n * W * Lc;
        for (int k = 0; k < A; ++k)
            if (!((post_done >> k) & 1u)) {
                const uint32_t ck = drop[k];
                uint16_t* cell = usage + (ck & 255u) * Lc + ((ck >> 8) & 255u);
                const uint16_t v = *cell;
                *cell = (uint16_t)(v + (v != 0xFFFFu));
            }
    }
    int cum = S.sc[kCumIn * E + e] + constraints;                 // (:572)
    int sc_out = sc;
    uint32_t done_mask;
    int success = 0;
    if (sc < cfg.max_step) {                                      // (:577-585)
        success = (all_done && cum == 0) ? 1 : 0;
        done_mask = post_done;
    } else {
        done_mask = all_mask;
    }
    S.donemask[e] = done_mask;
    const int term = (done_mask == all_mask) ? 1 : 0;
    int did_reset = 0;
    if (term && (flags & DMFB_STEP_AUTO_RESET)) {
        // DMFBenv.reset(new=False) (:589-597) fused into the step: new task now, updateHealth by the CTA later
        const uint32_t ep2 = episode + 1u;
        generate_layout(cfg, seed, cfg.env_base + n, ep2, drop);
        if (st.episode) st.episode[n] = ep2;
        if (st.start)
            for (int i = 0; i < A; ++i)
                reinterpret_cast<uint16_t*>(st.start)[(size_t)n * A + i] = (uint16_t)(drop[i] & 0xFFFFu);
        sc_out = 0;
        cum = 0;
        did_reset = 1;
        S.flag[e] = kFlagNewTask;
    }
    S.sc[kStepOut * E + e] = sc_out;
    S.sc[kCumOut * E + e] = cum;
    S.sc[kCons * E + e] = constraints;
    S.sc[kMisc * E + e] = success | (term << 8) | (did_reset << 24);
    S.sc[kTeam * E + e] = __float_as_int((float)(sum / (double)A));  // rollout.py:33
    if (illegal && out.status) atomicOr(out.status, 1);
}

template <int FOV_T>
__global__ void __launch_bounds__(kThreads)
dmfb_step_kernel(const __grid_constant__ dmfb_cfg_t cfg, const dmfb_state_t st, const void* __restrict__ actions,
                 int aes, const double* __restrict__ u, uint64_t seed, uint32_t flags, const dmfb_out_t out, int E)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const TileLayout L(cfg, E);
    const TileSmem S(smem_raw, L);
    const int64_t n0 = (int64_t)blockIdx.x * E;
    const int e_valid = (int)min((int64_t)E, (int64_t)st.n_envs - n0);
    const int A = L.A, SA = L.SA;
    const int W = cfg.width, Lc = cfg.length;
    const bool have_prob = st.health != nullptr;
    const bool have_draw = have_prob && (u != nullptr);

    // ---- phase A: every global input of the tile is fetched coalesced, in one round trip ----------
    {
        const uint32_t* gdrop = reinterpret_cast<const uint32_t*>(st.drop) + (size_t)n0 * A;
        const size_t gbase = (size_t)n0 * A;
        for (int j = threadIdx.x; j < e_valid * A; j += blockDim.x) {
            const int e = j / A, i = j - e * A;
            const uint32_t d = gdrop[j];
            S.drop[e * SA + i] = d;
            S.act[j] = (int8_t)load_action(actions, aes, gbase + j);
            if (have_prob) {
                S.prob[j] = st.health[((size_t)(n0 + e) * W + (d & 255u)) * Lc + ((d >> 8) & 255u)];
                if (have_draw) S.draw[j] = u[gbase + j];
            }
        }
        for (int e = threadIdx.x; e < e_valid; e += blockDim.x) {
            S.sc[kStepIn * E + e] = st.step_count[n0 + e];
            S.sc[kCumIn * E + e] = st.constraints[n0 + e];
            S.sc[kEpisode * E + e] = st.episode ? (int32_t)st.episode[n0 + e] : 0;
            S.flag[e] = ((flags & DMFB_STEP_FREEZE_TERM) && st.terminated[n0 + e]) ? kFlagFrozen : 0;
        }
    }
    load_tables(cfg, L, S);
    zero_tile(L, S);
    __syncthreads();

    // ---- phase B: dynamics, one thread per env, shared memory only ---------------------------------
    if ((int)threadIdx.x < e_valid)
        step_one_env(cfg, st, L, S, threadIdx.x, n0 + threadIdx.x, have_prob, have_draw, seed, flags, out);
    __syncthreads();

    // ---- phase C: coalesced write-back of state and of the small outputs ---------------------------
    {
        uint32_t* gdrop = reinterpret_cast<uint32_t*>(st.drop) + (size_t)n0 * A;
        const size_t gbase = (size_t)n0 * A;
        for (int j = threadIdx.x; j < e_valid * A; j += blockDim.x) {
            const int e = j / A, i = j - e * A;
            gdrop[j] = S.drop[e * SA + i];
            if (out.reward) out.reward[gbase + j] = S.rew[e * SA + i];
            if (out.done) out.done[gbase + j] = (uint8_t)((S.donemask[e] >> i) & 1u);
        }
        if (out.avail) {
            const int per_env = A * cfg.n_actions;
            uint8_t* gav = out.avail + (size_t)n0 * per_env;
            for (int j = threadIdx.x; j < e_valid * per_env; j += blockDim.x)
                gav[j] = (S.flag[j / per_env] & kFlagFrozen) ? 0 : 1;
        }
        for (int e = threadIdx.x; e < e_valid; e += blockDim.x) {
            const int64_t n = n0 + e;
            const int misc = S.sc[kMisc * E + e];
            const int term = (misc >> 8) & 1, did_reset = (misc >> 24) & 1;
            st.step_count[n] = S.sc[kStepOut * E + e];
            st.constraints[n] = S.sc[kCumOut * E + e];
            st.terminated[n] = (uint8_t)(term & !did_reset);
            if (out.team_reward) out.team_reward[n] = __int_as_float(S.sc[kTeam * E + e]);
            if (out.constraints) out.constraints[n] = S.sc[kCons * E + e];
            if (out.success) out.success[n] = (uint8_t)(misc & 1);
            if (out.terminated) out.terminated[n] = (uint8_t)term;
            if (out.padded) out.padded[n] = (uint8_t)((misc >> 16) & 1);
        }
        // fused auto-reset: updateHealth (dmfb.py:465-471) of the envs that just got a new task
        if ((flags & DMFB_STEP_AUTO_RESET) && st.usage) {
            const int cells = W * Lc;
            for (int e = 0; e < e_valid; ++e) {
                if (!(S.flag[e] & kFlagNewTask)) continue;
                uint16_t* usage = st.usage + (size_t)(n0 + e) * cells;
                double* health = st.health ? st.health + (size_t)(n0 + e) * cells : nullptr;
                const double* degrade = st.degrade ? st.degrade + (size_t)(n0 + e) * cells : nullptr;
                for (int k = threadIdx.x; k < cells; k += blockDim.x)
                    if (usage[k] > 50) {
                        if (health) health[k] = health[k] * (degrade ? degrade[k] : 1.0);
                        usage[k] = 0;
                    }
            }
        }
    }
    paint_agents<FOV_T>(cfg, L, S, e_valid);
    store_tile(out.obs + (size_t)n0 * A * L.D, S.tile, (uint32_t)(e_valid * A * L.D));
}

// --------------------------------------------------------------------- reset --

// mode 0: reset (new task), mode 1: restart (back to start cells), mode 2: observe only
template <int FOV_T>
__global__ void __launch_bounds__(kThreads)
dmfb_reset_kernel(const __grid_constant__ dmfb_cfg_t cfg, const dmfb_state_t st, const uint8_t* __restrict__ mask,
                  int mode, int new_task, const uint8_t* __restrict__ layouts, const double* __restrict__ degrade_in,
                  uint64_t seed, int8_t* __restrict__ obs, int E)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const TileLayout L(cfg, E);
    const TileSmem S(smem_raw, L);
    const int64_t n0 = (int64_t)blockIdx.x * E;
    const int e_valid = (int)min((int64_t)E, (int64_t)st.n_envs - n0);
    const int A = L.A, SA = L.SA;
    const int cells = cfg.width * cfg.length;

    int selected = 0;
    if ((int)threadIdx.x < e_valid) selected = (mask == nullptr) || (mask[n0 + threadIdx.x] != 0);
    const int n_selected = __syncthreads_count(selected);
    if (n_selected == 0) return;  // nothing to reset in this tile (the common case of a masked auto-reset)
    load_tables(cfg, L, S);
    if (obs) zero_tile(L, S);
    if ((int)threadIdx.x < e_valid) {
        const int e = threadIdx.x;
        const int64_t n = n0 + e;
        S.flag[e] = selected ? 0 : kFlagSkip;
        uint32_t* drop = S.drop + e * SA;
        uint32_t* gdrop = reinterpret_cast<uint32_t*>(st.drop) + (size_t)n * A;
        if (selected && mode != 2) {
            if (mode == 0) {
                const uint32_t episode = st.episode ? st.episode[n] + 1u : 0u;
                if (st.episode) st.episode[n] = episode;
                if (layouts) {
                    const uint32_t* lay = reinterpret_cast<const uint32_t*>(layouts) + (size_t)n * A;
                    for (int i = 0; i < A; ++i) drop[i] = lay[i];
                } else {
                    generate_layout(cfg, seed, cfg.env_base + n, episode, drop);
                }
                if (st.start)
                    for (int i = 0; i < A; ++i)
                        reinterpret_cast<uint16_t*>(st.start)[(size_t)n * A + i] = (uint16_t)(drop[i] & 0xFFFFu);
            } else {  // restart: droplets back to their start cells (dmfb.py:185-190)
                for (int i = 0; i < A; ++i)
                    drop[i] = (gdrop[i] & 0xFFFF0000u) | reinterpret_cast<const uint16_t*>(st.start)[(size_t)n * A + i];
            }
            for (int i = 0; i < A; ++i) gdrop[i] = drop[i];
            st.step_count[n] = 0;
            st.constraints[n] = 0;
            st.terminated[n] = 0;
        } else {
            for (int i = 0; i < A; ++i) drop[i] = gdrop[i];
        }
    }
    __syncthreads();

    // refresh(new) (dmfb.py:174-183): new -> health=1, usage=0, degrade redrawn; else updateHealth (:465-471)
    if (mode == 0 && (st.usage || st.health || st.degrade)) {
        for (int e = 0; e < e_valid; ++e) {
            if (S.flag[e]) continue;
            const int64_t n = n0 + e;
            uint16_t* usage = st.usage ? st.usage + (size_t)n * cells : nullptr;
            double* health = st.health ? st.health + (size_t)n * cells : nullptr;
            double* degrade = st.degrade ? st.degrade + (size_t)n * cells : nullptr;
            if (new_task) {
                const uint32_t episode = st.episode ? st.episode[n] : 0u;
                for (int k = threadIdx.x; k < cells; k += blockDim.x) {
                    if (usage) usage[k] = 0;
                    if (health) health[k] = 1.0;
                    if (degrade) {
                        double dg = 1.0;
                        if (degrade_in) {
                            dg = degrade_in[(size_t)n * cells + k];
                        } else if (cfg.b_degrade) {  // _random_health_statue (dmfb.py:157-164)
                            const uint4 r = env_random(seed, kStreamDegrade, cfg.env_base + n, episode, (uint32_t)k, 0u);
                            dg = u53(r.x, r.y) * 0.4 + 0.6;
                            if (u53(r.z, r.w) < 1.0 - cfg.per_degrade) dg = 1.0;
                        }
                        degrade[k] = dg;
                    }
                }
            } else if (usage) {
                for (int k = threadIdx.x; k < cells; k += blockDim.x) {
                    if (usage[k] > 50) {
                        if (health) health[k] = health[k] * (degrade ? degrade[k] : 1.0);
                        usage[k] = 0;
                    }
                }
            }
        }
    }
    if (obs == nullptr) return;
    paint_agents<FOV_T>(cfg, L, S, e_valid);
    int8_t* gobs = obs + (size_t)n0 * A * L.D;
    if (n_selected == e_valid) store_tile(gobs, S.tile, (uint32_t)(e_valid * A * L.D));
    else if (n_selected > 0) {
        // flag semantics for store_rows_masked: non-zero = store
        __syncthreads();
        if ((int)threadIdx.x < e_valid) S.flag[threadIdx.x] = (S.flag[threadIdx.x] == 0);
        store_rows_masked(gobs, S.tile, e_valid, A * L.D, S.flag);
    }
}

// ----------------------------------------------------------------- get_state --
// getglobalobs (dmfb.py:368-392): (3,W,L) per env, int8.  Tile of E2 envs staged in smem.
__global__ void __launch_bounds__(kThreads)
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
    }
    store_tile(out + (size_t)n0 * per_env, tile, (uint32_t)(e_valid * per_env));
}

int tile_envs_for(const dmfb_cfg_t& cfg)
{
    const int row = cfg.n_agents * cfg.obs_dim;
    int E = pick_tile_envs(row, 40 * 1024, 64);
    if (E > kThreads) E = kThreads;
    return E;
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
    if (!cfg || !st || st->n_envs < 0 || !st->drop || !st->step_count || !st->constraints || !st->terminated) {
        snprintf(g_last_error, sizeof(g_last_error), "null cfg/state pointer");
        return DMFB_ERR_BAD_ARG;
    }
    if (cfg->n_blocks != 0) {
        snprintf(g_last_error, sizeof(g_last_error), "n_blocks > 0 is not supported yet");
        return DMFB_ERR_BAD_ARG;
    }
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
    cfg->width = width; cfg->length = length; cfg->n_agents = n_agents; cfg->n_blocks = n_blocks; cfg->fov = fov;
    cfg->stall = stall ? 1 : 0; cfg->b_degrade = b_degrade ? 1 : 0; cfg->per_degrade = per_degrade;
    cfg->max_step = 2 * (width + length);
    cfg->n_actions = 5;
    cfg->obs_dim = 3 * fov * fov + 2;
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

#define DMFB_DISPATCH_FOV(fov, KERNEL, ...)            \
    switch (fov) {                                     \
    case 5: KERNEL<5> __VA_ARGS__; break;              \
    case 7: KERNEL<7> __VA_ARGS__; break;              \
    case 9: KERNEL<9> __VA_ARGS__; break;              \
    default: KERNEL<0> __VA_ARGS__; break;             \
    }

static int launch_reset(const dmfb_cfg_t* cfg, const dmfb_state_t* state, const uint8_t* mask, int mode, int new_task,
                        const uint8_t* layouts, const double* degrade, uint64_t seed, int8_t* obs, void* stream);

int dmfb_step(const dmfb_cfg_t* cfg, const dmfb_state_t* state, const void* actions, int action_elem_size,
              const double* u_inject, uint64_t seed, uint32_t flags, const dmfb_out_t* out, void* stream)
{
    int rc = check_common(cfg, state);
    if (rc) return rc;
    if (!actions || !out || !out->obs || (action_elem_size != 1 && action_elem_size != 4 && action_elem_size != 8)) {
        snprintf(g_last_error, sizeof(g_last_error), "dmfb_step: bad actions/out");
        return DMFB_ERR_BAD_ARG;
    }
    if (state->n_envs == 0) return DMFB_OK;
    const int E = tile_envs_for(*cfg);
    const TileLayout L(*cfg, E);
    const int grid = (state->n_envs + E - 1) / E;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
#define LAUNCH_STEP(F)                                                                                      \
    do {                                                                                                    \
        rc = set_smem(dmfb_step_kernel<F>, L.total);                                                        \
        if (rc) return rc;                                                                                  \
        dmfb_step_kernel<F><<<grid, kThreads, L.total, s>>>(*cfg, *state, actions, action_elem_size,        \
                                                            u_inject, seed, flags, *out, E);                \
    } while (0)
    switch (cfg->fov) {
    case 5: LAUNCH_STEP(5); break;
    case 7: LAUNCH_STEP(7); break;
    case 9: LAUNCH_STEP(9); break;
    default: LAUNCH_STEP(0); break;
    }
#undef LAUNCH_STEP
    g_launches.fetch_add(1);
    DMFB_CUDA_TRY(cudaGetLastError());
    return DMFB_OK;  // DMFB_STEP_AUTO_RESET is fused into the step kernel
}

static int launch_reset(const dmfb_cfg_t* cfg, const dmfb_state_t* state, const uint8_t* mask, int mode, int new_task,
                        const uint8_t* layouts, const double* degrade, uint64_t seed, int8_t* obs, void* stream)
{
    int rc = check_common(cfg, state);
    if (rc) return rc;
    if (state->n_envs == 0) return DMFB_OK;
    const int E = tile_envs_for(*cfg);
    const TileLayout L(*cfg, E);
    const int grid = (state->n_envs + E - 1) / E;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
#define LAUNCH_RESET(F)                                                                                     \
    do {                                                                                                    \
        rc = set_smem(dmfb_reset_kernel<F>, L.total);                                                       \
        if (rc) return rc;                                                                                  \
        dmfb_reset_kernel<F><<<grid, kThreads, L.total, s>>>(*cfg, *state, mask, mode, new_task, layouts,   \
                                                             degrade, seed, obs, E);                        \
    } while (0)
    switch (cfg->fov) {
    case 5: LAUNCH_RESET(5); break;
    case 7: LAUNCH_RESET(7); break;
    case 9: LAUNCH_RESET(9); break;
    default: LAUNCH_RESET(0); break;
    }
#undef LAUNCH_RESET
    g_launches.fetch_add(1);
    DMFB_CUDA_TRY(cudaGetLastError());
    return DMFB_OK;
}

int dmfb_reset(const dmfb_cfg_t* cfg, const dmfb_state_t* state, const uint8_t* mask, int new_task,
               const uint8_t* layouts, const double* degrade, uint64_t seed, int8_t* obs, void* stream)
{
    return launch_reset(cfg, state, mask, 0, new_task, layouts, degrade, seed, obs, stream);
}

int dmfb_restart(const dmfb_cfg_t* cfg, const dmfb_state_t* state, const uint8_t* mask, int8_t* obs, void* stream)
{
    if (state && !state->start) {
        snprintf(g_last_error, sizeof(g_last_error), "dmfb_restart needs state->start");
        return DMFB_ERR_BAD_ARG;
    }
    return launch_reset(cfg, state, mask, 1, 0, nullptr, nullptr, 0, obs, stream);
}

int dmfb_observe(const dmfb_cfg_t* cfg, const dmfb_state_t* state, int8_t* obs, void* stream)
{
    if (!obs) return DMFB_ERR_BAD_ARG;
    return launch_reset(cfg, state, nullptr, 2, 0, nullptr, nullptr, 0, obs, stream);
}

int dmfb_global_state(const dmfb_cfg_t* cfg, const dmfb_state_t* state, int8_t* out, void* stream)
{
    int rc = check_common(cfg, state);
    if (rc) return rc;
    if (!out) return DMFB_ERR_BAD_ARG;
    if (state->n_envs == 0) return DMFB_OK;
    const int per_env = 3 * cfg->width * cfg->length;
    const int E2 = pick_tile_envs(per_env, 32 * 1024, 64);
    const uint32_t tile_bytes = ((uint32_t)(E2 * per_env) + 15u) & ~15u;
    rc = set_smem(dmfb_global_state_kernel, tile_bytes);
    if (rc) return rc;
    const int grid = (state->n_envs + E2 - 1) / E2;
    dmfb_global_state_kernel<<<grid, kThreads, tile_bytes, static_cast<cudaStream_t>(stream)>>>(*cfg, *state, out, E2,
                                                                                                 tile_bytes);
    g_launches.fetch_add(1);
    DMFB_CUDA_TRY(cudaGetLastError());
    return DMFB_OK;
}

}  // extern "C"
