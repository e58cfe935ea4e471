/*
 * meda_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Scalar CPU restatement of the MEDA environment step of jesselasse/MARL-DMFB
 * (env/MEDA/meda.py).  Same role and rules as dmfb_oracle.c: only tests/,
 * __graft_entry__.smoke() and bench.py's CPU legs may use it.
 *
 * Parity status: PINNED by tests/test_oracle_golden.py against the meda_*.npz
 * traces recorded from the unmodified reference (tests/golden/make_golden.py).
 *
 * Coordinates follow the reference: x runs along `length`, y along `width`;
 * grids are [width][length] indexed [y][x] (meda.py:302-309).  A droplet is a
 * (2r+1)^2 square, r = 2 (meda.py:150,208-211), identified by its centre.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct orc_meda_cfg {
    int32_t width, length, n_agents, fov, b_degrade, obs_version; /* 0 = MEDAEnv, 1 = MEDAEnv_v0_1, 2 = MEDAEnv_v0_2 */
} orc_meda_cfg;

#define MAXA 64
#define RAD 2

static int meda_obs_dim(const orc_meda_cfg *c)
{
    return (c->obs_version == 2 ? 3 : 4) * c->fov * c->fov + 2;
}

/* Droplet.move (meda.py:106-138): step 3 on the axes, 2 on the diagonals, then pushed back on chip
 * (x against `length`, y against `width`). */
static void meda_move(int *xc, int *yc, int action, int width, int length)
{
    const int r = 3;
    if (action == 8) return;                 /* STALL */
    switch (action) {
    case 0: *yc -= r; break;                 /* N  */
    case 1: *xc += r; break;                 /* E  */
    case 2: *yc += r; break;                 /* S  */
    case 3: *xc -= r; break;                 /* W  */
    case 4: *xc += r - 1; *yc -= r - 1; break; /* NE */
    case 5: *xc += r - 1; *yc += r - 1; break; /* SE */
    case 6: *xc -= r - 1; *yc += r - 1; break; /* SW */
    case 7: *xc -= r - 1; *yc -= r - 1; break; /* NW */
    default: break;                          /* any other value: no branch taken in the reference */
    }
    if (*xc + RAD >= length) *xc += length - 1 - (*xc + RAD);
    else if (*xc - RAD < 0) *xc += 0 - (*xc - RAD);
    if (*yc + RAD >= width) *yc += width - 1 - (*yc + RAD);
    else if (*yc - RAD < 0) *yc += 0 - (*yc - RAD);
}

/* RoutingTaskManager.getMoveProb (meda.py:302-309): sequential float64 mean over the footprint */
static double meda_move_prob(const double *health, int length, int xc, int yc)
{
    double prob = 0.0;
    int count = 0;
    for (int y = yc - RAD; y <= yc + RAD; y++)
        for (int x = xc - RAD; x <= xc + RAD; x++) {
            prob += health[(size_t)y * length + x];
            count++;
        }
    return prob / (double)count;
}

static int clipi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* CPython set iteration order of a set of small non-negative ints inserted in ascending order
 * (meda.py:862-872: `observed.add(idx)` ... `for idx in observed`).  Emulates setobject.c:
 * open addressing, table of 8 slots growing to the next power of two > 4*used when fill*5 >= mask*3,
 * hash(i) = i, 9 linear probes only where i+9 <= mask, then i = (5i + 1 + perturb) & mask.  With fewer than 5 elements the table
 * keeps 8 slots and an index >= 8 wraps around (e.g. {0,3,8} iterates 0,8,3). */
static int cpython_set_order(uint32_t mask_bits, int n_max, int *order)
{
    int table[64];
    int size = 8, fill = 0;
    for (int k = 0; k < 64; k++) table[k] = -1;
    for (int v = 0; v < n_max; v++) {
        if (!((mask_bits >> v) & 1u)) continue;
        /* insert v */
        {
            size_t m = (size_t)size - 1, perturb = (size_t)v, i = (size_t)v & m;
            for (;;) {
                int placed = 0;
                const size_t lim = (i + 9 <= m) ? 9 : 0;   /* LINEAR_PROBES only when they fit below the mask */
                for (size_t p = 0; p <= lim; p++)
                    if (table[i + p] < 0) { table[i + p] = v; placed = 1; break; }
                if (placed) break;
                perturb >>= 5;
                i = (i * 5 + 1 + perturb) & m;
            }
            fill++;
        }
        if ((size_t)fill * 5 >= ((size_t)size - 1) * 3) {
            /* resize: new size = smallest power of two > used*4 (used <= 50000), reinsert in slot order */
            int old[64], oldsize = size, newsize = 8;
            memcpy(old, table, sizeof(old));
            while (newsize <= fill * 4) newsize <<= 1;
            size = newsize;
            for (int k = 0; k < 64; k++) table[k] = -1;
            for (int k = 0; k < oldsize; k++) {
                int w = old[k];
                if (w < 0) continue;
                size_t m = (size_t)size - 1, perturb = (size_t)w, i = (size_t)w & m;
                for (;;) {
                    int placed = 0;
                    const size_t lim = (i + 9 <= m) ? 9 : 0;
                    for (size_t p = 0; p <= lim; p++)
                        if (table[i + p] < 0) { table[i + p] = w; placed = 1; break; }
                    if (placed) break;
                    perturb >>= 5;
                    i = (i * 5 + 1 + perturb) & m;
                }
            }
        }
    }
    int n = 0;
    for (int k = 0; k < size; k++) if (table[k] >= 0) order[n++] = table[k];
    return n;
}

/* exported for the host layer's self-check of its python-built table */
int orc_cpython_set_order(uint32_t mask_bits, int n_max, int *order) { return cpython_set_order(mask_bits, n_max, order); }

/* one footprint of MEDAEnv.getOneObs: cells of the square centred (xc,yc), clipped to the chip
 * (meda.py:627-630), written into `layer` either where they fall inside the window (:633-635) or
 * clipped onto it (:669-671) */
static void meda_foot(int8_t *layer, int fov, int ox, int oy, int xc, int yc, int W, int L, int val, int clip_to_window)
{
    const int y_min = yc - RAD < 0 ? 0 : yc - RAD, y_max = yc + RAD >= W ? W - 1 : yc + RAD;
    const int x_min = xc - RAD < 0 ? 0 : xc - RAD, x_max = xc + RAD >= L ? L - 1 : xc + RAD;
    for (int y = y_min; y <= y_max; y++)
        for (int x = x_min; x <= x_max; x++) {
            int nx = x - ox, ny = y - oy;
            if (clip_to_window) {
                nx = clipi(nx, 0, fov - 1);
                ny = clipi(ny, 0, fov - 1);
                layer[ny * fov + nx] = (int8_t)val;
            } else if (0 <= nx && nx < fov && 0 <= ny && ny < fov) {
                layer[ny * fov + nx] = (int8_t)val;
            }
        }
}

/* MEDAEnv.getOneObs (meda.py:613-674): (4,fov,fov)+2, float64 in the reference, integral values -> int8 */
static void meda_obs_base(const orc_meda_cfg *c, const int *xc, const int *yc, const int *gx, const int *gy,
                          int agent, int8_t *obs)
{
    const int fov = c->fov, f2 = fov * fov, n = c->n_agents, W = c->width, L = c->length;
    memset(obs, 0, (size_t)(4 * f2 + 2));
    const int cx = xc[agent], cy = yc[agent];
    const int ox = cx - fov / 2, oy = cy - fov / 2;
    meda_foot(obs + 0 * f2, fov, ox, oy, cx, cy, W, L, agent + 1, 0);                  /* own droplet  (:626-635) */
    meda_foot(obs + 1 * f2, fov, ox, oy, gx[agent], gy[agent], W, L, agent + 1, 0);    /* own goal     (:637-646) */
    for (int idx = 0; idx < n; idx++)                                                  /* others       (:648-658) */
        if (idx != agent) meda_foot(obs + 2 * f2, fov, ox, oy, xc[idx], yc[idx], W, L, idx + 1, 0);
    for (int idx = 0; idx < n; idx++)                                                  /* others' goals, clipped (:660-671) */
        if (idx != agent) meda_foot(obs + 3 * f2, fov, ox, oy, gx[idx], gy[idx], W, L, idx + 1, 1);
    obs[4 * f2 + 0] = (int8_t)(gx[agent] - cx);                                        /* dir_VEC (:672) */
    obs[4 * f2 + 1] = (int8_t)(gy[agent] - cy);
}

/* MEDAEnv_v0_2.getOneObs (meda.py:850-897) and MEDAEnv_v0_1.getOneObs (meda.py:788-844).  v0_1 is v0_2 with one
 * more layer (own goal, index 1, :808-816) in front of the others' goals and with the direction entries
 * (dy / width, dx / length) as float64 (:840); their integer numerators (dy, dx) are emitted here. */
static void meda_obs_v0x(const orc_meda_cfg *c, const int *xc, const int *yc, const int *gx, const int *gy,
                         int agent, int8_t *obs)
{
    const int fov = c->fov, f2 = fov * fov, n = c->n_agents, W = c->width, L = c->length, hf = fov / 2;
    const int v1 = c->obs_version == 1;
    const int l_goals = v1 ? 2 : 1, l_border = v1 ? 3 : 2, n_layers = v1 ? 4 : 3;
    memset(obs, 0, (size_t)(n_layers * f2 + 2));
    const int cx = xc[agent], cy = yc[agent];
    const int ox = cx - hf, oy = cy - hf;
    uint32_t observed = 0;
    for (int idx = 0; idx < n; idx++)                       /* layer 0 (:863-869), footprints not clipped to the chip */
        for (int y = yc[idx] - RAD; y <= yc[idx] + RAD; y++)
            for (int x = xc[idx] - RAD; x <= xc[idx] + RAD; x++) {
                int nx = x - ox, ny = y - oy;
                if (0 <= nx && nx < fov && 0 <= ny && ny < fov) {
                    obs[0 * f2 + ny * fov + nx] = (int8_t)(idx + 1);
                    observed |= 1u << idx;
                }
            }
    if (v1)                                                 /* own goal, clipped to the chip, where inside the window */
        meda_foot(obs + 1 * f2, fov, ox, oy, gx[agent], gy[agent], W, L, agent + 1, 0);
    int order[MAXA];
    const int no = cpython_set_order(observed, n, order);   /* `for idx in observed` (:871-872) */
    for (int k = 0; k < no; k++) {
        const int idx = order[k];
        if (idx == agent) continue;                         /* observed.remove(agent_index) */
        for (int y = gy[idx] - RAD; y <= gy[idx] + RAD; y++)
            for (int x = gx[idx] - RAD; x <= gx[idx] + RAD; x++) {
                int nx = clipi(x - ox, 0, fov - 1), ny = clipi(y - oy, 0, fov - 1);
                obs[l_goals * f2 + ny * fov + nx] = (int8_t)(idx + 1);
            }
    }
    /* border layer (:880-891 / :826-839): the reference indexes the ROW axis with the x-derived bounds and uses
     * `width` for the x extent and `length` for the y extent; python slicing clamps oversize bounds */
    int8_t *bl = obs + l_border * f2;
    int leftbound = hf - cx, rightbound = hf - (W - 1 - cx);
    if (leftbound > 0) {
        for (int r = 0; r < leftbound && r < fov; r++) for (int q = 0; q < fov; q++) bl[r * fov + q] = 1;
    } else if (rightbound > 0) {
        for (int r = (fov - rightbound < 0 ? 0 : fov - rightbound); r < fov; r++)
            for (int q = 0; q < fov; q++) bl[r * fov + q] = 1;
    }
    int upbound = hf - cy, downbound = hf - (L - 1 - cy);
    if (upbound > 0) {
        for (int r = 0; r < fov; r++) for (int q = 0; q < upbound && q < fov; q++) bl[r * fov + q] = 1;
    } else if (downbound > 0) {
        for (int r = 0; r < fov; r++)
            for (int q = (fov - downbound < 0 ? 0 : fov - downbound); q < fov; q++) bl[r * fov + q] = 1;
    }
    if (v1) {
        obs[4 * f2 + 0] = (int8_t)(gy[agent] - cy);
        obs[4 * f2 + 1] = (int8_t)(gx[agent] - cx);
    } else {
        /* direction vector (:895): round((dy)/(width/30)), round((dx)/(length/30)) */
        obs[3 * f2 + 0] = (int8_t)(int)rint((double)(gy[agent] - cy) / ((double)W / 30.0));
        obs[3 * f2 + 1] = (int8_t)(int)rint((double)(gx[agent] - cx) / ((double)L / 30.0));
    }
}

static void meda_load(const orc_meda_cfg *c, const uint8_t *drop, int *xc, int *yc, int *gx, int *gy)
{
    for (int i = 0; i < c->n_agents; i++) {
        xc[i] = drop[4 * i + 0]; yc[i] = drop[4 * i + 1]; gx[i] = drop[4 * i + 2]; gy[i] = drop[4 * i + 3];
    }
}

static void meda_all_obs(const orc_meda_cfg *c, const uint8_t *drop, int8_t *obs)
{
    int xc[MAXA], yc[MAXA], gx[MAXA], gy[MAXA];
    meda_load(c, drop, xc, yc, gx, gy);
    const int D = meda_obs_dim(c);
    for (int i = 0; i < c->n_agents; i++) {
        if (c->obs_version != 0) meda_obs_v0x(c, xc, yc, gx, gy, i, obs + (size_t)i * D);
        else meda_obs_base(c, xc, yc, gx, gy, i, obs + (size_t)i * D);
    }
}

/* MEDAEnv.step (meda.py:513-539) -> moveDroplets (:241-259) -> moveOneDroplet (:261-292) -> calPunish (:321-330)
 * for ONE chip.  `fails` is kept as the punish COUNT (the reference keeps the float sum of -0.6's). */
static void meda_step_one(const orc_meda_cfg *c, uint8_t *drop, uint8_t *status, int32_t *step_count, int32_t *fails,
                          double *usage, const double *health, const int8_t *actions, const double *u,
                          int8_t *obs, double *reward, uint8_t *done, int32_t *punish_count_out, uint8_t *success_out)
{
    const int n = c->n_agents, W = c->width, L = c->length;
    int xc[MAXA], yc[MAXA], gx[MAXA], gy[MAXA], pun[MAXA];
    double rew[MAXA];
    meda_load(c, drop, xc, yc, gx, gy);
    *step_count += 1;                                            /* :514 */
    int success = 0;
    for (int i = 0; i < n; i++) {
        if (status[i]) { rew[i] = 0.0; continue; }               /* :248-249 */
        const int dx0 = xc[i] - gx[i], dy0 = yc[i] - gy[i];
        const int old2 = dx0 * dx0 + dy0 * dy0;                  /* distances[i]^2 */
        if (old2 < 16) {                                         /* distances[i] < r_i + r_goal = 4 (:272-277) */
            xc[i] = gx[i]; yc[i] = gy[i];
            rew[i] = 0.0;
            status[i] = 1;
        } else {
            const double prob = health ? meda_move_prob(health, L, xc[i], yc[i]) : 1.0;  /* :279 */
            const double draw = u ? u[i] : 0.0;
            if (draw <= prob) meda_move(&xc[i], &yc[i], actions[i], W, L);              /* :280-281 */
            const int dx1 = xc[i] - gx[i], dy1 = yc[i] - gy[i];
            const int new2 = dx1 * dx1 + dy1 * dy1;
            if (new2 < 16) rew[i] = 0.0;                         /* :283-290, comparisons on exact squares */
            else if (new2 == old2 && actions[i] == 8) rew[i] = -0.2;
            else if (new2 < old2) rew[i] = -0.08;
            else rew[i] = -0.4;
        }
    }
    /* calPunish (:321-330): centre distance < 1.5*(r_i+r_j) = 6 */
    int total = 0;
    for (int i = 0; i < n; i++) pun[i] = 0;
    for (int i = 0; i < n - 1; i++)
        for (int j = i + 1; j < n; j++) {
            const int dx = xc[i] - xc[j], dy = yc[i] - yc[j];
            if (dx * dx + dy * dy < 36) { pun[i]++; pun[j]++; total += 2; }
        }
    for (int i = 0; i < n; i++) {
        double p = 0.0;
        for (int k = 0; k < pun[i]; k++) p -= 0.6;              /* punish[i] -= 0.6, repeated */
        rew[i] = pun[i] ? rew[i] + p : rew[i];                   /* rewards[i] += punish[i] (int 0 when untouched) */
    }
    *fails += total;                                             /* :521 */
    int all = 1;
    for (int i = 0; i < n; i++) all &= status[i];
    if (all) {                                                   /* :522-525 */
        for (int i = 0; i < n; i++) rew[i] = rew[i] + 3.0;
        if (*fails == 0) for (int i = 0; i < n; i++) rew[i] = rew[i] + 3.0;
    }
    for (int i = 0; i < n; i++) { drop[4 * i + 0] = (uint8_t)xc[i]; drop[4 * i + 1] = (uint8_t)yc[i]; }
    if (obs) meda_all_obs(c, drop, obs);                         /* :528 */
    if (*step_count < W + L) {                                   /* max_step = w + l (:492,529-534) */
        if (all && *fails == 0) success = 1;
        for (int i = 0; i < n; i++) done[i] = status[i];
        if (usage)                                               /* addUsage (:591-598) */
            for (int i = 0; i < n; i++)
                if (!done[i])
                    for (int y = yc[i] - RAD; y <= yc[i] + RAD; y++)
                        for (int x = xc[i] - RAD; x <= xc[i] + RAD; x++) usage[(size_t)y * L + x] += 1.0;
    } else {
        for (int i = 0; i < n; i++) done[i] = 1;
    }
    for (int i = 0; i < n; i++) reward[i] = rew[i];
    *punish_count_out = total;
    *success_out = (uint8_t)success;
}

static void meda_update_health(const orc_meda_cfg *c, double *usage, double *health, const double *degrade)
{
    if (!c->b_degrade) return;                                   /* meda.py:600-602 */
    const int cells = c->width * c->length;
    for (int k = 0; k < cells; k++)
        if (usage[k] > 50.0) { health[k] = health[k] * degrade[k]; usage[k] = 0.0; }
}

int orc_meda_step(const orc_meda_cfg *c, int n_envs, uint8_t *drop, uint8_t *status, int32_t *step_count, int32_t *fails,
                  double *usage, const double *health, const int8_t *actions, const double *u, int8_t *obs,
                  double *reward, uint8_t *done, int32_t *punish_count, uint8_t *success)
{
    const int A = c->n_agents, cells = c->width * c->length, D = meda_obs_dim(c);
    for (int e = 0; e < n_envs; e++)
        meda_step_one(c, drop + (size_t)e * A * 4, status + (size_t)e * A, step_count + e, fails + e,
                      usage ? usage + (size_t)e * cells : NULL, health ? health + (size_t)e * cells : NULL,
                      actions + (size_t)e * A, u ? u + (size_t)e * A : NULL, obs ? obs + (size_t)e * A * D : NULL,
                      reward + (size_t)e * A, done + (size_t)e * A, punish_count + e, success + e);
    return 0;
}

/* MEDAEnv.reset (meda.py:541-550) with the tasks injected: counters zeroed, new tasks, obs, then updateHealth */
int orc_meda_reset(const orc_meda_cfg *c, int n_envs, const uint8_t *mask, const uint8_t *layouts, uint8_t *drop,
                   uint8_t *status, int32_t *step_count, int32_t *fails, double *usage, double *health,
                   const double *degrade, int8_t *obs)
{
    const int A = c->n_agents, cells = c->width * c->length, D = meda_obs_dim(c);
    for (int e = 0; e < n_envs; e++) {
        if (mask && !mask[e]) continue;
        step_count[e] = 0;
        fails[e] = 0;
        memcpy(drop + (size_t)e * A * 4, layouts + (size_t)e * A * 4, (size_t)A * 4);
        memset(status + (size_t)e * A, 0, (size_t)A);
        if (obs) meda_all_obs(c, drop + (size_t)e * A * 4, obs + (size_t)e * A * D);
        if (usage && health && degrade)
            meda_update_health(c, usage + (size_t)e * cells, health + (size_t)e * cells, degrade + (size_t)e * cells);
    }
    return 0;
}

int orc_meda_observe(const orc_meda_cfg *c, int n_envs, const uint8_t *drop, int8_t *obs)
{
    const int A = c->n_agents, D = meda_obs_dim(c);
    for (int e = 0; e < n_envs; e++) meda_all_obs(c, drop + (size_t)e * A * 4, obs + (size_t)e * A * D);
    return 0;
}

/* ---------------------------------------------------- task generator -- */
static inline uint64_t meda_splitmix(uint64_t *s)
{
    uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

/* refresh/addTask/_genLegalDroplet (meda.py:161-185,213-233): centres uniform in [r, dim-r-1]; a droplet
 * (destination) is redrawn while its centre is closer than 1.5*(2+2+2) = 9 to an earlier droplet
 * (destination); the destination is redrawn while it overlaps its own droplet. */
void orc_meda_gen_layout(const orc_meda_cfg *c, uint64_t *rng, uint8_t *layout)
{
    const int A = c->n_agents, W = c->width, L = c->length;
    int xs[MAXA], ys[MAXA], gxs[MAXA], gys[MAXA];
    for (int i = 0; i < A; i++) {
        for (;;) {  /* _genLegalDroplet(self.droplets) */
            int y = RAD + (int)(meda_splitmix(rng) % (uint64_t)(W - 2 * RAD));
            int x = RAD + (int)(meda_splitmix(rng) % (uint64_t)(L - 2 * RAD));
            int ok = 1;
            for (int j = 0; j < i; j++) { int dx = x - xs[j], dy = y - ys[j]; if (dx * dx + dy * dy < 81) { ok = 0; break; } }
            if (ok) { xs[i] = x; ys[i] = y; break; }
        }
        for (;;) {  /* _genLegalDroplet(self.destinations), repeated while it overlaps the droplet (:179-182) */
            int y = RAD + (int)(meda_splitmix(rng) % (uint64_t)(W - 2 * RAD));
            int x = RAD + (int)(meda_splitmix(rng) % (uint64_t)(L - 2 * RAD));
            int ok = 1;
            for (int j = 0; j < i; j++) { int dx = x - gxs[j], dy = y - gys[j]; if (dx * dx + dy * dy < 81) { ok = 0; break; } }
            if (!ok) continue;
            if (abs(x - xs[i]) <= 2 * RAD && abs(y - ys[i]) <= 2 * RAD) continue;  /* isDropletOverlap */
            gxs[i] = x; gys[i] = y;
            break;
        }
    }
    for (int i = 0; i < A; i++) {
        layout[4 * i + 0] = (uint8_t)xs[i]; layout[4 * i + 1] = (uint8_t)ys[i];
        layout[4 * i + 2] = (uint8_t)gxs[i]; layout[4 * i + 3] = (uint8_t)gys[i];
    }
}

/* CPU-baseline driver, see orc_dmfb_rollout */
typedef struct {
    const orc_meda_cfg *c;
    int e0, e1, steps;
    uint64_t seed, total;
    int8_t *obs;
} meda_roll_job;

static void *meda_roll_thread(void *arg)
{
    meda_roll_job *job = (meda_roll_job *)arg;
    const orc_meda_cfg *c = job->c;
    const int A = c->n_agents, cells = c->width * c->length, D = meda_obs_dim(c);
    double *usage = (double *)calloc((size_t)cells, sizeof(double));
    double *health = (double *)malloc((size_t)cells * sizeof(double));
    double *degrade = (double *)malloc((size_t)cells * sizeof(double));
    uint64_t total = 0;
    for (int e = job->e0; e < job->e1; e++) {
        uint64_t rng = job->seed * 0x100000001B3ull + (uint64_t)e;
        uint8_t drop[4 * MAXA], status[MAXA], done[MAXA], succ;
        int8_t acts[MAXA];
        double u[MAXA], rew[MAXA];
        int32_t sc = 0, fails = 0, cons;
        int8_t *obs = job->obs + (size_t)e * A * D;
        for (int k = 0; k < cells; k++) {
            usage[k] = 0.0; health[k] = 1.0;
            degrade[k] = c->b_degrade ? (double)(meda_splitmix(&rng) >> 11) * (1.0 / 9007199254740992.0) * 0.4 + 0.6 : 1.0;
        }
        memset(status, 0, sizeof(status));
        orc_meda_gen_layout(c, &rng, drop);
        for (int t = 0; t < job->steps; t++) {
            for (int i = 0; i < A; i++) {
                acts[i] = (int8_t)(meda_splitmix(&rng) % 9u);
                u[i] = (double)(meda_splitmix(&rng) >> 11) * (1.0 / 9007199254740992.0);
            }
            meda_step_one(c, drop, status, &sc, &fails, usage, c->b_degrade ? health : NULL, acts, u, obs, rew, done,
                          &cons, &succ);
            int all = 1;
            for (int i = 0; i < A; i++) { all &= done[i]; total += (uint64_t)(rew[i] < 0.0); }
            if (all) {
                sc = 0; fails = 0;
                memset(status, 0, sizeof(status));
                orc_meda_gen_layout(c, &rng, drop);
                meda_all_obs(c, drop, obs);
                meda_update_health(c, usage, health, degrade);
            }
        }
    }
    free(usage); free(health); free(degrade);
    job->total = total;
    return NULL;
}

int64_t orc_meda_rollout(const orc_meda_cfg *c, int n_envs, int steps, uint64_t seed, int8_t *obs, int n_threads,
                         uint64_t *checksum)
{
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 1024) n_threads = 1024;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
    meda_roll_job *jobs = (meda_roll_job *)malloc(sizeof(meda_roll_job) * (size_t)n_threads);
    for (int t = 0; t < n_threads; t++) {
        jobs[t].c = c; jobs[t].steps = steps; jobs[t].seed = seed; jobs[t].obs = obs; jobs[t].total = 0;
        jobs[t].e0 = (int)((int64_t)n_envs * t / n_threads);
        jobs[t].e1 = (int)((int64_t)n_envs * (t + 1) / n_threads);
        pthread_create(&th[t], NULL, meda_roll_thread, &jobs[t]);
    }
    uint64_t total = 0;
    for (int t = 0; t < n_threads; t++) { pthread_join(th[t], NULL); total += jobs[t].total; }
    free(th); free(jobs);
    if (checksum) *checksum = total;
    return (int64_t)n_envs * steps * c->n_agents;
}
