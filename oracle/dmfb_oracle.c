/*
 * dmfb_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Scalar CPU restatement of the DMFB environment step of jesselasse/MARL-DMFB
 * (env/DMFB/dmfb.py).  It exists only so that tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference leg can check (and time) the
 * CUDA path against something that follows the reference line by line.  The
 * product (marl-dmfb_b200/) never links, imports or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py replays every trace under
 * tests/golden/ (npz files) — recorded from the unmodified Python reference by
 * tests/golden/make_golden.py — through this file and requires bit equality
 * of positions, observations, dones, constraints, success, usage, health and
 * of the float64 rewards.
 *
 * Each function cites the reference lines it restates (paths relative to the
 * reference repo root).  One chip is handled at a time with the same loop
 * structure as the reference; batching is a plain loop over chips (optionally
 * OpenMP for the timing legs).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct orc_dmfb_cfg {
    int32_t width, length, n_agents, fov, stall, b_degrade, n_blocks;
    int32_t obs_version; /* 0 = DMFBenv.getOneObs, 1 = DMFBenv_v0_1.getOneObs */
} orc_dmfb_cfg;

static int orc_obs_dim(const orc_dmfb_cfg *c) { return (c->obs_version == 1 ? 4 : 3) * c->fov * c->fov + 2; }

#define MAXA 64

/* Droplet.move + Action (dmfb.py:26-31,103-124) */
static int orc_move(int *x, int *y, int action, int width, int length)
{
    switch (action) {
    case 0: break;            /* STALL */
    case 4: *y += 1; break;   /* UP    */
    case 3: *y -= 1; break;   /* DOWN  */
    case 2: *x -= 1; break;   /* LEFT  */
    case 1: *x += 1; break;   /* RIGHT */
    default: return -1;       /* TypeError('action is illegal') */
    }
    if (*x > width - 1) *x = width - 1; else if (*x < 0) *x = 0;
    if (*y > length - 1) *y = length - 1; else if (*y < 0) *y = 0;
    return 0;
}

/* RoutingTaskManager._isTouchingBlocks (dmfb.py:301-308); blocks = [n_blocks][2] (x_min, y_min), 2x2 each (:246) */
static int orc_touching_blocks(const uint8_t *blocks, int n_blocks, int x, int y)
{
    for (int b = 0; b < n_blocks; b++) {
        int x_min = blocks[2 * b], y_min = blocks[2 * b + 1];
        if (x >= x_min && x <= x_min + 1 && y >= y_min && y <= y_min + 1) return 1;
    }
    return 0;
}

/* RoutingTaskManager._isinvalidaction (dmfb.py:310-323): any pair of droplets
 * at squared distance 0 (the Gram-matrix EDM of :200-205 is exact on small ints). */
static int orc_any_pair_equal(const int *x, const int *y, int n)
{
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++)
            if (i != j) {
                int dx = x[i] - x[j], dy = y[i] - y[j];
                if (dx * dx + dy * dy == 0) return 1;
            }
    return 0;
}

/* RoutingTaskManager.getOneObs (dmfb.py:395-457) + DMFBenv.getOneObs (:614-620):
 * int8 (3,fov,fov) flattened, then the 2-byte direction vector. */
static void orc_one_obs(const orc_dmfb_cfg *c, const int *x, const int *y, const int *gx, const int *gy,
                        const uint8_t *blocks, int agent, int8_t *obs)
{
    const int fov = c->fov, hf = fov / 2, n = c->n_agents;
    const int f2 = fov * fov;
    memset(obs, 0, (size_t)(3 * f2 + 2));
    const int cx = x[agent], cy = y[agent];
    const int ox = cx - fov / 2, oy = cy - fov / 2;
    /* layer 0: every droplet inside the window (:410-413) */
    for (int idx = 0; idx < n; idx++) {
        int rx = x[idx] - ox, ry = y[idx] - oy;
        if (0 <= rx && rx < fov && 0 <= ry && ry < fov) obs[0 * f2 + rx * fov + ry] = (int8_t)(idx + 1);
    }
    /* layer 1: goals of the other visible droplets, clipped into the window (:416-420);
     * abs(d) < fov/2 is a true division in the reference -> compare 2*abs(d) < fov */
    for (int idx = 0; idx < n; idx++) {
        if (idx != agent && 2 * abs(x[idx] - cx) < fov && 2 * abs(y[idx] - cy) < fov) {
            int rx = gx[idx] - ox, ry = gy[idx] - oy;
            if (rx < 0) rx = 0; if (rx > fov - 1) rx = fov - 1;
            if (ry < 0) ry = 0; if (ry > fov - 1) ry = fov - 1;
            obs[1 * f2 + rx * fov + ry] = (int8_t)(idx + 1);
        }
    }
    /* layer 2: blocks at ABSOLUTE chip coordinates, as the reference writes them (:422-426) ... */
    for (int b = 0; blocks && b < c->n_blocks; b++)
        for (int i = blocks[2 * b]; i <= blocks[2 * b] + 1; i++)
            for (int j = blocks[2 * b + 1]; j <= blocks[2 * b + 1] + 1; j++)
                if (0 <= i && i < fov && 0 <= j && j < fov) obs[2 * f2 + i * fov + j] = 1;
    /* ... + off-chip boundary (:428-439) */
    int leftbound = hf - cx, rightbound = hf - (c->width - 1 - cx);
    if (leftbound > 0) {
        for (int r = 0; r < leftbound && r < fov; r++)
            for (int q = 0; q < fov; q++) obs[2 * f2 + r * fov + q] = 1;
    } else if (rightbound > 0) {
        for (int r = (fov - rightbound < 0 ? 0 : fov - rightbound); r < fov; r++)
            for (int q = 0; q < fov; q++) obs[2 * f2 + r * fov + q] = 1;
    }
    int upbound = hf - cy, downbound = hf - (c->length - 1 - cy);
    if (upbound > 0) {
        for (int r = 0; r < fov; r++)
            for (int q = 0; q < upbound && q < fov; q++) obs[2 * f2 + r * fov + q] = 1;
    } else if (downbound > 0) {
        for (int r = 0; r < fov; r++)
            for (int q = (fov - downbound < 0 ? 0 : fov - downbound); q < fov; q++) obs[2 * f2 + r * fov + q] = 1;
    }
    /* direction vector (:442-454): python round() == round-half-even on the double == rint() */
    int drx = gx[agent] - cx, dry = gy[agent] - cy;
    if (abs(drx) > hf) {
        double scale = (double)(c->width - hf) / (double)(10 - hf);
        if (drx > 0) drx = (int)rint((double)(drx - hf) / scale) + hf;
        else drx = (int)rint((double)(drx + hf) / scale) - hf;
    }
    if (abs(dry) > hf) {
        double scale = (double)(c->length - hf) / (double)(10 - hf);
        if (dry > 0) dry = (int)rint((double)(dry - hf) / scale) + hf;
        else dry = (int)rint((double)(dry + hf) / scale) - hf;
    }
    obs[3 * f2 + 0] = (int8_t)drx;
    obs[3 * f2 + 1] = (int8_t)dry;
}

/* DMFBenv_v0_1.getOneObs (dmfb.py:727-835; `--version 0.1`, common/config.py:6-8): float64 (4,fov,fov) + 2 in the
 * reference; the layers are integral and are emitted as int8, the two direction entries
 * ((tar_y - cy) / length, (tar_x - cx) / width) (:833) are emitted as their integer numerators. */
static void orc_one_obs_v01(const orc_dmfb_cfg *c, const int *x, const int *y, const int *gx, const int *gy,
                            const uint8_t *blocks, int agent, int8_t *obs)
{
    const int fov = c->fov, hf = fov / 2, n = c->n_agents, f2 = fov * fov;
    memset(obs, 0, (size_t)(4 * f2 + 2));
    const int cx = x[agent], cy = y[agent];
    const int ox = cx - hf, oy = cy - hf;
    /* layer 0 (:741-747): every droplet inside the window; the OTHER visible ones are remembered */
    int see_idx[MAXA], see_x[MAXA], see_y[MAXA], see_dist[MAXA], n_see = 0;
    for (int idx = 0; idx < n; idx++) {
        int rx = x[idx] - ox, ry = y[idx] - oy;
        if (0 <= rx && rx < fov && 0 <= ry && ry < fov) {
            obs[0 * f2 + rx * fov + ry] = (int8_t)(idx + 1);
            if (idx != agent) {
                see_idx[n_see] = idx; see_x[n_see] = rx; see_y[n_see] = ry;
                see_dist[n_see] = abs(x[idx] - gx[idx]) + abs(y[idx] - gy[idx]);
                n_see++;
            }
        }
    }
    /* layer 1 (:751-761): own goal, projected onto the window for fewer than 10 droplets, else only if inside */
    {
        int rx = gx[agent] - ox, ry = gy[agent] - oy;
        if (n < 10) {
            rx = rx < 0 ? 0 : (rx > fov - 1 ? fov - 1 : rx);
            ry = ry < 0 ? 0 : (ry > fov - 1 ? fov - 1 : ry);
            obs[1 * f2 + rx * fov + ry] = (int8_t)(agent + 1);
        } else if (0 <= rx && rx < fov && 0 <= ry && ry < fov) {
            obs[1 * f2 + rx * fov + ry] = (int8_t)(agent + 1);
        }
    }
    /* layer 2 (:764-808): goals of the visible others, nearest-to-goal first (list.sort is stable), each drawn
     * where the ray droplet -> goal leaves the window; an occupied cell pushes the mark to a free 4-neighbour */
    for (int a = 1; a < n_see; a++) {                    /* stable insertion sort by distance */
        int ti = see_idx[a], tx = see_x[a], ty = see_y[a], td = see_dist[a], b = a - 1;
        while (b >= 0 && see_dist[b] > td) {
            see_idx[b + 1] = see_idx[b]; see_x[b + 1] = see_x[b]; see_y[b + 1] = see_y[b]; see_dist[b + 1] = see_dist[b];
            b--;
        }
        see_idx[b + 1] = ti; see_x[b + 1] = tx; see_y[b + 1] = ty; see_dist[b + 1] = td;
    }
    int8_t *l2 = obs + 2 * f2;
    for (int s = 0; s < n_see; s++) {
        const int idx = see_idx[s], sx = see_x[s], sy = see_y[s];
        const int dx = gx[idx] - x[idx], dy = gy[idx] - y[idx];
        const int boundx = dx >= 0 ? fov - 1 - sx : -sx;
        const int boundy = dy >= 0 ? fov - 1 - sy : -sy;
        int clipdx, clipdy;
        if (abs(dx) <= abs(boundx) && abs(dy) <= abs(boundy)) { clipdx = dx; clipdy = dy; }
        else if (dx == 0) { clipdx = 0; clipdy = boundy; }
        else if (dy == 0) { clipdx = boundx; clipdy = 0; }
        else {
            /* python: dx / dy * boundy == (dx / dy) * boundy in float64; dy * boundx / dx == (dy * boundx) / dx */
            const double qx = (double)dx / (double)dy * (double)boundy;
            const double qy = (double)(dy * boundx) / (double)dx;
            if (dx >= 0) { int v = (int)ceil(qx); clipdx = boundx < v ? boundx : v; }
            else { int v = (int)floor(qx); clipdx = boundx > v ? boundx : v; }
            if (dy >= 0) { int v = (int)ceil(qy); clipdy = boundy < v ? boundy : v; }
            else { int v = (int)floor(qy); clipdy = boundy > v ? boundy : v; }
        }
        const int i = sx + clipdx, j = sy + clipdy;
        /* 0 <= i, j < fov: the clipped offsets keep the sign of (dx, dy) and never pass the window bounds */
#define L2AT(I, J) l2[(I) * fov + (J)]
        if (L2AT(i, j) == 0) { L2AT(i, j) = (int8_t)(idx + 1); continue; }
        if (i == sx && j == sy) continue;
        if (i + 1 < fov && L2AT(i + 1, j) == 0) { L2AT(i + 1, j) = (int8_t)(idx + 1); continue; }
        if (i - 1 >= 0 && L2AT(i - 1, j) == 0) { L2AT(i - 1, j) = (int8_t)(idx + 1); continue; }
        if (j + 1 < fov && L2AT(i, j + 1) == 0) { L2AT(i, j + 1) = (int8_t)(idx + 1); continue; }
        if (j - 1 >= 0 && L2AT(i, j - 1) == 0) { L2AT(i, j - 1) = (int8_t)(idx + 1); continue; }
#undef L2AT
    }
    /* layer 3: blocks at ABSOLUTE chip coordinates (:813-817) + off-chip boundary (:819-831) */
    for (int b = 0; blocks && b < c->n_blocks; b++)
        for (int i = blocks[2 * b]; i <= blocks[2 * b] + 1; i++)
            for (int j = blocks[2 * b + 1]; j <= blocks[2 * b + 1] + 1; j++)
                if (0 <= i && i < fov && 0 <= j && j < fov) obs[3 * f2 + i * fov + j] = 1;
    int leftbound = hf - cx, rightbound = hf - (c->width - 1 - cx);
    if (leftbound > 0) {
        for (int r = 0; r < leftbound && r < fov; r++)
            for (int q = 0; q < fov; q++) obs[3 * f2 + r * fov + q] = 1;
    } else if (rightbound > 0) {
        for (int r = (fov - rightbound < 0 ? 0 : fov - rightbound); r < fov; r++)
            for (int q = 0; q < fov; q++) obs[3 * f2 + r * fov + q] = 1;
    }
    int upbound = hf - cy, downbound = hf - (c->length - 1 - cy);
    if (upbound > 0) {
        for (int r = 0; r < fov; r++)
            for (int q = 0; q < upbound && q < fov; q++) obs[3 * f2 + r * fov + q] = 1;
    } else if (downbound > 0) {
        for (int r = 0; r < fov; r++)
            for (int q = (fov - downbound < 0 ? 0 : fov - downbound); q < fov; q++) obs[3 * f2 + r * fov + q] = 1;
    }
    obs[4 * f2 + 0] = (int8_t)(gy[agent] - cy);          /* numerator of (tar_y - cy) / length */
    obs[4 * f2 + 1] = (int8_t)(gx[agent] - cx);          /* numerator of (tar_x - cx) / width  */
}

static void orc_load(const orc_dmfb_cfg *c, const uint8_t *drop, int *x, int *y, int *gx, int *gy)
{
    for (int i = 0; i < c->n_agents; i++) {
        x[i] = drop[4 * i + 0]; y[i] = drop[4 * i + 1];
        gx[i] = drop[4 * i + 2]; gy[i] = drop[4 * i + 3];
    }
}

/* DMFBenv.getObs (dmfb.py:622-626) for chip state `drop` */
static void orc_all_obs(const orc_dmfb_cfg *c, const uint8_t *drop, const uint8_t *blocks, int8_t *obs)
{
    int x[MAXA], y[MAXA], gx[MAXA], gy[MAXA];
    orc_load(c, drop, x, y, gx, gy);
    const int D = orc_obs_dim(c);
    for (int i = 0; i < c->n_agents; i++) {
        if (c->obs_version == 1) orc_one_obs_v01(c, x, y, gx, gy, blocks, i, obs + (size_t)i * D);
        else orc_one_obs(c, x, y, gx, gy, blocks, i, obs + (size_t)i * D);
    }
}

/* DMFBenv.step (dmfb.py:560-587) -> RoutingTaskManager.moveDroplets (:253-299) ->
 * moveOneDroplet (:325-359) -> addUsage (:459-463), for ONE chip.
 * Returns -1 on an illegal action. */
static int orc_step_one(const orc_dmfb_cfg *c, uint8_t *drop, const uint8_t *blocks, int32_t *step_count,
                        int32_t *cum_constraints, double *usage, const double *health, const int8_t *actions, const double *u,
                        int record, int8_t *obs, double *reward, uint8_t *done, int32_t *constraints_out,
                        uint8_t *success_out)
{
    const int n = c->n_agents, W = c->width, L = c->length;
    int x[MAXA], y[MAXA], gx[MAXA], gy[MAXA], dist[MAXA];
    int px[MAXA], py[MAXA], cxs[MAXA], cys[MAXA], pre_done[MAXA], sta[MAXA], dyn[MAXA];
    double rew[MAXA];
    orc_load(c, drop, x, y, gx, gy);
    for (int i = 0; i < n; i++) dist[i] = abs(x[i] - gx[i]) + abs(y[i] - gy[i]); /* Droplet.distance :93-95 */

    *step_count += 1;                                     /* :561 */
    int success = 0;
    for (int i = 0; i < n; i++) pre_done[i] = (dist[i] == 0); /* getTaskStatus :278,365-366 */
    for (int i = 0; i < n; i++) {                          /* moveOneDroplet :325-359 */
        int ox = x[i], oy = y[i];
        double r;
        if (c->stall && dist[i] == 0) {
            r = 0.0;
        } else {
            double prob = health ? health[(size_t)x[i] * L + y[i]] : 1.0; /* getMoveProb :361-363 */
            double draw = u ? u[i] : 0.0;
            if (draw <= prob) {                            /* random.random() <= prob :335 */
                if (orc_move(&x[i], &y[i], actions[i], W, L)) return -1;
                if (blocks && orc_touching_blocks(blocks, c->n_blocks, x[i], y[i])) { x[i] = ox; y[i] = oy; } /* :338-340 */
                if (orc_any_pair_equal(x, y, n)) { x[i] = ox; y[i] = oy; } /* :341-343 */
            }
            int nd = abs(x[i] - gx[i]) + abs(y[i] - gy[i]);
            if (nd == dist[i] && dist[i] == 0) r = -0.1;
            else if (nd == dist[i] && actions[i] == 0) r = -0.25;
            else if (nd < dist[i]) r = -0.1;
            else r = -0.4;
            dist[i] = nd;
        }
        rew[i] = r; px[i] = ox; py[i] = oy; cxs[i] = x[i]; cys[i] = y[i];
    }
    /* comflic_static (:254-261): norm(cur_i-cur_j) < 2, each unordered pair counted to both */
    for (int i = 0; i < n; i++) { sta[i] = 0; dyn[i] = 0; }
    for (int i = 0; i < n - 1; i++)
        for (int j = i + 1; j < n; j++) {
            int dx = cxs[i] - cxs[j], dy = cys[i] - cys[j];
            if (dx * dx + dy * dy < 4) { sta[i]++; sta[j]++; }
        }
    /* comflic_dynamic (:263-271): norm(past_i-cur_j) < 2 over ordered pairs, counted to both */
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++)
            if (i != j) {
                int dx = px[i] - cxs[j], dy = py[i] - cys[j];
                if (dx * dx + dy * dy < 4) { dyn[i]++; dyn[j]++; }
            }
    int constraints = 0;
    for (int i = 0; i < n; i++) constraints += sta[i] + dyn[i];          /* :287 */
    for (int i = 0; i < n; i++) rew[i] = (rew[i] - (double)(2 * sta[i])) - (double)(2 * dyn[i]); /* :288 */
    if (c->stall)
        for (int i = 0; i < n; i++) if (pre_done[i]) rew[i] = 0.0;      /* :289-292 */
    int all_done = 1;
    for (int i = 0; i < n; i++) if (dist[i] != 0) all_done = 0;
    if (all_done) {                                                        /* :293-296 */
        for (int i = 0; i < n; i++) rew[i] = rew[i] + 10.0;
        if (constraints == 0) for (int i = 0; i < n; i++) rew[i] = rew[i] + 10.0;
    }
    if (record && usage)                                                   /* addUsage :459-463 */
        for (int i = 0; i < n; i++) if (dist[i] != 0) usage[(size_t)x[i] * L + y[i]] += 1.0;
    *cum_constraints += constraints;                                       /* :572 */
    for (int i = 0; i < n; i++) { drop[4 * i + 0] = (uint8_t)x[i]; drop[4 * i + 1] = (uint8_t)y[i]; }
    if (obs) orc_all_obs(c, drop, blocks, obs);                            /* :576 */
    if (*step_count < 2 * (W + L)) {                                       /* :577-585, max_step :508 */
        if (all_done && *cum_constraints == 0) success = 1;
        for (int i = 0; i < n; i++) done[i] = (uint8_t)(dist[i] == 0);
    } else {
        for (int i = 0; i < n; i++) done[i] = 1;
    }
    for (int i = 0; i < n; i++) reward[i] = rew[i];
    *constraints_out = constraints;
    *success_out = (uint8_t)success;
    return 0;
}

/* RoutingTaskManager.updateHealth (dmfb.py:465-471) */
static void orc_update_health(const orc_dmfb_cfg *c, double *usage, double *health, const double *degrade)
{
    const int cells = c->width * c->length;
    for (int k = 0; k < cells; k++)
        if (usage[k] > 50.0) {
            if (health) health[k] = health[k] * (degrade ? degrade[k] : 1.0);
            usage[k] = 0.0;
        }
}

/* ------------------------------------------------------------ batch API -- */

int orc_dmfb_step(const orc_dmfb_cfg *c, int n_envs, uint8_t *drop, const uint8_t *blocks, int32_t *step_count, int32_t *cum_constraints,
                  double *usage, const double *health, const int8_t *actions, const double *u, int record,
                  int8_t *obs, double *reward, uint8_t *done, int32_t *constraints, uint8_t *success)
{
    const int A = c->n_agents, cells = c->width * c->length, D = orc_obs_dim(c);
    int err = 0;
    for (int e = 0; e < n_envs; e++) {
        int rc = orc_step_one(c, drop + (size_t)e * A * 4, blocks ? blocks + (size_t)e * c->n_blocks * 2 : NULL,
                              step_count + e, cum_constraints + e,
                              usage ? usage + (size_t)e * cells : NULL, health ? health + (size_t)e * cells : NULL,
                              actions + (size_t)e * A, u ? u + (size_t)e * A : NULL, record,
                              obs ? obs + (size_t)e * A * D : NULL, reward + (size_t)e * A, done + (size_t)e * A,
                              constraints + e, success + e);
        if (rc) err |= 1;
    }
    return err ? -1 : 0;
}

/* DMFBenv.reset(new) (dmfb.py:589-597) with the task (layout) injected:
 * refresh (:174-183): new task; new -> health=1, usage=0, degrade redrawn (injected);
 * else updateHealth.  Then getObs. */
int orc_dmfb_reset(const orc_dmfb_cfg *c, int n_envs, const uint8_t *mask, int new_task, const uint8_t *layouts,
                   const uint8_t *block_layouts, const double *degrade_in, uint8_t *drop, uint8_t *blocks,
                   int32_t *step_count, int32_t *cum_constraints, double *usage, double *health, double *degrade,
                   int8_t *obs)
{
    const int A = c->n_agents, cells = c->width * c->length, D = orc_obs_dim(c);
    for (int e = 0; e < n_envs; e++) {
        if (mask && !mask[e]) continue;
        step_count[e] = 0;
        cum_constraints[e] = 0;
        memcpy(drop + (size_t)e * A * 4, layouts + (size_t)e * A * 4, (size_t)A * 4);
        if (blocks && block_layouts)
            memcpy(blocks + (size_t)e * c->n_blocks * 2, block_layouts + (size_t)e * c->n_blocks * 2, (size_t)c->n_blocks * 2);
        if (new_task) {
            for (int k = 0; k < cells; k++) {
                if (health) health[(size_t)e * cells + k] = 1.0;
                if (usage) usage[(size_t)e * cells + k] = 0.0;
                if (degrade) degrade[(size_t)e * cells + k] = degrade_in ? degrade_in[(size_t)e * cells + k] : 1.0;
            }
        } else if (usage) {
            orc_update_health(c, usage + (size_t)e * cells, health ? health + (size_t)e * cells : NULL,
                              degrade ? degrade + (size_t)e * cells : NULL);
        }
        if (obs) orc_all_obs(c, drop + (size_t)e * A * 4, blocks ? blocks + (size_t)e * c->n_blocks * 2 : NULL,
                             obs + (size_t)e * A * D);
    }
    return 0;
}

int orc_dmfb_observe(const orc_dmfb_cfg *c, int n_envs, const uint8_t *drop, const uint8_t *blocks, int8_t *obs)
{
    const int A = c->n_agents, D = orc_obs_dim(c);
    for (int e = 0; e < n_envs; e++)
        orc_all_obs(c, drop + (size_t)e * A * 4, blocks ? blocks + (size_t)e * c->n_blocks * 2 : NULL, obs + (size_t)e * A * D);
    return 0;
}

/* RoutingTaskManager.getglobalobs (dmfb.py:368-392) as int8 [3,W,L] per chip */
int orc_dmfb_global_state(const orc_dmfb_cfg *c, int n_envs, const uint8_t *drop, const uint8_t *blocks, int8_t *out)
{
    const int A = c->n_agents, W = c->width, L = c->length;
    for (int e = 0; e < n_envs; e++) {
        int8_t *g = out + (size_t)e * 3 * W * L;
        memset(g, 0, (size_t)3 * W * L);
        const uint8_t *d = drop + (size_t)e * A * 4;
        for (int b = 0; blocks && b < c->n_blocks; b++) {      /* add_blocks_In_gloabal_Obs (:376-381) */
            const uint8_t *bl = blocks + ((size_t)e * c->n_blocks + b) * 2;
            for (int i = bl[0]; i <= bl[0] + 1; i++)
                for (int j = bl[1]; j <= bl[1] + 1; j++) g[2 * W * L + i * L + j] = 1;
        }
        for (int i = 0; i < A; i++) {
            g[0 * W * L + d[4 * i + 0] * L + d[4 * i + 1]] = (int8_t)(i + 1);
            g[1 * W * L + d[4 * i + 2] * L + d[4 * i + 3]] = (int8_t)(i + 1);
        }
    }
    return 0;
}

/* ---------------------------------------------------- task generator -- */
/* splitmix64: only the *distribution* of generated tasks has to match the reference
 * (its own generator is numpy's global Mersenne Twister, dmfb.py:209-210). */
static inline uint64_t orc_splitmix(uint64_t *s)
{
    uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

/* _Generate_Start_End (dmfb.py:207-226): 2A uniform cells; the whole set is redrawn until the
 * minimum pairwise squared distance is > 2.  layout = [A][4] (x,y,gx,gy). */
void orc_dmfb_gen_layout(const orc_dmfb_cfg *c, uint64_t *rng, uint8_t *layout)
{
    const int A = c->n_agents, m = 2 * A;
    int px[2 * MAXA], py[2 * MAXA];
    for (;;) {
        for (int k = 0; k < m; k++) py[k] = (int)(orc_splitmix(rng) % (uint64_t)c->length);
        for (int k = 0; k < m; k++) px[k] = (int)(orc_splitmix(rng) % (uint64_t)c->width);
        int ok = 1;
        for (int i = 0; i < m && ok; i++)
            for (int j = i + 1; j < m; j++) {
                int dx = px[i] - px[j], dy = py[i] - py[j];
                if (dx * dx + dy * dy <= 2) { ok = 0; break; }
            }
        if (ok) break;
    }
    for (int i = 0; i < A; i++) {
        layout[4 * i + 0] = (uint8_t)px[i]; layout[4 * i + 1] = (uint8_t)py[i];
        layout[4 * i + 2] = (uint8_t)px[A + i]; layout[4 * i + 3] = (uint8_t)py[A + i];
    }
}

/* GenRandomBlocks (dmfb.py:228-251): each 2x2 block uniform with x_min in [0,W-4], y_min in [0,L-4], redrawn while
 * it contains a start/goal cell or overlaps an earlier block. */
void orc_dmfb_gen_blocks(const orc_dmfb_cfg *c, uint64_t *rng, const uint8_t *layout, uint8_t *blocks)
{
    const int A = c->n_agents;
    if (c->width < 5 || c->length < 5 || c->n_blocks * 4.0 / (c->width * c->length) > 0.2) return;
    for (int b = 0; b < c->n_blocks; b++) {
        for (;;) {
            int y = (int)(orc_splitmix(rng) % (uint64_t)(c->length - 3));
            int x = (int)(orc_splitmix(rng) % (uint64_t)(c->width - 3));
            int bad = 0;
            for (int i = 0; i < 2 * A && !bad; i++) {
                int px = layout[4 * (i % A) + (i < A ? 0 : 2)], py = layout[4 * (i % A) + (i < A ? 1 : 3)];
                if (px >= x && px <= x + 1 && py >= y && py <= y + 1) bad = 1;
            }
            for (int k = 0; k < b && !bad; k++) {
                int ox = blocks[2 * k], oy = blocks[2 * k + 1];
                if (!(x > ox + 1 || ox > x + 1) && !(y > oy + 1 || oy > y + 1)) bad = 1;
            }
            if (!bad) { blocks[2 * b] = (uint8_t)x; blocks[2 * b + 1] = (uint8_t)y; break; }
        }
    }
}

/* CPU-baseline driver (bench.py cpu_baseline / --impl reference): n_envs chips, `steps` lock-step
 * env steps with uniform random actions, auto-reset when an episode ends (all done or step limit),
 * degradation draws uniform.  Every step writes the full observation tensor, like the GPU path.
 * Chips are split over n_threads POSIX threads (the reference itself is single-threaded Python;
 * independent chips are its only parallelism).  Returns the number of agent-steps executed;
 * *checksum defeats dead-code elimination. */
#include <pthread.h>

typedef struct {
    const orc_dmfb_cfg *c;
    int e0, e1, steps;
    uint64_t seed, total;
    int8_t *obs;
} orc_roll_job;

static void *orc_dmfb_roll_thread(void *arg)
{
    orc_roll_job *job = (orc_roll_job *)arg;
    const orc_dmfb_cfg *c = job->c;
    const int A = c->n_agents, cells = c->width * c->length, D = orc_obs_dim(c);
    double *usage = (double *)calloc((size_t)cells, sizeof(double));
    double *health = (double *)malloc((size_t)cells * sizeof(double));
    double *degrade = (double *)malloc((size_t)cells * sizeof(double));
    uint64_t total = 0;
    for (int e = job->e0; e < job->e1; e++) {
        uint64_t rng = job->seed * 0x100000001B3ull + (uint64_t)e;
        uint8_t drop[4 * MAXA];
        int8_t acts[MAXA];
        double u[MAXA], rew[MAXA];
        uint8_t done[MAXA], succ;
        int32_t sc = 0, cc = 0, cons;
        int8_t *obs = job->obs + (size_t)e * A * D;
        for (int k = 0; k < cells; k++) {
            usage[k] = 0.0; health[k] = 1.0;
            degrade[k] = c->b_degrade ? (double)(orc_splitmix(&rng) >> 11) * (1.0 / 9007199254740992.0) * 0.4 + 0.6 : 1.0;
        }
        orc_dmfb_gen_layout(c, &rng, drop);
        for (int t = 0; t < job->steps; t++) {
            for (int i = 0; i < A; i++) {
                acts[i] = (int8_t)(orc_splitmix(&rng) % 5u);
                u[i] = (double)(orc_splitmix(&rng) >> 11) * (1.0 / 9007199254740992.0);
            }
            orc_step_one(c, drop, NULL, &sc, &cc, usage, c->b_degrade ? health : NULL, acts, u, 1, obs, rew, done,
                         &cons, &succ);
            int all = 1;
            for (int i = 0; i < A; i++) { all &= done[i]; total += (uint64_t)(rew[i] < 0.0); }
            if (all) {
                sc = 0; cc = 0;
                orc_dmfb_gen_layout(c, &rng, drop);
                orc_update_health(c, usage, health, degrade);
                orc_all_obs(c, drop, NULL, obs);
            }
        }
    }
    free(usage); free(health); free(degrade);
    job->total = total;
    return NULL;
}

int64_t orc_dmfb_rollout(const orc_dmfb_cfg *c, int n_envs, int steps, uint64_t seed, int8_t *obs /* [n_envs,A,D] */,
                         int n_threads, uint64_t *checksum)
{
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 1024) n_threads = 1024;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
    orc_roll_job *jobs = (orc_roll_job *)malloc(sizeof(orc_roll_job) * (size_t)n_threads);
    for (int t = 0; t < n_threads; t++) {
        jobs[t].c = c; jobs[t].steps = steps; jobs[t].seed = seed; jobs[t].obs = obs; jobs[t].total = 0;
        jobs[t].e0 = (int)((int64_t)n_envs * t / n_threads);
        jobs[t].e1 = (int)((int64_t)n_envs * (t + 1) / n_threads);
        pthread_create(&th[t], NULL, orc_dmfb_roll_thread, &jobs[t]);
    }
    uint64_t total = 0;
    for (int t = 0; t < n_threads; t++) { pthread_join(th[t], NULL); total += jobs[t].total; }
    free(th); free(jobs);
    if (checksum) *checksum = total;
    return (int64_t)n_envs * steps * c->n_agents;
}
