/*
 * dmfb_b200.h — C ABI of the B200-native batched DMFB / MEDA environment step.
 *
 * This is the drop-in boundary for ONE hot path of jesselasse/MARL-DMFB: the
 * environment step of env/DMFB/dmfb.py and env/MEDA/meda.py (apply moves ->
 * fluidic-constraint / collision resolution -> usage / degradation -> rewards,
 * dones, avail mask -> per-agent fov x fov observation + global state),
 * batched over N independent chips ("envs") whose state lives in HBM as
 * struct-of-arrays.  The reference has no FFI layer of its own (it is pure
 * Python), so each entry point below cites the reference *method* it replaces
 * (file:line relative to the reference repo root).
 *
 * Conventions
 *  - Plain C: pointers and sizes only, no torch / C++ types.
 *  - Every buffer is allocated and owned by the caller (device memory unless
 *    the function name says _host); the library keeps no state about env
 *    batches (what it does keep per process: a launch counter, a thread-local
 *    error string, the one-time cudaFuncSetAttribute bookkeeping of its kernels
 *    and the tuning knobs DMFB_TILE_ENVS / DMFB_NO_PDL / MEDA_WARP_ENVS /
 *    MEDA_WARPS_PER_CTA read once from the environment), never allocates behind
 *    the caller's back (except the explicit dmfb_host_* / meda_host_* handle
 *    API) and never frees caller memory.
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *    Calls on one env batch are ordered by that stream; the library is
 *    re-entrant across batches / devices.  dmfb_step is launched with
 *    programmatic stream serialization: back-to-back steps overlap the
 *    prologue of step t+1 with the tail of step t, every global access of
 *    step t+1 still waits for step t to complete.
 *  - Return value: DMFB_OK (0) or a DMFB_ERR_* code.  The Python host layer
 *    maps the codes onto the exception types the reference raises.
 *  - Coordinates: DMFB x in [0,width), y in [0,length) (dmfb.py:103-124).
 *    MEDA x in [0,length), y in [0,width) (meda.py:131-138); grids are
 *    indexed [y][x] there (meda.py:302-309).
 */
#ifndef DMFB_B200_H
#define DMFB_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DMFB_ABI_VERSION 4
#define DMFB_MAX_DIM 128    /* max chip width / length (cells) */
#define DMFB_MAX_AGENTS 32  /* max droplets per chip */
#define DMFB_MAX_FOV 19     /* max field of view (cells) */
#define DMFB_MAX_BLOCKS 32  /* max 2x2 obstacles per chip */
#define DMFB_L2_WORDS 12    /* ceil(19*19/32) bit-mask words per boundary pattern */

enum dmfb_status {
    DMFB_OK = 0,
    DMFB_ERR_FOV_TOO_LARGE = 1,      /* RuntimeError('Fov is too large')            dmfb.py:139-140 */
    DMFB_ERR_TOO_MANY_DROPLETS = 2,  /* TypeError / RuntimeError('Too many droplets') dmfb.py:144-146, meda.py:151-154 */
    DMFB_ERR_BAD_ARG = 3,            /* null pointer, bad size, unsupported dimension */
    DMFB_ERR_DIV_ZERO = 4,           /* ZeroDivisionError: fov//2 == 10               dmfb.py:446 */
    DMFB_ERR_CUDA = 5,               /* launch / runtime error (see dmfb_last_cuda_error) */
    DMFB_ERR_CHIP_TOO_SMALL = 6      /* AssertionError: width >= 5 and length >= 5     dmfb.py:489 */
};

/* sticky status bits (dmfb_out_t.status / meda_out_t.status, dmfb_state_t.gen_status / meda_state_t.gen_status) */
#define DMFB_STATUS_ILLEGAL_ACTION 1  /* an action outside the action set was applied: TypeError (dmfb.py:115-116) */
#define DMFB_STATUS_SAMPLER_GAVE_UP 2 /* a task / obstacle generator found no legal draw within its attempt budget (the
                                         reference would loop for ever, dmfb.py:212-224,246-250, meda.py:213-233): the env
                                         keeps its previous layout; the host layer raises RuntimeError */

/* step flags */
#define DMFB_STEP_RECORD_USAGE 1u  /* DMFBenv.step(record=True): addUsage, dmfb.py:570-571 */
#define DMFB_STEP_FREEZE_TERM 2u   /* lock-step rollouts: envs whose `terminated` flag was already set
                                      are not stepped and emit the zero padding of rollout.py:131-141 */
#define DMFB_STEP_AUTO_RESET 4u    /* vectorised rollouts: envs that terminate in this step (all done or step
                                      limit) get DMFBenv.reset(new=False) right after it; their obs rows then hold
                                      the first observation of the new episode, reward/done/info those of the
                                      finished step.  Bit 31 of the flags is reserved for the library. */
#define DMFB_STEP_SKIP_TASK_SEARCH 8u /* with DMFB_STEP_AUTO_RESET on dense 10-droplet chips (dmfb_state_t.next_task): this
                                      call does not launch the run-ahead task search kernel behind the step.  A caller
                                      that launches it every P-th step only passes DMFB_STEP_SEARCH_SHARE(P) on that
                                      step - P times the attempts per open search - and this flag on the others; the
                                      tasks drawn do not depend on when the search runs.  Ignored elsewhere. */
#define DMFB_STEP_SEARCH_SHARE(p) (((uint32_t)(p) & 0xFFu) << 8)

/* ------------------------------------------------------------------ DMFB -- */

/* Static description of a batch of identical chips; fill with dmfb_cfg_init().
 * Passed by value to the kernels (it carries the small lookup tables). */
typedef struct dmfb_cfg {
    int32_t width, length;  /* chip cells: x in [0,width), y in [0,length) */
    int32_t n_agents;       /* droplets per chip (A) */
    int32_t n_blocks;       /* 2x2 obstacles per chip (0 in every shipped config).  Like GenRandomBlocks (dmfb.py:232-234)
                               dmfb_cfg_init stores 0 here when 4*n_blocks/(W*L) > 0.2. */
    int32_t fov;            /* side of the square partial observation */
    int32_t stall;          /* reference ctor arg `stall` (dmfb.py:331) */
    int32_t b_degrade;      /* electrode degradation on/off */
    int32_t max_step;       /* 2*(width+length), dmfb.py:508 */
    int32_t n_actions;      /* 5, dmfb.py:26-31 */
    int32_t obs_dim;        /* 3*fov*fov+2, dmfb.py:633-640 (4*fov*fov+2 for DMFB_OBS_V01) */
    int32_t l2_words;       /* ceil(fov*fov/32) */
    int32_t obs_version;    /* DMFB_OBS_BASE (dmfb_cfg_init) or DMFB_OBS_V01 (dmfb_cfg_set_obs_version) */
    double per_degrade;     /* fraction of degrading electrodes, dmfb.py:157-164 */
    int64_t env_base;       /* global index of env 0 of this batch (multi-GPU sharding; RNG stream id) */
    /* dirct table: dir_x[d + width-1] for d = goal_x - x (dmfb.py:442-454), same for y */
    int8_t dir_x[2 * DMFB_MAX_DIM];
    int8_t dir_y[2 * DMFB_MAX_DIM];
    /* boundary-layer bit patterns (dmfb.py:428-439): code c in [0, 2*(fov/2)]:
     * 0 none, 1..hf = first c rows (cols), hf+1..2hf = last c-hf rows (cols).
     * Bit q = x*fov + y of l2_row[c] / l2_col[c] is set when row x / col y is off-chip. */
    uint32_t l2_row[DMFB_MAX_FOV][DMFB_L2_WORDS];
    uint32_t l2_col[DMFB_MAX_FOV][DMFB_L2_WORDS];
} dmfb_cfg_t;

/* Per-env state, struct-of-arrays over N envs, device pointers. */
typedef struct dmfb_state {
    int32_t n_envs;
    int32_t usage_log_cap;  /* entries per env in usage_log (0 = no log) */
    uint8_t* drop;          /* [N,A,4] = x, y, goal_x, goal_y              (Droplet, dmfb.py:74-79) */
    uint8_t* start;         /* [N,A,2] start cells for restart(), may be NULL (dmfb.py:185-190) */
    int32_t* step_count;    /* [N]                                          (dmfb.py:510,561) */
    int32_t* constraints;   /* [N] episode-cumulative constraint count      (dmfb.py:511,572) */
    uint8_t* terminated;    /* [N] 1 once all(dones) was returned by a step (rollout.py:34-35) */
    uint32_t* episode;      /* [N] episode counter (RNG stream selector) */
    uint32_t* usage;        /* [N,W,L] actuation counts m_usage (updated with fire-and-forget RED.ADD), may be NULL
                               when !b_degrade (dmfb.py:148,459-463) */
    double* health;         /* [N,W,L] m_health, NULL == all 1.0            (dmfb.py:147,361-363) */
    double* degrade;        /* [N,W,L] m_degrade, NULL == all 1.0           (dmfb.py:151,157-166) */
    uint8_t* blocks;        /* [N,n_blocks,2] (x_min,y_min) of the 2x2 blocks (Block, dmfb.py:34-41,246); NULL when n_blocks==0 */
    /* Optional (both NULL = off) log of the actuated cells.  Nothing reads m_usage between two resets (updateHealth runs
     * at reset, dmfb.py:465-471), so a step appends the A cells it would increment (x | y<<8, 0xFFFF = none) - 2A
     * coalesced bytes instead of A scattered read-modify-writes in an array that does not fit L2 - and the reset that
     * needs the counters replays the env's log into `usage` first.  usage + log together are always the reference's
     * m_usage; dmfb_flush_usage() folds the log in for a reader.  A full log falls back to direct increments. */
    uint16_t* usage_log;    /* [N, usage_log_cap, A] */
    int32_t* usage_log_len; /* [N] entries in use */
    /* Optional (both NULL = off) task prefetch for DMFB_STEP_AUTO_RESET.  _Generate_Start_End redraws the whole point set
     * until it is legal (dmfb.py:212-224), a geometric number of attempts whose tail would keep the whole launch waiting
     * for the unluckiest of the envs that reset in a step.  Attempt k of (seed, env, episode) is a pure function, so the
     * search for the NEXT episode's task can run ahead: every step examines a bounded number of attempts for envs
     * whose next task is not known yet (at the end of the step kernel, or, for dense 10-droplet chips, in a second
     * kernel that dmfb_step launches behind it) and parks the first accepted one here; the reset then only picks it up
     * (or finishes the search where it stopped).  The task drawn is the same first accepted attempt either way. */
    uint32_t* next_task;    /* [N,A] packed x | y<<8 | goal_x<<16 | goal_y<<24 of the next episode's task */
    uint32_t* next_cursor;  /* [N] zero-initialised: attempts already examined; bit 31 = next_task is valid */
    int32_t* gen_status;    /* [1] sticky DMFB_STATUS_* bits raised by the generators, may be NULL */
    /* Optional (NULL = off) bit map of the degraded cells: bit k (k = x*length + y) of an env's ceil(W*L/32) words is set
     * iff health[k] != 1.0.  getMoveProb (dmfb.py:361-363) reads one float64 per droplet and step at a random place of an
     * array that cannot live in L2 (1.3 GB for 64K 50x50 chips); the bit map can (20 MB), and a clear bit answers
     * "1.0" without touching the array.  dmfb_reset keeps it in step with health (new_task clears it, updateHealth sets
     * bits); a caller that writes `health` itself calls dmfb_sync_health_bits() afterwards. */
    uint32_t* health_bits;  /* [N, ceil(W*L/32)] */
} dmfb_state_t;

/* Per-step outputs, device pointers; any pointer except `obs` may be NULL. */
typedef struct dmfb_out {
    int8_t* obs;            /* [N,A,obs_dim]   getObs(), dmfb.py:622-626 (16-byte aligned for the fast path) */
    float* reward;          /* [N,A]           rewards dict values, dmfb.py:573-574 */
    double* reward_f64;     /* [N,A]           same, float64 bit-exact with the reference (optional) */
    float* team_reward;     /* [N]             sum(r)/len(r), rollout.py:33 */
    uint8_t* done;          /* [N,A]           dones dict values, dmfb.py:577-585 */
    uint8_t* avail;         /* [N,A,n_actions] all ones; zeros on padded steps (rollout.py:22,138-139) */
    int32_t* constraints;   /* [N]             info['constraints'] of THIS step, dmfb.py:586 */
    uint8_t* success;       /* [N]             info['success'], dmfb.py:579-580 */
    uint8_t* terminated;    /* [N]             all(dones), rollout.py:34-35 */
    uint8_t* padded;        /* [N]             1 when the env was frozen (DMFB_STEP_FREEZE_TERM) */
    int32_t* status;        /* [1]             sticky DMFB_STATUS_ILLEGAL_ACTION (TypeError, dmfb.py:115-116) */
} dmfb_out_t;

/* Validate arguments like DMFBenv.__init__/RoutingTaskManager.__init__ (dmfb.py:128-155,
 * 487-508) and fill the derived fields and tables. */
int dmfb_cfg_init(dmfb_cfg_t* cfg, int width, int length, int n_agents, int n_blocks, int fov,
                  int stall, int b_degrade, double per_degrade);

/* Observation variant.  DMFB_OBS_BASE: DMFBenv.getOneObs -> RoutingTaskManager.getOneObs (dmfb.py:395-457,614-620).
 * DMFB_OBS_V01: DMFBenv_v0_1.getOneObs (dmfb.py:723-835, selected by `--version 0.1`, common/config.py:6-8):
 * (4,fov,fov) = droplets / own goal / goals of the visible others drawn where the ray to the goal leaves the
 * window / obstacles + border, then 2 direction entries.  The reference returns float64 with the direction
 * ((tar_y-y)/length, (tar_x-x)/width); the layers are integral and are emitted as int8, the direction entries as
 * their integer numerators (tar_y-y, tar_x-x).  Sets obs_version and obs_dim; returns DMFB_ERR_BAD_ARG otherwise. */
#define DMFB_OBS_BASE 0
#define DMFB_OBS_V01 1
int dmfb_cfg_set_obs_version(dmfb_cfg_t* cfg, int obs_version);

/* DMFBenv.step (dmfb.py:560-587) for every env of the batch.
 *  actions   [N,A] device, element size `action_elem_size` in {1,4,8} bytes (int8/int32/int64), values 0..4
 *  u_inject  [N,A] float64 move-success draws that replace random.random() (dmfb.py:335), or NULL:
 *            then draws come from Philox4x32-10 keyed by (seed, env, episode, step, agent);
 *            ignored when state->health is NULL (probability 1). */
int dmfb_step(const dmfb_cfg_t* cfg, const dmfb_state_t* state, const void* actions,
              int action_elem_size, const double* u_inject, uint64_t seed, uint32_t flags,
              const dmfb_out_t* out, void* stream);

/* DMFBenv.reset(new) (dmfb.py:589-597 -> refresh :174-183 -> Generate_task :168 / updateHealth :465)
 *  mask      [N] uint8, reset only envs with mask != 0; NULL = all envs
 *  new_task  1: reference `new=True` (health=1, usage=0, redraw degrade); 0: updateHealth()
 *  layouts   [N,A,4] uint8 (x,y,gx,gy) injected task, or NULL: on-device generator equivalent to
 *            _Generate_Start_End (dmfb.py:207-226; uniform cells, whole set rejected until every
 *            pairwise squared distance among the 2A points is > 2)
 *  block_layouts [N,n_blocks,2] uint8 (x_min,y_min) injected obstacles, or NULL: on-device generator equivalent to
 *            GenRandomBlocks (dmfb.py:228-251; redrawn while a block covers a start/goal cell or overlaps a block)
 *  degrade   [N,W,L] float64 injected degradation factors (only read when new_task && b_degrade), or NULL:
 *            on-device equivalent of _random_health_statue (dmfb.py:157-164)
 *  obs       [N,A,obs_dim] receives getObs() of the reset envs (rows of other envs untouched), may be NULL */
int dmfb_reset(const dmfb_cfg_t* cfg, const dmfb_state_t* state, const uint8_t* mask, int new_task,
               const uint8_t* layouts, const uint8_t* block_layouts, const double* degrade, uint64_t seed,
               int8_t* obs, void* stream);

/* Folds state->usage_log into state->usage for every env and empties the log (no-op without a log). */
int dmfb_flush_usage(const dmfb_cfg_t* cfg, const dmfb_state_t* state, void* stream);

/* Rebuilds state->health_bits from state->health (no-op when either is NULL). */
int dmfb_sync_health_bits(const dmfb_cfg_t* cfg, const dmfb_state_t* state, void* stream);

/* DMFBenv.getObs() of the current state for all envs (dmfb.py:614-626). */
int dmfb_observe(const dmfb_cfg_t* cfg, const dmfb_state_t* state, int8_t* obs, void* stream);

/* RoutingTaskManager.getglobalobs() (dmfb.py:368-392) as int8 [N,3,W,L] — the `get_state` analogue. */
int dmfb_global_state(const dmfb_cfg_t* cfg, const dmfb_state_t* state, int8_t* out, void* stream);

/* DMFBenv.restart() (dmfb.py:599-605): droplets back to their start cells, counters zeroed,
 * health untouched.  Needs state->start. */
int dmfb_restart(const dmfb_cfg_t* cfg, const dmfb_state_t* state, const uint8_t* mask, int8_t* obs,
                 void* stream);

/* ------------------------------------------------------------------ MEDA -- */

#define MEDA_OBS_BASE 0 /* MEDAEnv.getOneObs      (4,fov,fov)+2, meda.py:613-674 (emitted as int8; all values integral) */
#define MEDA_OBS_V01 1  /* MEDAEnv_v0_1.getOneObs (4,fov,fov)+2, meda.py:788-844; the two direction entries are emitted as the
                         * integer numerators (dy, dx) of the reference's (dy / width, dx / length) */
#define MEDA_OBS_V02 2  /* MEDAEnv_v0_2.getOneObs (3,fov,fov)+2 int8, meda.py:850-897 */

typedef struct meda_cfg {
    int32_t width, length;  /* grid is m_health[width][length] indexed [y][x] */
    int32_t n_agents;
    int32_t fov;
    int32_t b_degrade;
    int32_t obs_version;    /* MEDA_OBS_BASE, MEDA_OBS_V01 or MEDA_OBS_V02 */
    int32_t max_step;       /* width+length, meda.py:492 */
    int32_t n_actions;      /* 9, meda.py:23-32 */
    int32_t obs_dim;        /* 4*fov^2+2 (base) or 3*fov^2+2 (v0_2) */
    int32_t radius;         /* 2, meda.py:150 */
    double per_degrade;
    int64_t env_base;
    /* v0_2 direction vector: dir_y[d+width-1] = round(d/(width/30)), dir_x[d+length-1] = round(d/(length/30)) (meda.py:895) */
    int8_t dir_x[2 * DMFB_MAX_DIM];
    int8_t dir_y[2 * DMFB_MAX_DIM];
    /* v0_2 writes the others' goals in CPython set iteration order of the observed indices (meda.py:871-878);
     * for A > 8 that is not ascending, see meda_set_order / the `set_order` argument of meda_step. */
} meda_cfg_t;

typedef struct meda_state {
    int32_t n_envs;
    int32_t usage_log_cap;  /* entries per env in usage_log (0 = no log) */
    uint8_t* drop;          /* [N,A,4] = x_center, y_center, goal x_center, goal y_center (meda.py:35-47) */
    uint8_t* start;         /* [N,A,2] may be NULL */
    uint8_t* status;        /* [N,A] sticky arrival flags (meda.py:159,277) */
    int32_t* step_count;    /* [N] */
    int32_t* fails;         /* [N] episode-cumulative punish count; reference `fails` == -0.6*count (meda.py:521) */
    uint8_t* terminated;    /* [N] */
    uint32_t* episode;      /* [N] */
    uint32_t* usage;        /* [N,W,L] */
    double* health;         /* [N,W,L], NULL == all 1.0 */
    double* degrade;        /* [N,W,L], NULL == all 1.0 */
    /* Optional log of the actuated droplets, same idea as dmfb_state_t.usage_log: a step appends the centre of every
     * droplet whose 5x5 footprint it would increment (x | y<<8, 0xFFFF = none); meda_reset / meda_flush_usage replay it. */
    uint16_t* usage_log;    /* [N, usage_log_cap, A] */
    int32_t* usage_log_len; /* [N] */
    /* Optional scratch for DMFB_STEP_AUTO_RESET (both NULL = a masked meda_reset over the whole batch after the step):
     * the step appends the envs that just terminated to reset_list, and a small second kernel resets exactly those. */
    int32_t* reset_list;    /* [N] */
    int32_t* reset_count;   /* [2], zero-initialised: entries in reset_list, finished-CTA ticket */
    int32_t* gen_status;    /* [1] sticky DMFB_STATUS_* bits raised by the task generator, may be NULL */
    /* Optional (NULL = off) bit map of the degraded cells, bit y*length + x set iff health[y][x] != 1.0 (same idea as
     * dmfb_state_t.health_bits): getMoveProb (meda.py:302-309) averages 25 float64 cells per droplet and step; 25 clear
     * bits answer "1.0" (25 ones sum to 25.0 exactly) without the gather.  meda_sync_health_bits() after writing health. */
    uint32_t* health_bits;  /* [N, ceil(W*L/32)] */
} meda_state_t;

typedef struct meda_out {
    int8_t* obs;            /* [N,A,obs_dim] */
    float* reward;          /* [N,A] */
    double* reward_f64;     /* [N,A] optional */
    float* team_reward;     /* [N] */
    uint8_t* done;          /* [N,A] */
    uint8_t* avail;         /* [N,A,9] */
    int32_t* constraints;   /* [N] punish count of THIS step; info['constraints'] == -0.6*count (meda.py:256,538) */
    uint8_t* success;       /* [N] */
    uint8_t* terminated;    /* [N] */
    uint8_t* padded;        /* [N] */
    int32_t* status;        /* [1] sticky bit 0: action outside 0..8 seen */
} meda_out_t;

int meda_cfg_init(meda_cfg_t* cfg, int width, int length, int n_agents, int fov, int b_degrade,
                  double per_degrade, int obs_version);
/* MEDAEnv.step (meda.py:513-539). `set_order` is NULL or a device table [2^A][A] uint8 (v0_1 / v0_2, A>8). */
int meda_step(const meda_cfg_t* cfg, const meda_state_t* state, const void* actions,
              int action_elem_size, const double* u_inject, uint64_t seed, uint32_t flags,
              const uint8_t* set_order, const meda_out_t* out, void* stream);
/* MEDAEnv.reset (meda.py:541-550): refresh tasks, obs, then updateHealth. */
int meda_reset(const meda_cfg_t* cfg, const meda_state_t* state, const uint8_t* mask, int new_chip,
               const uint8_t* layouts, const double* degrade, uint64_t seed, const uint8_t* set_order,
               int8_t* obs, void* stream);
int meda_observe(const meda_cfg_t* cfg, const meda_state_t* state, const uint8_t* set_order,
                 int8_t* obs, void* stream);
/* MEDAEnv.restart (meda.py:552-561 -> RoutingTaskManager.restart :170-173): droplets back to their start squares,
 * status and step_count cleared, `fails` kept (as in the reference).  Needs state->start. */
int meda_restart(const meda_cfg_t* cfg, const meda_state_t* state, const uint8_t* mask, const uint8_t* set_order,
                 int8_t* obs, void* stream);
/* Folds state->usage_log into state->usage for every env and empties the log (no-op without a log). */
int meda_flush_usage(const meda_cfg_t* cfg, const meda_state_t* state, void* stream);
/* Rebuilds state->health_bits from state->health (no-op when either is NULL). */
int meda_sync_health_bits(const meda_cfg_t* cfg, const meda_state_t* state, void* stream);
/* Host helper: iteration order of the CPython set {i : bit i of mask_bits} built by ascending insertion
 * (MEDAEnv_v0_2 iterates such a set, meda.py:862-872).  out[0..n_max) = elements in iteration order, 0xFF padded.
 * The device table `set_order` is [2^A][A] uint8 with row m = meda_set_order(m, A, ...). */
int meda_set_order(uint32_t mask_bits, int n_max, uint8_t* out);

/* ------------------------------------------------------------- utilities -- */

int dmfb_abi_version(void);
/* Text of the last CUDA error seen by this thread inside the library ("" if none). */
const char* dmfb_last_cuda_error(void);
/* Number of kernel launches issued by the library from this process so far (bench bookkeeping). */
uint64_t dmfb_launch_count(void);

/* ---------------------------------------------- host-buffer entry points -- */
/* Reference-facing variant of the same path with HOST buffers: the handle owns the device state,
 * pinned staging and streams; every call copies its inputs host->device and its results
 * device->host before returning (this is what bench.py's `e2e` times). */
typedef struct dmfb_host_env dmfb_host_env_t;

int dmfb_host_create(const dmfb_cfg_t* cfg, int n_envs, int device, int n_chunks, dmfb_host_env_t** out);
void dmfb_host_destroy(dmfb_host_env_t* h);
/* host pointers; obs [N,A,obs_dim] */
int dmfb_host_reset(dmfb_host_env_t* h, int new_task, const uint8_t* layouts, const double* degrade,
                    uint64_t seed, int8_t* obs);
/* actions int8 [N,A] host; outputs host (any but obs may be NULL) */
int dmfb_host_step(dmfb_host_env_t* h, const int8_t* actions, const double* u_inject, uint64_t seed,
                   uint32_t flags, int8_t* obs, float* reward, uint8_t* done, int32_t* constraints,
                   uint8_t* success);
/* Packed observation transfer for dmfb_host_step.  n_threads > 0 turns it on: `dma_percent` % of the envs still
 * arrive unpacked by DMA, the rest crosses PCIe as 4-bit cells (cell values are 0..n_agents, so n_agents <= 15) and
 * is expanded into `obs` by a pool of n_threads host threads (the caller's thread included) while that DMA runs.
 * The result in `obs` is byte-identical.  n_threads <= 0 turns it off again. */
int dmfb_host_set_transfer(dmfb_host_env_t* h, int n_threads, int dma_percent);
/* The host half of that transfer on its own (no GPU involved): expands n_records packed records - ceil(cells/2) bytes
 * of two 4-bit cells each (cell 2j in the low nibble of byte j), then 2 raw bytes, `packed_stride` bytes apart - into
 * contiguous int8 records of cells+2 bytes with n_threads threads. */
int dmfb_host_unpack_records(const uint8_t* packed, size_t packed_stride, int8_t* out, int cells, size_t n_records,
                             int n_threads);
/* The same for MEDA: MEDAEnv.step / reset (meda.py:513-550) with host buffers.  obs [N,A,obs_dim] int8; `constraints`
 * is the punish COUNT of the step (info['constraints'] == -0.6 * count, meda.py:256,538).  DMFB_STEP_AUTO_RESET resets
 * the envs that terminate inside the step launch.  The handle builds the CPython-set order table itself when the
 * observation variant needs one (v0_1 / v0_2 with more than 8 droplets; at most 16 droplets then). */
typedef struct meda_host_env meda_host_env_t;
int meda_host_create(const meda_cfg_t* cfg, int n_envs, int device, meda_host_env_t** out);
void meda_host_destroy(meda_host_env_t* h);
int meda_host_reset(meda_host_env_t* h, int new_chip, const uint8_t* layouts, const double* degrade, uint64_t seed,
                    int8_t* obs);
int meda_host_step(meda_host_env_t* h, const int8_t* actions, const double* u_inject, uint64_t seed, uint32_t flags,
                   int8_t* obs, float* reward, uint8_t* done, int32_t* constraints, uint8_t* success);

/* pinned host memory helpers (cudaHostAlloc / cudaFreeHost) so that callers outside torch can
 * give the copies a DMA-able buffer */
void* dmfb_host_alloc_pinned(size_t bytes);
void dmfb_host_free_pinned(void* p);

#ifdef __cplusplus
}
#endif
#endif /* DMFB_B200_H */
