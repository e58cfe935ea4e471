// host_api_meda.cu — the MEDA hot path behind HOST buffers (meda_host_* in include/dmfb_b200.h).
//
// Reference-facing shape of MEDAEnv.step / reset (env/MEDA/meda.py:513-550) for N chips: actions come from host memory,
// observations / rewards / dones / info go back to host memory, every call.  The handle owns the device-resident state
// and one stream; inputs are copied H2D and results D2H inside the call (plain DMA: a MEDA observation row is
// 4 * 19 * 19 + 2 bytes per droplet, so this path is bound by the D2H copy exactly like dmfb_host_step).
#include <new>

#include "common.cuh"

using namespace dmfb;

struct meda_host_env {
    meda_cfg_t cfg;
    int n_envs = 0, device = 0;
    uint8_t *drop = nullptr, *start = nullptr, *status = nullptr, *terminated = nullptr;
    int32_t *step_count = nullptr, *fails = nullptr, *usage_log_len = nullptr, *gen_status = nullptr;
    uint32_t *episode = nullptr, *usage = nullptr, *health_bits = nullptr;
    uint16_t* usage_log = nullptr;
    double *health = nullptr, *degrade = nullptr;
    uint8_t* set_order = nullptr;       // [2^A][A] for the v0_1 / v0_2 observations with more than 8 droplets
    int8_t* d_actions = nullptr;
    double* d_u = nullptr;
    int8_t* d_obs = nullptr;
    float* d_reward = nullptr;
    uint8_t *d_done = nullptr, *d_succ = nullptr, *d_layouts = nullptr;
    int32_t* d_cons = nullptr;
    cudaStream_t stream = nullptr;
};

namespace {

template <typename T>
int dev_alloc(T** p, size_t count)
{
    DMFB_CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(p), count * sizeof(T)));
    DMFB_CUDA_TRY(cudaMemset(*p, 0, count * sizeof(T)));
    return DMFB_OK;
}

meda_state_t state_of(const meda_host_env* h)
{
    meda_state_t s{};
    s.n_envs = h->n_envs;
    s.drop = h->drop; s.start = h->start; s.status = h->status; s.step_count = h->step_count; s.fails = h->fails;
    s.terminated = h->terminated; s.episode = h->episode; s.usage = h->usage; s.health = h->health; s.degrade = h->degrade;
    s.usage_log = h->usage_log; s.usage_log_len = h->usage_log_len; s.usage_log_cap = h->usage_log ? h->cfg.max_step : 0;
    s.gen_status = h->gen_status; s.health_bits = h->health_bits;
    return s;
}

}  // namespace

extern "C" {

int meda_host_create(const meda_cfg_t* cfg, int n_envs, int device, meda_host_env_t** out)
{
    if (!cfg || !out || n_envs <= 0) return DMFB_ERR_BAD_ARG;
    if (cfg->obs_version != MEDA_OBS_BASE && cfg->n_agents > 16) return DMFB_ERR_BAD_ARG;
    DMFB_CUDA_TRY(cudaSetDevice(device));
    meda_host_env* h = new (std::nothrow) meda_host_env();
    if (!h) return DMFB_ERR_BAD_ARG;
    h->cfg = *cfg; h->n_envs = n_envs; h->device = device;
    const size_t N = (size_t)n_envs, A = (size_t)cfg->n_agents, cells = (size_t)cfg->width * cfg->length;
    int rc = DMFB_OK;
#define TRY_ALLOC(p, n) if ((rc = dev_alloc(&h->p, (n))) != DMFB_OK) { meda_host_destroy(h); return rc; }
    TRY_ALLOC(drop, N * A * 4) TRY_ALLOC(start, N * A * 2) TRY_ALLOC(status, N * A) TRY_ALLOC(terminated, N)
    TRY_ALLOC(step_count, N) TRY_ALLOC(fails, N) TRY_ALLOC(episode, N) TRY_ALLOC(gen_status, 1)
    if (cfg->b_degrade) {
        TRY_ALLOC(usage, N * cells) TRY_ALLOC(health, N * cells) TRY_ALLOC(degrade, N * cells)
        TRY_ALLOC(usage_log, N * (size_t)cfg->max_step * A) TRY_ALLOC(usage_log_len, N)
        TRY_ALLOC(health_bits, N * ((cells + 31) / 32))
    }
    TRY_ALLOC(d_actions, N * A) TRY_ALLOC(d_u, N * A) TRY_ALLOC(d_obs, N * A * (size_t)cfg->obs_dim)
    TRY_ALLOC(d_reward, N * A) TRY_ALLOC(d_done, N * A) TRY_ALLOC(d_cons, N) TRY_ALLOC(d_succ, N)
    TRY_ALLOC(d_layouts, N * A * 4)
    if (cfg->obs_version != MEDA_OBS_BASE && cfg->n_agents > 8) {     // CPython set order table (meda.py:862-878)
        const size_t rows = (size_t)1 << cfg->n_agents;
        TRY_ALLOC(set_order, rows * A)
        uint8_t* tab = new (std::nothrow) uint8_t[rows * A];
        if (!tab) { meda_host_destroy(h); return DMFB_ERR_BAD_ARG; }
        for (size_t m = 0; m < rows; ++m) meda_set_order((uint32_t)m, cfg->n_agents, tab + m * A);
        cudaError_t e = cudaMemcpy(h->set_order, tab, rows * A, cudaMemcpyHostToDevice);
        delete[] tab;
        if (e != cudaSuccess) { meda_host_destroy(h); return cuda_fail(e, "set_order upload"); }
    }
#undef TRY_ALLOC
    cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { meda_host_destroy(h); return cuda_fail(e, "cudaStreamCreate"); }
    *out = h;
    return DMFB_OK;
}

void meda_host_destroy(meda_host_env_t* h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamDestroy(h->stream);
    void* ptrs[] = {h->drop, h->start, h->status, h->terminated, h->step_count, h->fails, h->usage_log_len, h->gen_status,
                    h->episode, h->usage, h->health_bits, h->usage_log, h->health, h->degrade, h->set_order, h->d_actions,
                    h->d_u, h->d_obs, h->d_reward, h->d_done, h->d_succ, h->d_layouts, h->d_cons};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    delete h;
}

int meda_host_reset(meda_host_env_t* h, int new_chip, const uint8_t* layouts, const double* degrade, uint64_t seed, int8_t* obs)
{
    if (!h) return DMFB_ERR_BAD_ARG;
    DMFB_CUDA_TRY(cudaSetDevice(h->device));
    const size_t N = (size_t)h->n_envs, A = (size_t)h->cfg.n_agents, D = (size_t)h->cfg.obs_dim;
    const size_t cells = (size_t)h->cfg.width * h->cfg.length;
    cudaStream_t s = h->stream;
    const uint8_t* d_lay = nullptr;
    if (layouts) {
        DMFB_CUDA_TRY(cudaMemcpyAsync(h->d_layouts, layouts, N * A * 4, cudaMemcpyHostToDevice, s));
        d_lay = h->d_layouts;
    }
    const double* d_deg = nullptr;
    if (degrade && new_chip && h->degrade) {      // staged in the degrade array itself, read in place by the kernel
        DMFB_CUDA_TRY(cudaMemcpyAsync(h->degrade, degrade, N * cells * sizeof(double), cudaMemcpyHostToDevice, s));
        d_deg = h->degrade;
    }
    const meda_state_t st = state_of(h);
    int rc = meda_reset(&h->cfg, &st, nullptr, new_chip, d_lay, d_deg, seed, h->set_order, h->d_obs, s);
    if (rc) return rc;
    if (obs) DMFB_CUDA_TRY(cudaMemcpyAsync(obs, h->d_obs, N * A * D, cudaMemcpyDeviceToHost, s));
    int32_t gs = 0;
    DMFB_CUDA_TRY(cudaMemcpyAsync(&gs, h->gen_status, sizeof(gs), cudaMemcpyDeviceToHost, s));
    DMFB_CUDA_TRY(cudaStreamSynchronize(s));
    if (gs & DMFB_STATUS_SAMPLER_GAVE_UP) {
        DMFB_CUDA_TRY(cudaMemsetAsync(h->gen_status, 0, sizeof(gs), s));
        snprintf(g_last_error, sizeof(g_last_error), "meda task generator found no legal layout");
        return DMFB_ERR_BAD_ARG;
    }
    return DMFB_OK;
}

int meda_host_step(meda_host_env_t* h, const int8_t* actions, const double* u_inject, uint64_t seed, uint32_t flags,
                   int8_t* obs, float* reward, uint8_t* done, int32_t* constraints, uint8_t* success)
{
    if (!h || !actions || !obs) return DMFB_ERR_BAD_ARG;
    DMFB_CUDA_TRY(cudaSetDevice(h->device));
    const size_t N = (size_t)h->n_envs, A = (size_t)h->cfg.n_agents, D = (size_t)h->cfg.obs_dim;
    cudaStream_t s = h->stream;
    DMFB_CUDA_TRY(cudaMemcpyAsync(h->d_actions, actions, N * A, cudaMemcpyHostToDevice, s));
    const double* d_u = nullptr;
    if (u_inject) {
        DMFB_CUDA_TRY(cudaMemcpyAsync(h->d_u, u_inject, N * A * sizeof(double), cudaMemcpyHostToDevice, s));
        d_u = h->d_u;
    }
    const meda_state_t st = state_of(h);
    meda_out_t o{};
    o.obs = h->d_obs; o.reward = h->d_reward; o.done = h->d_done; o.constraints = h->d_cons; o.success = h->d_succ;
    int rc = meda_step(&h->cfg, &st, h->d_actions, 1, d_u, seed, flags, h->set_order, &o, s);
    if (rc) return rc;
    DMFB_CUDA_TRY(cudaMemcpyAsync(obs, h->d_obs, N * A * D, cudaMemcpyDeviceToHost, s));
    if (reward) DMFB_CUDA_TRY(cudaMemcpyAsync(reward, o.reward, N * A * sizeof(float), cudaMemcpyDeviceToHost, s));
    if (done) DMFB_CUDA_TRY(cudaMemcpyAsync(done, o.done, N * A, cudaMemcpyDeviceToHost, s));
    if (constraints) DMFB_CUDA_TRY(cudaMemcpyAsync(constraints, o.constraints, N * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    if (success) DMFB_CUDA_TRY(cudaMemcpyAsync(success, o.success, N, cudaMemcpyDeviceToHost, s));
    DMFB_CUDA_TRY(cudaStreamSynchronize(s));
    return DMFB_OK;
}

}  // extern "C"
