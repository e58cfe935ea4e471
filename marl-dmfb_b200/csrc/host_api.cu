// host_api.cu — the same DMFB hot path behind HOST buffers (dmfb_host_* in include/dmfb_b200.h).
//
// This is the reference-facing shape of the call: like DMFBenv.step (env/DMFB/dmfb.py:560-587) it takes
// actions from host memory and returns observations / rewards / dones / info in host memory.  The handle
// owns the device-resident state; every call copies inputs H2D and results D2H.  The env batch is cut
// into chunks, one CUDA stream each, so that the D2H copy of chunk c overlaps the kernels of chunk c+1
// (PCIe is the bound of this path: 245 obs bytes per agent-step have to cross it).
#include <new>
#include <vector>

#include "common.cuh"

using namespace dmfb;

struct dmfb_host_env {
    dmfb_cfg_t cfg;
    int n_envs = 0, device = 0, n_chunks = 1;
    // device state
    uint8_t *drop = nullptr, *start = nullptr, *terminated = nullptr;
    int32_t *step_count = nullptr, *constraints = nullptr;
    uint32_t* episode = nullptr;
    uint32_t* usage = nullptr;
    uint8_t* blocks = nullptr;
    double *health = nullptr, *degrade = nullptr;
    // device staging of per-step inputs / outputs
    int8_t* d_actions = nullptr;
    double* d_u = nullptr;
    int8_t* d_obs = nullptr;
    float* d_reward = nullptr;
    uint8_t* d_done = nullptr;
    int32_t* d_cons = nullptr;
    uint8_t* d_succ = nullptr;
    uint8_t* d_layouts = nullptr;
    std::vector<cudaStream_t> streams;
    std::vector<int> lo;  // chunk boundaries, size n_chunks+1
};

namespace {

template <typename T>
int dev_alloc(T** p, size_t count)
{
    DMFB_CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(p), count * sizeof(T)));
    DMFB_CUDA_TRY(cudaMemset(*p, 0, count * sizeof(T)));
    return DMFB_OK;
}

dmfb_state_t sub_state(const dmfb_host_env* h, int lo, int hi)
{
    const size_t A = (size_t)h->cfg.n_agents, cells = (size_t)h->cfg.width * h->cfg.length;
    dmfb_state_t s{};
    s.n_envs = hi - lo;
    s.drop = h->drop + (size_t)lo * A * 4;
    s.start = h->start + (size_t)lo * A * 2;
    s.step_count = h->step_count + lo;
    s.constraints = h->constraints + lo;
    s.terminated = h->terminated + lo;
    s.episode = h->episode + lo;
    s.usage = h->usage ? h->usage + (size_t)lo * cells : nullptr;
    s.health = h->health ? h->health + (size_t)lo * cells : nullptr;
    s.degrade = h->degrade ? h->degrade + (size_t)lo * cells : nullptr;
    s.blocks = h->blocks ? h->blocks + (size_t)lo * h->cfg.n_blocks * 2 : nullptr;
    return s;
}

}  // namespace

extern "C" {

void* dmfb_host_alloc_pinned(size_t bytes)
{
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    return p;
}

void dmfb_host_free_pinned(void* p)
{
    if (p) cudaFreeHost(p);
}

int dmfb_host_create(const dmfb_cfg_t* cfg, int n_envs, int device, int n_chunks, dmfb_host_env_t** out)
{
    if (!cfg || !out || n_envs <= 0) return DMFB_ERR_BAD_ARG;
    DMFB_CUDA_TRY(cudaSetDevice(device));
    dmfb_host_env* h = new (std::nothrow) dmfb_host_env();
    if (!h) return DMFB_ERR_BAD_ARG;
    h->cfg = *cfg;
    h->n_envs = n_envs;
    h->device = device;
    // chunk boundaries on multiples of 64 envs keep every chunk's obs slice 16-byte aligned
    const int units = (n_envs + 63) / 64;
    if (n_chunks < 1) n_chunks = 1;
    if (n_chunks > units) n_chunks = units;
    h->n_chunks = n_chunks;
    for (int c = 0; c <= n_chunks; ++c) {
        long long b = (long long)units * c / n_chunks * 64;
        h->lo.push_back((int)(b > n_envs ? n_envs : b));
    }
    h->lo[n_chunks] = n_envs;
    const size_t N = (size_t)n_envs, A = (size_t)cfg->n_agents, cells = (size_t)cfg->width * cfg->length;
    int rc = DMFB_OK;
#define TRY_ALLOC(p, n) if ((rc = dev_alloc(&h->p, (n))) != DMFB_OK) { dmfb_host_destroy(h); return rc; }
    TRY_ALLOC(drop, N * A * 4)
    TRY_ALLOC(start, N * A * 2)
    TRY_ALLOC(terminated, N)
    TRY_ALLOC(step_count, N)
    TRY_ALLOC(constraints, N)
    TRY_ALLOC(episode, N)
    if (cfg->n_blocks) TRY_ALLOC(blocks, N * (size_t)cfg->n_blocks * 2)
    if (cfg->b_degrade) {
        TRY_ALLOC(usage, N * cells)
        TRY_ALLOC(health, N * cells)
        TRY_ALLOC(degrade, N * cells)
    }
    TRY_ALLOC(d_actions, N * A)
    TRY_ALLOC(d_u, N * A)
    TRY_ALLOC(d_obs, N * A * (size_t)cfg->obs_dim)
    TRY_ALLOC(d_reward, N * A)
    TRY_ALLOC(d_done, N * A)
    TRY_ALLOC(d_cons, N)
    TRY_ALLOC(d_succ, N)
    TRY_ALLOC(d_layouts, N * A * 4)
#undef TRY_ALLOC
    for (int c = 0; c < n_chunks; ++c) {
        cudaStream_t s;
        cudaError_t e = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
        if (e != cudaSuccess) { dmfb_host_destroy(h); return cuda_fail(e, "cudaStreamCreate"); }
        h->streams.push_back(s);
    }
    *out = h;
    return DMFB_OK;
}

void dmfb_host_destroy(dmfb_host_env_t* h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    for (cudaStream_t s : h->streams) cudaStreamDestroy(s);
    void* ptrs[] = {h->drop, h->start, h->terminated, h->step_count, h->constraints, h->episode, h->usage, h->health,
                    h->degrade, h->blocks, h->d_actions, h->d_u, h->d_obs, h->d_reward, h->d_done, h->d_cons, h->d_succ,
                    h->d_layouts};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    delete h;
}

int dmfb_host_reset(dmfb_host_env_t* h, int new_task, const uint8_t* layouts, const double* degrade, uint64_t seed,
                    int8_t* obs)
{
    if (!h) return DMFB_ERR_BAD_ARG;
    DMFB_CUDA_TRY(cudaSetDevice(h->device));
    const size_t A = (size_t)h->cfg.n_agents, D = (size_t)h->cfg.obs_dim, cells = (size_t)h->cfg.width * h->cfg.length;
    for (int c = 0; c < h->n_chunks; ++c) {
        const int lo = h->lo[c], hi = h->lo[c + 1];
        if (hi <= lo) continue;
        cudaStream_t s = h->streams[c];
        dmfb_cfg_t cfg = h->cfg;
        cfg.env_base += lo;
        dmfb_state_t st = sub_state(h, lo, hi);
        const uint8_t* d_lay = nullptr;
        if (layouts) {
            DMFB_CUDA_TRY(cudaMemcpyAsync(h->d_layouts + (size_t)lo * A * 4, layouts + (size_t)lo * A * 4,
                                          (size_t)(hi - lo) * A * 4, cudaMemcpyHostToDevice, s));
            d_lay = h->d_layouts + (size_t)lo * A * 4;
        }
        const double* d_deg = nullptr;
        if (degrade && new_task && h->degrade) {
            // stage the injected factors in the degrade array itself, then let the kernel read them in place
            DMFB_CUDA_TRY(cudaMemcpyAsync(h->degrade + (size_t)lo * cells, degrade + (size_t)lo * cells,
                                          (size_t)(hi - lo) * cells * sizeof(double), cudaMemcpyHostToDevice, s));
            d_deg = h->degrade + (size_t)lo * cells;
        }
        int rc = dmfb_reset(&cfg, &st, nullptr, new_task, d_lay, nullptr, d_deg, seed, h->d_obs + (size_t)lo * A * D, s);
        if (rc) return rc;
        if (obs)
            DMFB_CUDA_TRY(cudaMemcpyAsync(obs + (size_t)lo * A * D, h->d_obs + (size_t)lo * A * D,
                                          (size_t)(hi - lo) * A * D, cudaMemcpyDeviceToHost, s));
    }
    for (cudaStream_t s : h->streams) DMFB_CUDA_TRY(cudaStreamSynchronize(s));
    return DMFB_OK;
}

int dmfb_host_step(dmfb_host_env_t* h, const int8_t* actions, const double* u_inject, uint64_t seed, uint32_t flags,
                   int8_t* obs, float* reward, uint8_t* done, int32_t* constraints, uint8_t* success)
{
    if (!h || !actions || !obs) return DMFB_ERR_BAD_ARG;
    DMFB_CUDA_TRY(cudaSetDevice(h->device));
    const size_t A = (size_t)h->cfg.n_agents, D = (size_t)h->cfg.obs_dim;
    for (int c = 0; c < h->n_chunks; ++c) {
        const int lo = h->lo[c], hi = h->lo[c + 1];
        if (hi <= lo) continue;
        const size_t n = (size_t)(hi - lo);
        cudaStream_t s = h->streams[c];
        dmfb_cfg_t cfg = h->cfg;
        cfg.env_base += lo;
        dmfb_state_t st = sub_state(h, lo, hi);
        DMFB_CUDA_TRY(cudaMemcpyAsync(h->d_actions + (size_t)lo * A, actions + (size_t)lo * A, n * A,
                                      cudaMemcpyHostToDevice, s));
        const double* d_u = nullptr;
        if (u_inject) {
            DMFB_CUDA_TRY(cudaMemcpyAsync(h->d_u + (size_t)lo * A, u_inject + (size_t)lo * A, n * A * sizeof(double),
                                          cudaMemcpyHostToDevice, s));
            d_u = h->d_u + (size_t)lo * A;
        }
        dmfb_out_t o{};
        o.obs = h->d_obs + (size_t)lo * A * D;
        o.reward = h->d_reward + (size_t)lo * A;
        o.done = h->d_done + (size_t)lo * A;
        o.constraints = h->d_cons + lo;
        o.success = h->d_succ + lo;
        int rc = dmfb_step(&cfg, &st, h->d_actions + (size_t)lo * A, 1, d_u, seed, flags, &o, s);
        if (rc) return rc;
        DMFB_CUDA_TRY(cudaMemcpyAsync(obs + (size_t)lo * A * D, o.obs, n * A * D, cudaMemcpyDeviceToHost, s));
        if (reward)
            DMFB_CUDA_TRY(cudaMemcpyAsync(reward + (size_t)lo * A, o.reward, n * A * sizeof(float), cudaMemcpyDeviceToHost, s));
        if (done) DMFB_CUDA_TRY(cudaMemcpyAsync(done + (size_t)lo * A, o.done, n * A, cudaMemcpyDeviceToHost, s));
        if (constraints)
            DMFB_CUDA_TRY(cudaMemcpyAsync(constraints + lo, o.constraints, n * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
        if (success) DMFB_CUDA_TRY(cudaMemcpyAsync(success + lo, o.success, n, cudaMemcpyDeviceToHost, s));
    }
    for (cudaStream_t s : h->streams) DMFB_CUDA_TRY(cudaStreamSynchronize(s));
    return DMFB_OK;
}

}  // extern "C"
