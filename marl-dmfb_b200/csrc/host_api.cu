// host_api.cu — the same DMFB hot path behind HOST buffers (dmfb_host_* in include/dmfb_b200.h).
//
// This is the reference-facing shape of the call: like DMFBenv.step (env/DMFB/dmfb.py:560-587) it takes
// actions from host memory and returns observations / rewards / dones / info in host memory.  The handle
// owns the device-resident state; every call copies inputs H2D and results D2H.  The env batch is cut
// into chunks, one CUDA stream each, so that the D2H copy of chunk c overlaps the kernels of chunk c+1
// (PCIe is the bound of this path: 245 obs bytes per agent-step have to cross it).
//
// Packed transfer (dmfb_host_set_transfer): observation cells are small integers (0..n_agents), so a share of the
// batch crosses PCIe as 4-bit cells (half the bytes) and is expanded into the caller's buffer by the host cores,
// WHILE the rest of the batch arrives unpacked by DMA: the two writers of the host buffer - copy engine and CPU
// threads - work in parallel, which beats either alone.
#include <new>
#include <vector>

#include "common.cuh"
#include "host_unpack.h"

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
    uint16_t* usage_log = nullptr;      // [N, max_step, A] actuated cells since the last reset (dmfb_state_t)
    int32_t* usage_log_len = nullptr;
    uint32_t* next_task = nullptr;      // task prefetch (dmfb_state_t.next_task / next_cursor)
    uint32_t* next_cursor = nullptr;
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
    // packed transfer (off while pool == nullptr)
    dmfb::UnpackPool* pool = nullptr;
    int dma_percent = 100;
    uint8_t* d_packed = nullptr;   // [N*A][packed_record_bytes]
    uint8_t* h_packed = nullptr;   // pinned mirror
    cudaStream_t dma_stream = nullptr;
    cudaEvent_t ev_step = nullptr;
    std::vector<cudaEvent_t> ev_chunk;
};

constexpr int kPackChunks = 16;

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
    s.usage_log_cap = h->usage_log ? h->cfg.max_step : 0;
    s.usage_log = h->usage_log ? h->usage_log + (size_t)lo * h->cfg.max_step * A : nullptr;
    s.usage_log_len = h->usage_log_len ? h->usage_log_len + lo : nullptr;
    s.next_task = h->next_task + (size_t)lo * A;
    s.next_cursor = h->next_cursor + lo;
    return s;
}

// One thread per 32-bit word of the packed records: 8 cells (or the direction bytes / padding at the record's end).
__global__ void __launch_bounds__(256)
pack_obs_kernel(const int8_t* __restrict__ obs, uint8_t* __restrict__ packed, int cells, int words_per_record,
                size_t n_words)
{
    const size_t gw = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gw >= n_words) return;
    const size_t rec = gw / (size_t)words_per_record;
    const int w = (int)(gw - rec * (size_t)words_per_record);
    const int8_t* in = obs + rec * (size_t)(cells + 2);
    const int nb = (cells + 1) >> 1;
    uint32_t v = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int b = 4 * w + q;
        uint32_t byte = 0;
        if (b < nb) {
            const uint32_t lo = (uint32_t)in[2 * b] & 15u;
            const uint32_t hi = (2 * b + 1 < cells) ? ((uint32_t)in[2 * b + 1] & 15u) : 0u;
            byte = lo | (hi << 4);
        } else if (b < nb + 2) {
            byte = (uint32_t)(uint8_t)in[cells + (b - nb)];
        }
        v |= byte << (8 * q);
    }
    reinterpret_cast<uint32_t*>(packed)[gw] = v;
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
    TRY_ALLOC(next_task, N * A)
    TRY_ALLOC(next_cursor, N)
    if (cfg->n_blocks) TRY_ALLOC(blocks, N * (size_t)cfg->n_blocks * 2)
    if (cfg->b_degrade) {
        TRY_ALLOC(usage, N * cells)
        TRY_ALLOC(health, N * cells)
        TRY_ALLOC(degrade, N * cells)
        TRY_ALLOC(usage_log, N * (size_t)cfg->max_step * A)
        TRY_ALLOC(usage_log_len, N)
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
    dmfb::unpack_pool_destroy(h->pool);
    if (h->dma_stream) cudaStreamDestroy(h->dma_stream);
    if (h->ev_step) cudaEventDestroy(h->ev_step);
    for (cudaEvent_t e : h->ev_chunk) cudaEventDestroy(e);
    if (h->d_packed) cudaFree(h->d_packed);
    if (h->h_packed) cudaFreeHost(h->h_packed);
    void* ptrs[] = {h->drop, h->start, h->terminated, h->step_count, h->constraints, h->episode, h->usage, h->health,
                    h->degrade, h->blocks, h->usage_log, h->usage_log_len, h->next_task, h->next_cursor, h->d_actions, h->d_u, h->d_obs, h->d_reward, h->d_done, h->d_cons, h->d_succ,
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

int dmfb_host_set_transfer(dmfb_host_env_t* h, int n_threads, int dma_percent)
{
    if (!h || dma_percent < 0 || dma_percent > 100) return DMFB_ERR_BAD_ARG;
    DMFB_CUDA_TRY(cudaSetDevice(h->device));
    if (n_threads <= 0) {                      // back to the plain DMA path
        dmfb::unpack_pool_destroy(h->pool);
        h->pool = nullptr;
        return DMFB_OK;
    }
    if (h->cfg.n_agents > 15) {                // cell values 0..n_agents must fit 4 bits
        snprintf(g_last_error, sizeof(g_last_error), "packed transfer needs n_agents <= 15");
        return DMFB_ERR_BAD_ARG;
    }
    const size_t recs = (size_t)h->n_envs * h->cfg.n_agents;
    const size_t bytes = recs * dmfb::packed_record_bytes(h->cfg.obs_dim - 2);
    if (!h->d_packed) {
        DMFB_CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&h->d_packed), bytes));
        DMFB_CUDA_TRY(cudaHostAlloc(reinterpret_cast<void**>(&h->h_packed), bytes, cudaHostAllocDefault));
        DMFB_CUDA_TRY(cudaStreamCreateWithFlags(&h->dma_stream, cudaStreamNonBlocking));
        DMFB_CUDA_TRY(cudaEventCreateWithFlags(&h->ev_step, cudaEventDisableTiming));
        for (int c = 0; c < kPackChunks; ++c) {
            cudaEvent_t e;
            DMFB_CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            h->ev_chunk.push_back(e);
        }
    }
    dmfb::unpack_pool_destroy(h->pool);
    h->pool = dmfb::unpack_pool_create(n_threads);
    h->dma_percent = dma_percent;
    return DMFB_OK;
}

int dmfb_host_unpack_records(const uint8_t* packed, size_t packed_stride, int8_t* out, int cells, size_t n_records,
                             int n_threads)
{
    if (!packed || !out || cells < 1 || packed_stride < (size_t)(cells + 1) / 2 + 2) return DMFB_ERR_BAD_ARG;
    dmfb::UnpackPool* pool = dmfb::unpack_pool_create(n_threads);
    dmfb::unpack_records(pool, packed, packed_stride, out, cells, n_records);
    dmfb::unpack_pool_destroy(pool);
    return DMFB_OK;
}

// dmfb_host_step with the packed transfer on: one launch for the whole batch, then envs [0, n_dma) leave by DMA
// straight into the caller's buffer while envs [n_dma, N) are packed on the device, copied in kPackChunks pieces
// and expanded by the pool as the pieces arrive.
static int host_step_packed(dmfb_host_env_t* h, const int8_t* actions, const double* u_inject, uint64_t seed,
                            uint32_t flags, int8_t* obs, float* reward, uint8_t* done, int32_t* constraints,
                            uint8_t* success)
{
    const size_t N = (size_t)h->n_envs, A = (size_t)h->cfg.n_agents, D = (size_t)h->cfg.obs_dim;
    const int cells = (int)D - 2;
    const size_t stride = dmfb::packed_record_bytes(cells);
    cudaStream_t s = h->streams[0], sd = h->dma_stream;
    size_t n_dma = (N * (size_t)h->dma_percent / 100) & ~(size_t)63;    // 64-env granularity keeps 16-byte alignment
    if (h->dma_percent == 100) n_dma = N;
    DMFB_CUDA_TRY(cudaMemcpyAsync(h->d_actions, actions, N * A, cudaMemcpyHostToDevice, s));
    const double* d_u = nullptr;
    if (u_inject) {
        DMFB_CUDA_TRY(cudaMemcpyAsync(h->d_u, u_inject, N * A * sizeof(double), cudaMemcpyHostToDevice, s));
        d_u = h->d_u;
    }
    dmfb_state_t st = sub_state(h, 0, h->n_envs);
    dmfb_out_t o{};
    o.obs = h->d_obs; o.reward = h->d_reward; o.done = h->d_done; o.constraints = h->d_cons; o.success = h->d_succ;
    int rc = dmfb_step(&h->cfg, &st, h->d_actions, 1, d_u, seed, flags, &o, s);
    if (rc) return rc;
    DMFB_CUDA_TRY(cudaEventRecord(h->ev_step, s));
    // unpacked share + the small outputs: second stream, so that the copy engine runs beside the pack kernel
    DMFB_CUDA_TRY(cudaStreamWaitEvent(sd, h->ev_step, 0));
    if (n_dma) DMFB_CUDA_TRY(cudaMemcpyAsync(obs, h->d_obs, n_dma * A * D, cudaMemcpyDeviceToHost, sd));
    if (reward) DMFB_CUDA_TRY(cudaMemcpyAsync(reward, o.reward, N * A * sizeof(float), cudaMemcpyDeviceToHost, sd));
    if (done) DMFB_CUDA_TRY(cudaMemcpyAsync(done, o.done, N * A, cudaMemcpyDeviceToHost, sd));
    if (constraints) DMFB_CUDA_TRY(cudaMemcpyAsync(constraints, o.constraints, N * sizeof(int32_t), cudaMemcpyDeviceToHost, sd));
    if (success) DMFB_CUDA_TRY(cudaMemcpyAsync(success, o.success, N, cudaMemcpyDeviceToHost, sd));
    // packed share
    const size_t rec0 = n_dma * A, n_rec = (N - n_dma) * A;
    size_t bound[kPackChunks + 1];
    for (int c = 0; c <= kPackChunks; ++c) bound[c] = rec0 + n_rec * (size_t)c / kPackChunks;
    if (n_rec) {
        const int wpr = (int)(stride / 4);
        for (int c = 0; c < kPackChunks; ++c) {                               // pack + copy piece by piece: the first
            const size_t q0 = bound[c], q1 = bound[c + 1];                    // piece reaches the host threads early
            if (q1 > q0) {
                const size_t n_words = (q1 - q0) * (size_t)wpr;
                pack_obs_kernel<<<(unsigned)((n_words + 255) / 256), 256, 0, s>>>(h->d_obs + q0 * D, h->d_packed + q0 * stride,
                                                                                 cells, wpr, n_words);
                g_launches.fetch_add(1);
                DMFB_CUDA_TRY(cudaGetLastError());
                DMFB_CUDA_TRY(cudaMemcpyAsync(h->h_packed + q0 * stride, h->d_packed + q0 * stride, (q1 - q0) * stride,
                                              cudaMemcpyDeviceToHost, s));
            }
            DMFB_CUDA_TRY(cudaEventRecord(h->ev_chunk[c], s));
        }
        for (int c = 0; c < kPackChunks; ++c) {
            DMFB_CUDA_TRY(cudaEventSynchronize(h->ev_chunk[c]));
            dmfb::unpack_records(h->pool, h->h_packed + bound[c] * stride, stride, obs + bound[c] * D, cells,
                                 bound[c + 1] - bound[c]);
        }
    }
    DMFB_CUDA_TRY(cudaStreamSynchronize(sd));
    DMFB_CUDA_TRY(cudaStreamSynchronize(s));
    return DMFB_OK;
}

int dmfb_host_step(dmfb_host_env_t* h, const int8_t* actions, const double* u_inject, uint64_t seed, uint32_t flags,
                   int8_t* obs, float* reward, uint8_t* done, int32_t* constraints, uint8_t* success)
{
    if (!h || !actions || !obs) return DMFB_ERR_BAD_ARG;
    DMFB_CUDA_TRY(cudaSetDevice(h->device));
    if (h->pool) return host_step_packed(h, actions, u_inject, seed, flags, obs, reward, done, constraints, success);
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
