// Several shards inside ONE process (include/hac_index.h, hac_shards_*): what faiss' IndexShards does with its own
// C++ thread pool (index_cpu_to_gpu_multiple with co.shard, /root/reference/src/test_HAConvDR_topiocqa.py:55-66).
// One persistent host thread per shard drives that shard's device; the search itself is the single-shard one
// (hac_search_device) with the cross-shard threshold exchange armed, followed by ONE merge kernel on the first
// device that reads every shard's list in place over NVLink.  Built on the public C ABI only.
#include <float.h>
#include <string.h>

#include <chrono>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/hac_index.h"

namespace hac {
int set_last_error(int code, const std::string& msg);      // hac_api.cu: the calling thread's hac_last_error()
}

namespace {

constexpr int kMaxShards = 16;

double now_ms() {
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

struct DeviceScope {
    int prev = -1;
    explicit DeviceScope(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
    }
    ~DeviceScope() {
        int cur = -1;
        cudaGetDevice(&cur);
        if (prev >= 0 && cur != prev) cudaSetDevice(prev);
    }
};

struct Shard {
    hac_index* idx = nullptr;
    int device = 0;
    cudaStream_t stream = nullptr;
    float* q = nullptr;            // [q_cap] floats: this device's copy of the queries
    int64_t q_cap = 0;
    float* D = nullptr;            // [res_cap] this shard's result lists, read by the merge kernel
    int64_t* I = nullptr;
    int64_t res_cap = 0;
    uint64_t* words = nullptr;     // threshold-exchange words, written by this shard, read by the others
    std::thread th;
    int rc = HAC_OK;
    std::string err;
    double t_start_ms = 0.0, t_search_ms = 0.0, t_done_ms = 0.0;   // since the search call began
};

}  // namespace

struct hac_shards {
    int n = 0, d = 0;
    std::vector<Shard> sh;
    bool peer_ok = true;           // every device can load / store every other device's memory
    bool exchange = true;
    int64_t words_cap = 0;         // queries the exchange buffers hold
    int64_t epoch = 0;
    // merged result on the first device, and (no peer access) the gathered lists
    float* Dm = nullptr;
    int64_t* Im = nullptr;
    int64_t out_cap = 0;
    float* Dg = nullptr;
    int64_t* Ig = nullptr;
    int64_t gather_cap = 0;
    // job hand-off to the shard threads
    std::mutex mu;
    std::condition_variable cv_job, cv_done;
    uint64_t job_seq = 0;
    int pending = 0;
    bool stop = false;
    int64_t job_nq = 0;
    const float* job_q = nullptr;
    bool job_q_on_host = true;
    int job_k = 0;
    int64_t job_epoch = 0;
    double job_t0 = 0.0;
    float phase_ms[4] = {0, 0, 0, 0};   // shards done (slowest), merge done, results on the host, first shard done
};

namespace {

void shard_thread(hac_shards* g, int s) {
    Shard& me = g->sh[s];
    cudaSetDevice(me.device);
    uint64_t seen = 0;
    for (;;) {
        {
            std::unique_lock<std::mutex> lk(g->mu);
            g->cv_job.wait(lk, [&] { return g->stop || g->job_seq != seen; });
            if (g->stop) return;
            seen = g->job_seq;
        }
        me.t_start_ms = now_ms() - g->job_t0;
        const int64_t nq = g->job_nq;
        const size_t q_bytes = (size_t)nq * g->d * sizeof(float);
        int rc = HAC_OK;
        std::string err;
        // every shard fetches the queries itself: the copies run on G PCIe links (or NVLink, device source) at once
        cudaError_t e = cudaMemcpyAsync(me.q, g->job_q, q_bytes,
                                        g->job_q_on_host ? cudaMemcpyHostToDevice : cudaMemcpyDefault, me.stream);
        if (e != cudaSuccess) {
            rc = HAC_E_CUDA;
            err = std::string("query copy: ") + cudaGetErrorString(e);
            cudaGetLastError();
        }
        if (rc == HAC_OK) {
            hac_set_option(me.idx, "exchange_epoch", g->job_epoch);
            me.t_search_ms = now_ms() - g->job_t0;
            rc = hac_search_device(me.idx, nq, me.q, g->job_k, me.D, me.I, me.stream);
            if (rc != HAC_OK) err = hac_last_error();
            hac_set_option(me.idx, "exchange_epoch", 0);
        }
        me.rc = rc;
        me.err = err;
        me.t_done_ms = now_ms() - g->job_t0;
        {
            std::lock_guard<std::mutex> lk(g->mu);
            if (--g->pending == 0) g->cv_done.notify_all();
        }
    }
}

int fail_here(int code, const std::string& msg) { return hac::set_last_error(code, msg); }

#define CUS(call)                                                                                       \
    do {                                                                                                \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess) {                                                                       \
            cudaGetLastError();                                                                         \
            return fail_here(e__ == cudaErrorMemoryAllocation ? HAC_E_NOMEM : HAC_E_CUDA,               \
                             std::string(#call) + ": " + cudaGetErrorString(e__));                      \
        }                                                                                               \
    } while (0)

int64_t grow_to(int64_t need, int64_t floor_) {
    int64_t cap = floor_;
    while (cap < need) cap *= 2;
    return cap;
}

// per-search buffers; grown geometrically, never shrunk
int ensure_buffers(hac_shards* g, int64_t nq, int k) {
    const int64_t q_elems = nq * g->d, r_elems = nq * k;
    for (Shard& s : g->sh) {
        DeviceScope ds(s.device);
        if (q_elems > s.q_cap) {
            if (s.q) cudaFree(s.q);
            s.q = nullptr;
            s.q_cap = 0;
            const int64_t cap = grow_to(q_elems, 256 * (int64_t)g->d);
            CUS(cudaMalloc(&s.q, (size_t)cap * sizeof(float)));
            s.q_cap = cap;
        }
        if (r_elems > s.res_cap) {
            if (s.D) cudaFree(s.D);
            if (s.I) cudaFree(s.I);
            s.D = nullptr;
            s.I = nullptr;
            s.res_cap = 0;
            const int64_t cap = grow_to(r_elems, 1 << 16);
            CUS(cudaMalloc(&s.D, (size_t)cap * sizeof(float)));
            CUS(cudaMalloc(&s.I, (size_t)cap * sizeof(int64_t)));
            s.res_cap = cap;
        }
    }
    {
        DeviceScope ds(g->sh[0].device);
        if (r_elems > g->out_cap) {
            if (g->Dm) cudaFree(g->Dm);
            if (g->Im) cudaFree(g->Im);
            g->Dm = nullptr;
            g->Im = nullptr;
            g->out_cap = 0;
            const int64_t cap = grow_to(r_elems, 1 << 16);
            CUS(cudaMalloc(&g->Dm, (size_t)cap * sizeof(float)));
            CUS(cudaMalloc(&g->Im, (size_t)cap * sizeof(int64_t)));
            g->out_cap = cap;
        }
        if (!g->peer_ok && r_elems * g->n > g->gather_cap) {
            if (g->Dg) cudaFree(g->Dg);
            if (g->Ig) cudaFree(g->Ig);
            g->Dg = nullptr;
            g->Ig = nullptr;
            g->gather_cap = 0;
            const int64_t cap = grow_to(r_elems * g->n, 1 << 16);
            CUS(cudaMalloc(&g->Dg, (size_t)cap * sizeof(float)));
            CUS(cudaMalloc(&g->Ig, (size_t)cap * sizeof(int64_t)));
            g->gather_cap = cap;
        }
    }
    if (g->exchange && g->peer_ok && g->n > 1 && nq > g->words_cap) {
        // new (zeroed) words on every device, then every shard learns its own and its peers' buffers
        const int64_t cap = grow_to(nq, 4096);
        for (Shard& s : g->sh) {
            DeviceScope ds(s.device);
            hac_set_threshold_exchange(s.idx, nullptr, nullptr, 0, 0);
            if (s.words) cudaFree(s.words);
            s.words = nullptr;
        }
        g->words_cap = 0;
        for (Shard& s : g->sh) {
            DeviceScope ds(s.device);
            const size_t bytes = (size_t)cap * HAC_EXCHANGE_WORDS_PER_QUERY * sizeof(uint64_t);
            CUS(cudaMalloc(&s.words, bytes));
            CUS(cudaMemset(s.words, 0, bytes));
            CUS(cudaDeviceSynchronize());
        }
        for (int a = 0; a < g->n; ++a) {
            const uint64_t* peers[kMaxShards];
            int np = 0;
            for (int b = 0; b < g->n; ++b)
                if (b != a) peers[np++] = g->sh[b].words;
            int rc = hac_set_threshold_exchange(g->sh[a].idx, g->sh[a].words, peers, np, cap * HAC_EXCHANGE_WORDS_PER_QUERY);
            if (rc != HAC_OK) return rc;
        }
        g->words_cap = cap;
    }
    return HAC_OK;
}

int search_impl(hac_shards* g, int64_t nq, const float* q, bool q_on_host, int k, float* D, int64_t* I, bool out_on_host,
                cudaStream_t out_stream) {
    if (g == nullptr) return fail_here(HAC_E_INVALID, "shards_search: null group");
    if (nq < 0 || (nq > 0 && (!q || !D || !I))) return fail_here(HAC_E_INVALID, "shards_search: null buffer");
    if (k <= 0 || k > HAC_MAX_K) return fail_here(HAC_E_INVALID, "shards_search: k outside [1, HAC_MAX_K]");
    if ((int64_t)g->n * k > 16384) return fail_here(HAC_E_INVALID, "shards_search: n_shards * k exceeds 16384");
    if (nq == 0) return HAC_OK;
    const double t0 = now_ms();
    int rc = ensure_buffers(g, nq, k);
    if (rc != HAC_OK) return rc;
    g->epoch = g->epoch % 0x0FFFFFFF + 1;                  // 1 .. 2^28 - 1, never 0 (= off)
    const bool armed = g->exchange && g->peer_ok && g->n > 1 && g->words_cap >= nq;
    {
        std::lock_guard<std::mutex> lk(g->mu);
        g->job_nq = nq;
        g->job_q = q;
        g->job_q_on_host = q_on_host;
        g->job_k = k;
        g->job_epoch = armed ? g->epoch : 0;
        g->job_t0 = t0;
        g->pending = g->n;
        ++g->job_seq;
    }
    g->cv_job.notify_all();
    {
        std::unique_lock<std::mutex> lk(g->mu);
        g->cv_done.wait(lk, [&] { return g->pending == 0; });
    }
    double first = 1e30, last = 0.0;
    for (const Shard& s : g->sh) {
        first = std::min(first, s.t_done_ms);
        last = std::max(last, s.t_done_ms);
    }
    g->phase_ms[0] = (float)last;
    g->phase_ms[3] = (float)first;
    for (int s = 0; s < g->n; ++s)                          // every shard has finished: report the first failure
        if (g->sh[s].rc != HAC_OK)
            return fail_here(g->sh[s].rc, "shard " + std::to_string(s) + " (cuda:" + std::to_string(g->sh[s].device) +
                                              "): " + g->sh[s].err);
    // the shard searches returned after their streams drained, so the lists are complete: one merge kernel on the
    // first device reads them where they lie
    Shard& s0 = g->sh[0];
    DeviceScope ds(s0.device);
    cudaStream_t ms = out_on_host ? s0.stream : out_stream;
    float* Dout = out_on_host ? g->Dm : D;
    int64_t* Iout = out_on_host ? g->Im : I;
    if (g->peer_ok) {
        const float* Dp[kMaxShards];
        const int64_t* Ip[kMaxShards];
        for (int s = 0; s < g->n; ++s) {
            Dp[s] = g->sh[s].D;
            Ip[s] = g->sh[s].I;
        }
        rc = hac_merge_topk_peers_device(s0.device, g->n, nq, k, Dp, Ip, k, Dout, Iout, ms);
    } else {
        const size_t r_elems = (size_t)nq * k;
        for (int s = 0; s < g->n; ++s) {
            CUS(cudaMemcpyPeerAsync(g->Dg + s * r_elems, s0.device, g->sh[s].D, g->sh[s].device, r_elems * sizeof(float), ms));
            CUS(cudaMemcpyPeerAsync(g->Ig + s * r_elems, s0.device, g->sh[s].I, g->sh[s].device, r_elems * sizeof(int64_t), ms));
        }
        rc = hac_merge_topk_device(s0.device, g->n, nq, k, g->Dg, g->Ig, k, Dout, Iout, ms);
    }
    if (rc != HAC_OK) return rc;
    if (out_on_host) {
        CUS(cudaMemcpyAsync(D, Dout, (size_t)nq * k * sizeof(float), cudaMemcpyDeviceToHost, ms));
        g->phase_ms[1] = (float)(now_ms() - t0);           // merge enqueued (the D2H copies follow it on the stream)
        CUS(cudaMemcpyAsync(I, Iout, (size_t)nq * k * sizeof(int64_t), cudaMemcpyDeviceToHost, ms));
    }
    CUS(cudaStreamSynchronize(ms));
    g->phase_ms[2] = (float)(now_ms() - t0);
    if (!out_on_host) g->phase_ms[1] = g->phase_ms[2];
    return HAC_OK;
}

}  // namespace

extern "C" {

int hac_shards_create(hac_index* const* shards, int n, hac_shards** out) {
    if (out == nullptr) return fail_here(HAC_E_INVALID, "shards_create: null out pointer");
    *out = nullptr;
    if (shards == nullptr || n < 1 || n > kMaxShards) return fail_here(HAC_E_INVALID, "shards_create: 1..16 shards");
    for (int i = 0; i < n; ++i) {
        if (shards[i] == nullptr) return fail_here(HAC_E_INVALID, "shards_create: null shard");
        if (hac_dim(shards[i]) != hac_dim(shards[0])) return fail_here(HAC_E_INVALID, "shards_create: dimensions differ");
        for (int j = 0; j < i; ++j)
            if (shards[j] == shards[i]) return fail_here(HAC_E_INVALID, "shards_create: the same shard twice");
    }
    hac_shards* g = new hac_shards();
    g->n = n;
    g->d = hac_dim(shards[0]);
    g->sh.resize(n);
    for (int i = 0; i < n; ++i) {
        g->sh[i].idx = shards[i];
        g->sh[i].device = hac_device(shards[i]);
    }
    // kernels of a read (merge) and write (threshold exchange) memory of b
    for (int a = 0; a < n; ++a)
        for (int b = 0; b < n; ++b)
            if (g->sh[a].device != g->sh[b].device && hac_enable_peer_access(g->sh[a].device, g->sh[b].device) != HAC_OK)
                g->peer_ok = false;
    for (int i = 0; i < n; ++i) {
        DeviceScope ds(g->sh[i].device);
        cudaError_t e = cudaStreamCreateWithFlags(&g->sh[i].stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) {
            cudaGetLastError();
            for (int j = 0; j < i; ++j) cudaStreamDestroy(g->sh[j].stream);
            delete g;
            return fail_here(HAC_E_CUDA, std::string("shards_create: ") + cudaGetErrorString(e));
        }
    }
    for (int i = 0; i < n; ++i) g->sh[i].th = std::thread(shard_thread, g, i);
    *out = g;
    return HAC_OK;
}

int hac_shards_destroy(hac_shards* g) {
    if (g == nullptr) return HAC_OK;
    {
        std::lock_guard<std::mutex> lk(g->mu);
        g->stop = true;
    }
    g->cv_job.notify_all();
    for (Shard& s : g->sh)
        if (s.th.joinable()) s.th.join();
    for (Shard& s : g->sh) {
        DeviceScope ds(s.device);
        cudaStreamSynchronize(s.stream);
        hac_set_threshold_exchange(s.idx, nullptr, nullptr, 0, 0);   // the shard handles outlive the group
        if (s.q) cudaFree(s.q);
        if (s.D) cudaFree(s.D);
        if (s.I) cudaFree(s.I);
        if (s.words) cudaFree(s.words);
        cudaStreamDestroy(s.stream);
    }
    {
        DeviceScope ds(g->sh[0].device);
        if (g->Dm) cudaFree(g->Dm);
        if (g->Im) cudaFree(g->Im);
        if (g->Dg) cudaFree(g->Dg);
        if (g->Ig) cudaFree(g->Ig);
    }
    delete g;
    return HAC_OK;
}

int hac_shards_search(hac_shards* g, int64_t nq, const float* q_host, int k, float* D_host, int64_t* I_host) {
    return search_impl(g, nq, q_host, true, k, D_host, I_host, true, nullptr);
}

int hac_shards_search_device(hac_shards* g, int64_t nq, const float* q_dev, int k, float* D_dev, int64_t* I_dev,
                             void* stream) {
    if (g != nullptr && nq > 0 && q_dev != nullptr) {
        // the queries are read by every shard's own stream: the caller's stream must have produced them
        DeviceScope ds(g->sh[0].device);
        CUS(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
    }
    return search_impl(g, nq, q_dev, false, k, D_dev, I_dev, false, static_cast<cudaStream_t>(stream));
}

int hac_shards_set_exchange(hac_shards* g, int on) {
    if (g == nullptr) return fail_here(HAC_E_INVALID, "shards_set_exchange: null group");
    g->exchange = on != 0;
    return HAC_OK;
}

int hac_shards_peer_access(const hac_shards* g) { return g != nullptr && g->peer_ok ? 1 : 0; }

int hac_shards_last_phases(const hac_shards* g, float* out_ms, int n) {
    if (g == nullptr || out_ms == nullptr || n < 0) return fail_here(HAC_E_INVALID, "shards_last_phases: bad argument");
    for (int i = 0; i < n && i < 4; ++i) out_ms[i] = g->phase_ms[i];
    return HAC_OK;
}

}  // extern "C"
