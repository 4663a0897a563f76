// ptb200.cu -- host side of libptb200.so: the C ABI declared in include/ptb200.h.
//
// Owns the handle (tables, dataset, per-vector workspaces, staging buffers), validates shapes the
// way the reference's Python layer does, stages host inputs through one pinned block, and launches
// the kernels in ptb_kernels.cuh / ptb_ts_kernels.cuh.  No CPU compute path exists here: without a
// CUDA device every computing entry point fails with PTB_ECUDA.
#include "../../include/ptb200.h"

#include <cuda_runtime.h>
#include <sys/mman.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "ptb_kernels.cuh"
#include "ptb_ss_kernels.cuh"
#include "ptb_ts_kernels.cuh"

using namespace ptb;

namespace {

thread_local std::string g_create_error;
// slots of the handle's device counter block d_work (ints): the persistent points kernel owns 0 and 1
constexpr int WORK_FINISH = 8, WORK_GATHER_ERR = 9;
constexpr unsigned long long GATHER_TIMEOUT_NS = 20ull * 1000000000ull;
std::atomic<unsigned long long> g_alloc_gen{1};  // bumped whenever a device or pinned buffer moves: captured graphs hold raw pointers

struct DevBuf {  // grow-only device buffer
    void *ptr = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        cap = 0;
        const size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&ptr, want);
        if (e == cudaSuccess) cap = want;
        ++g_alloc_gen;
        return e;
    }
    void release() {
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        cap = 0;
    }
    template <class T>
    T *as() const { return static_cast<T *>(ptr); }
};

struct PinBuf {  // grow-only pinned host buffer
    void *ptr = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (ptr) cudaFreeHost(ptr);
        ptr = nullptr;
        cap = 0;
        const size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMallocHost(&ptr, want);
        if (e == cudaSuccess) cap = want;
        ++g_alloc_gen;  // a captured copy node may hold the old address
        return e;
    }
    void release() {
        if (ptr) cudaFreeHost(ptr);
        ptr = nullptr;
        cap = 0;
    }
};

}  // namespace

struct ptb_model {
    ptb_config cfg{};
    int sm_count = 148;
    int nz = 0;
    double dk = 0, dg = 0;
    std::vector<double> ze, zm, mu, ks, gs;
    DevBuf d_tab;  // ze | mu | gs | ks | ldmu200 | ldz200
    double *d_ze = nullptr, *d_mu = nullptr, *d_gs = nullptr, *d_ks = nullptr, *d_ldmu = nullptr, *d_ldz = nullptr;
    DevBuf d_W;

    // dataset (set_data)
    bool has_data = false;
    int64_t npt = 0, nlc = 0, npb = 0, nep = 0;
    const double *d_time = nullptr;  // borrowed device pointer or d_time_own
    DevBuf d_time_own, d_meta;       // meta: lcids32[npt] | pbids | epids | nsamples | exptimes
    int32_t *d_lcids = nullptr, *d_pbids = nullptr, *d_epids = nullptr, *d_nsamples = nullptr;
    double *d_exptimes = nullptr;
    int ns_max = 1;
    int64_t nblk64 = 0;                // 64-point classification blocks
    DevBuf d_bmeta;                    // bmin[nblk64] | bmax[nblk64] | blc[nblk64]
    double *d_bmin = nullptr, *d_bmax = nullptr;
    int32_t *d_blc = nullptr;
    // the same at 16-point granularity ("cells"): the supersampled kernel classifies finer (a Kepler long-cadence block
    // of 64 points spans 1.3 d, more than a third of a hot Jupiter's period); built only when some nsamples > 1
    int64_t ncell = 0;
    DevBuf d_cmeta, d_cobs;            // cmin | cmax | clc ; cchi | cnoise
    double *d_cmin = nullptr, *d_cmax = nullptr, *d_cchi = nullptr;
    int32_t *d_clc = nullptr, *d_cnoise = nullptr;
    std::vector<int32_t> h_lcids;      // host copy (empty when nlc == 1)
    std::vector<int64_t> h_nsamples;
    std::vector<double> h_exptimes;

    // observations (set_obs)
    bool has_obs = false;
    int64_t nblocks = 0;
    const double *d_obs = nullptr;
    DevBuf d_obs_own, d_blk, d_nblk, d_bobs;  // d_bobs: bchi[nblk64] | bnoise[nblk64]
    double *d_bchi = nullptr;
    int32_t *d_bnoise = nullptr;
    bool blk_trivial = true;

    // per-vector workspaces
    DevBuf d_orb, d_ldrec, d_ldp, d_istar, d_flux, d_partial, d_isig2, d_lnl, d_xyc;
    DevBuf d_tsw, d_tsrec, d_sort;
    DevBuf d_tsgeo;                  // TSModel geometry pass [npv][npt]
    DevBuf d_lpf;                    // mapped LPF parameters (k_lpf_map)
    DevBuf d_basis, d_blmeta;        // LPF baseline: basis[nbasis][npt]; cstart[nlc] | ncoef[nlc] (int32)
    int64_t bl_nbasis = 0;           // 0: no baseline registered
    DevBuf d_cells;                  // per-vector table cells of the tabulated-profile interpolation
    DevBuf d_dummy;                  // zero limb-darkening coefficients of the eclipse model (uniform disk)
    bool ecl_mode = false;           // the next launch_rr_setup expands about mid-eclipse (ptb_eclipse_evaluate)
    double ecl_rstar = 1.0;
    const double *ecl_rstar_v = nullptr;  // per-vector stellar radii (device) for ptb_es_evaluate
    DevBuf d_rec, d_work;            // RoadRunner per-vector records; work counters of the persistent kernel
    bool work_dirty = false;         // a CUDA error was seen since the counters were last known to be re-armed
    int recstride = 0, rec_ld = 0;   // record stride / offset of the ld rows (doubles) of the last setup
    cudaStream_t side_stream = nullptr;  // the orbit solve runs here, concurrently with the table contraction
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    size_t ldm_smem = 0;
    bool keep_stages = true;         // write ldp / istar taps (ptb_get_stage); off in throughput runs
    int pt_occ[32] = {};  // resident CTAs per SM of each k_rr_points instantiation (0 = not queried)
    size_t pt_occ_smem[32] = {};
    bool xyc_injected = false;
    int64_t xyc_npv = 0;
    int64_t last_npv = 0, last_npb = 0, last_flux_count = 0;

    // staging of host inputs: one pinned block + one device block per call
    PinBuf h_stage;
    DevBuf d_stage;
    cudaEvent_t stage_ev = nullptr;  // completion of the last copy out of h_stage
    bool stage_pending = false;

    // optional per-kernel timing (ptb_set_profiling)
    bool profiling = false;
    static constexpr int TRING = 256;
    std::vector<cudaEvent_t> tev;  // TRING x 4 events
    int64_t tcalls = 0;            // timed calls since profiling was enabled

    // managed host results (ptb_bind_host_result): delta transfer of the flux into caller-owned pinned memory.  Several
    // buffers can be bound at once (a caller that keeps the previous result alive alternates between two or three);
    // each carries its own lit-block bitmap describing what IT currently holds.
    struct HostBinding {
        void *buf = nullptr, *dev = nullptr;  // host address / its device mapping
        int64_t count = 0;
        size_t esize = 0;
        bool valid = false;              // buf holds a complete previous result and d_lit its lit-block bitmap
        DevBuf d_lit;                    // lit bitmap words | 8-byte counter of blocks written
    };
    std::vector<HostBinding *> hr;
    PinBuf h_hrstat;                 // the counter, read back with the result
    std::vector<cudaEvent_t> ev_part;  // pipelined host delivery: points kernel of part i done / all deltas done
    int64_t hr_last_bytes = 0, hr_delta_calls = 0, hr_full_calls = 0;

    // CUDA-graph replay of launch-bound calls (small populations): the per-call kernel sequence is captured once
    // per argument signature and relaunched with ONE API call; host arguments travel through the same pinned block
    struct GraphEntry {
        std::vector<unsigned long long> key;
        cudaGraphExec_t exec = nullptr;
        int64_t launches = 0;
        int nchunks = 1;
        unsigned long long used = 0;
    };
    std::vector<GraphEntry> graphs;
    std::vector<std::vector<unsigned long long>> seen_keys;  // signatures that have run once (buffers are sized)
    cudaStream_t cap_stream = nullptr;
    bool capturing = false;
    bool graphs_enabled = true;
    unsigned long long data_gen = 1, graph_clock = 0;
    int64_t graph_replays = 0, graph_captures = 0;

    std::string err;
    int64_t launches = 0;
};

namespace {

int fail(ptb_model *h, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (h) h->err = buf;
    else g_create_error = buf;
    return code;
}

#define CU(call)                                                                                      \
    do {                                                                                              \
        cudaError_t _e = (call);                                                                      \
        if (_e != cudaSuccess) {                                                                      \
            if (h) h->work_dirty = true; /* a kernel may have died before re-arming its counters */   \
            return fail(h, _e == cudaErrorMemoryAllocation ? PTB_ENOMEM : PTB_ECUDA, "%s failed: %s", \
                        #call, cudaGetErrorString(_e));                                               \
        }                                                                                             \
    } while (0)

// A call classifies each argument pointer once: the graph key and the stager ask about the same pointers.
struct PtrMemo {
    const void *p[16];
    bool dev[16];
    int n = 0;
};
thread_local PtrMemo g_ptr_memo;

bool is_device_ptr(const void *p) {
    if (!p) return false;
    PtrMemo &m = g_ptr_memo;
    for (int i = 0; i < m.n; ++i)
        if (m.p[i] == p) return m.dev[i];
    cudaPointerAttributes at;
    bool dev = false;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) cudaGetLastError();
    else dev = at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
    if (m.n < 16) {
        m.p[m.n] = p;
        m.dev[m.n] = dev;
        m.n++;
    }
    return dev;
}

// Stager: host arrays are packed into the pinned block and shipped with ONE async copy;
// device arrays are used in place.
struct Stager {
    ptb_model *h;
    cudaStream_t st;
    struct Item { const void *src; size_t off, bytes; };
    std::vector<Item> items;
    size_t total = 0;
    explicit Stager(ptb_model *h_, cudaStream_t s) : h(h_), st(s) {}
    // returns a token: device pointers pass through, host ones are resolved after commit()
    struct Ref { const void *dev; size_t off; bool staged; };
    Ref add(const void *p, size_t bytes) {
        if (!p) return {nullptr, 0, false};
        if (is_device_ptr(p)) return {p, 0, false};
        const size_t off = total;
        items.push_back({p, off, bytes});
        total += (bytes + 255) & ~size_t(255);
        return {nullptr, off, true};
    }
    // pack the host arrays into the pinned block (after the previous copy out of it has finished)
    int pack() {
        if (total == 0) return PTB_OK;
        if (h->stage_pending) {  // the previous call's copy must have left the pinned block
            CU(cudaEventSynchronize(h->stage_ev));
            h->stage_pending = false;
        }
        CU(h->h_stage.reserve(total));
        CU(h->d_stage.reserve(total));
        for (auto &it : items) memcpy(static_cast<char *>(h->h_stage.ptr) + it.off, it.src, it.bytes);
        return PTB_OK;
    }
    // mark the pinned block busy until the work just queued on `s` has consumed it
    int fence(cudaStream_t s) {
        if (total == 0) return PTB_OK;
        if (!h->stage_ev) CU(cudaEventCreateWithFlags(&h->stage_ev, cudaEventDisableTiming));
        CU(cudaEventRecord(h->stage_ev, s));
        h->stage_pending = true;
        return PTB_OK;
    }
    int commit() {
        if (total == 0) return PTB_OK;
        if (int rc = pack()) return rc;
        CU(cudaMemcpyAsync(h->d_stage.ptr, h->h_stage.ptr, total, cudaMemcpyHostToDevice, st));
        if (h->capturing) return PTB_OK;  // the event is recorded on the launching stream after the graph launch
        return fence(st);
    }
    template <class T>
    const T *get(const Ref &r) const {
        if (!r.staged) return static_cast<const T *>(r.dev);
        return reinterpret_cast<const T *>(static_cast<const char *>(h->d_stage.ptr) + r.off);
    }
};

int set_device(ptb_model *h) {
    g_ptr_memo.n = 0;  // every entry point starts here: pointer classifications do not outlive a call
    CU(cudaSetDevice(h->cfg.device));
    return PTB_OK;
}

// timing marks: 0/1 bracket the per-vector setup, 2/3 the dominant kernel
int mark(ptb_model *h, int i, cudaStream_t st) {
    if (!h->profiling) return PTB_OK;
    CU(cudaEventRecord(h->tev[(h->tcalls % ptb_model::TRING) * 4 + i], st));
    if (i == 3) h->tcalls++;
    return PTB_OK;
}

// create_z_grid (common.py:131-149)
void make_z_grid(double zcut, int nin, int nedge, std::vector<double> &ze, std::vector<double> &zm) {
    const int n = nin + nedge;
    ze.assign(n, 0.0);
    zm.assign(n, 0.0);
    const double mucut = std::sqrt(1.0 - zcut * zcut);
    const double dz = zcut / nin, dmu = mucut / nedge;
    for (int i = 0; i < nin - 1; ++i) ze[i] = (i + 1) * dz;
    for (int i = 0; i <= nedge; ++i) {
        const double v = i * dmu;
        ze[n - 1 - i] = std::sqrt(1 - v * v);
    }
    for (int i = 0; i + 1 < n; ++i) zm[i + 1] = 0.5 * (ze[i] + ze[i + 1]);
}

// numpy/numba linspace: start + i*step, last element = stop
std::vector<double> linspace(double a, double b, int n) {
    std::vector<double> v(n);
    if (n > 1) {
        const double step = (b - a) / (n - 1);
        for (int i = 0; i < n; ++i) v[i] = a + i * step;
        v[n - 1] = b;
    } else if (n == 1) v[0] = a;
    return v;
}

ptb_model::HostBinding *find_binding(ptb_model *h, const void *host, size_t count) {
    if (!host) return nullptr;
    for (auto *b : h->hr)
        if (b->buf == host && (size_t)b->count == count) return b;
    return nullptr;
}

// Device result -> host.  Plain buffers get one copy-engine transfer.  The bound managed buffer
// (ptb_bind_host_result) gets a full transfer the first time and k_host_delta afterwards: only blocks
// that differ from 1.0 now, or did after the previous call, cross PCIe.  Returns after the data has landed.
int deliver_host(ptb_model *h, void *host, const void *dsrc, size_t count, size_t esize, cudaStream_t st) {
    ptb_model::HostBinding *B = find_binding(h, host, count);
    if (!B) {
        CU(cudaMemcpyAsync(host, dsrc, count * esize, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        return PTB_OK;
    }
    const long long nwords = ((long long)count + 32 * HD_BLOCK - 1) / (32 * HD_BLOCK);
    const bool delta = B->valid && B->esize == esize;
    B->valid = false;  // until this transfer has completed
    const size_t lit_bytes = ((size_t)nwords * 4 + 15) & ~size_t(15);
    if (lit_bytes + 16 > B->d_lit.cap && delta) return fail(h, PTB_ESTATE, "host result: bitmap lost");
    CU(B->d_lit.reserve(lit_bytes + 16));
    CU(h->h_hrstat.reserve(16));
    unsigned *lit = B->d_lit.as<unsigned>();
    unsigned long long *nwr = reinterpret_cast<unsigned long long *>(static_cast<char *>(B->d_lit.ptr) + lit_bytes);
    CU(cudaMemsetAsync(nwr, 0, 8, st));
    const unsigned grid = (unsigned)std::min<long long>((nwords + 7) / 8, (long long)h->sm_count * 8);
    if (esize == 8)
        k_host_delta<double><<<grid, 256, 0, st>>>(static_cast<const double *>(dsrc), static_cast<double *>(B->dev), lit, nwr,
                                                   (long long)count, 0, nwords, delta ? 0 : 1);
    else
        k_host_delta<float><<<grid, 256, 0, st>>>(static_cast<const float *>(dsrc), static_cast<float *>(B->dev), lit, nwr,
                                                  (long long)count, 0, nwords, delta ? 0 : 1);
    h->launches++;
    CU(cudaGetLastError());
    if (!delta) CU(cudaMemcpyAsync(host, dsrc, count * esize, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(h->h_hrstat.ptr, nwr, 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    B->esize = esize;
    B->valid = true;
    if (delta) {
        h->hr_last_bytes = (int64_t)(*static_cast<unsigned long long *>(h->h_hrstat.ptr)) * HD_BLOCK * (int64_t)esize;
        h->hr_delta_calls++;
    } else {
        h->hr_last_bytes = (int64_t)(count * esize);
        h->hr_full_calls++;
    }
    return PTB_OK;
}

struct ModelArgs {
    int64_t npv, kcols, nld;
    const double *k, *ld, *istar, *t0, *p, *a, *inc, *e, *w;
};

}  // namespace

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" {

void ptb_default_config(ptb_config *cfg) {
    if (!cfg) return;
    memset(cfg, 0, sizeof *cfg);
    cfg->device = 0;
    cfg->ldlaw = PTB_LD_QUADRATIC;
    cfg->nk = 256;
    cfg->nzin = 20;
    cfg->nzlimb = 20;
    cfg->ng = 100;
    cfg->kmin = 0.005;
    cfg->kmax = 0.5;
    cfg->zcut = 0.7;
    cfg->precompute_weights = 0;
    cfg->precision = 0;
}

int ptb_version(void) { return PTB_VERSION; }

const char *ptb_last_error(const ptb_model *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int ptb_create(const ptb_config *cfg, ptb_model **out) {
    ptb_model *h = nullptr;
    if (!cfg || !out) return fail(h, PTB_EINVAL, "ptb_create: null argument");
    *out = nullptr;
    const bool law_ok = (cfg->ldlaw >= 0 && cfg->ldlaw <= PTB_LD_POWER_2_PM) || cfg->ldlaw == PTB_LD_PROFILES;
    if (!law_ok) return fail(h, PTB_ENOTIMPL, "unknown limb darkening law id %d", cfg->ldlaw);
    if (cfg->nk < 2 || cfg->ng < 2 || cfg->nzin < 2 || cfg->nzlimb < 1 || !(cfg->kmax > cfg->kmin) || !(cfg->kmin > 0) ||
        !(cfg->zcut > 0 && cfg->zcut < 1))
        return fail(h, PTB_EINVAL, "invalid integration grid (nk=%d ng=%d nzin=%d nzlimb=%d klims=(%g,%g) zcut=%g)",
                    cfg->nk, cfg->ng, cfg->nzin, cfg->nzlimb, cfg->kmin, cfg->kmax, cfg->zcut);
    if (cfg->precision != 0 && cfg->precision != 1) return fail(h, PTB_EINVAL, "precision=%d: 0 (fp64) or 1 (opt-in fp32) expected", cfg->precision);
    const int nz = cfg->nzin + cfg->nzlimb;
    if (((size_t)cfg->ng * nz) % 2 != 0 || (size_t)cfg->ng * nz * 16 > 200 * 1024)
        return fail(h, PTB_EINVAL, "ng*nz = %d*%d: two table rows must fit 200 KB of shared memory and be 16-byte multiples",
                    cfg->ng, nz);
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(h, PTB_ECUDA, "no CUDA device available (%s); libptb200 has no CPU fallback",
                    ce != cudaSuccess ? cudaGetErrorString(ce) : "device count 0");
    }
    if (cfg->device < 0 || cfg->device >= ndev) return fail(h, PTB_EINVAL, "device %d out of range [0,%d)", cfg->device, ndev);

    h = new ptb_model();
    h->cfg = *cfg;
    h->nz = nz;
    auto bail = [&](int code) {
        g_create_error = h->err;
        ptb_destroy(h);
        return code;
    };
    if (set_device(h) != PTB_OK) return bail(PTB_ECUDA);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, cfg->device) == cudaSuccess) h->sm_count = prop.multiProcessorCount;
    if (prop.major < 10) {
        fail(h, PTB_ECUDA, "device %d is sm_%d%d; libptb200 is built for sm_100a only", cfg->device, prop.major, prop.minor);
        return bail(PTB_ECUDA);
    }

    make_z_grid(cfg->zcut, cfg->nzin, cfg->nzlimb, h->ze, h->zm);
    h->mu.resize(nz);
    for (int i = 0; i < nz; ++i) h->mu[i] = std::sqrt(1 - h->zm[i] * h->zm[i]);
    h->ks = linspace(cfg->kmin, cfg->kmax, cfg->nk);
    h->gs = linspace(0.0, 1.0 - 1e-7, cfg->ng);
    h->dk = (cfg->kmax - cfg->kmin) / cfg->nk;  // NOT the linspace step (common.py:223)
    h->dg = h->gs[1] - h->gs[0];
    std::vector<double> ldmu = linspace(1.0, 0.0, 200), ldz(200);
    for (int i = 0; i < 200; ++i) ldz[i] = std::sqrt(1 - ldmu[i] * ldmu[i]);

    std::vector<double> tab;
    auto push = [&](const std::vector<double> &v) {
        size_t off = tab.size();
        tab.insert(tab.end(), v.begin(), v.end());
        if (tab.size() % 2) tab.push_back(0.0);
        return off;
    };
    const size_t o_ze = push(h->ze), o_mu = push(h->mu), o_gs = push(h->gs), o_ks = push(h->ks), o_lm = push(ldmu),
                 o_lz = push(ldz);
    auto cu = [&](cudaError_t e, const char *what) {
        if (e == cudaSuccess) return true;
        fail(h, e == cudaErrorMemoryAllocation ? PTB_ENOMEM : PTB_ECUDA, "%s failed: %s", what, cudaGetErrorString(e));
        return false;
    };
    if (!cu(h->d_tab.reserve(tab.size() * 8), "cudaMalloc(tables)")) return bail(PTB_ECUDA);
    if (!cu(cudaMemcpy(h->d_tab.ptr, tab.data(), tab.size() * 8, cudaMemcpyHostToDevice), "cudaMemcpy(tables)"))
        return bail(PTB_ECUDA);
    double *base = h->d_tab.as<double>();
    h->d_ze = base + o_ze;
    h->d_mu = base + o_mu;
    h->d_gs = base + o_gs;
    h->d_ks = base + o_ks;
    h->d_ldmu = base + o_lm;
    h->d_ldz = base + o_lz;

    const size_t wcount = (size_t)cfg->nk * cfg->ng * nz;
    if (!cu(h->d_W.reserve(wcount * 8), "cudaMalloc(weights)")) return bail(PTB_ENOMEM);
    const int rows = cfg->nk * cfg->ng;
    k_weight_table<<<(rows + 127) / 128, 128>>>(h->d_ks, h->d_gs, h->d_ze, cfg->nk, cfg->ng, nz, h->d_W.as<double>());
    h->launches++;
    if (!cu(cudaGetLastError(), "k_weight_table launch")) return bail(PTB_ECUDA);
    if (!cu(cudaDeviceSynchronize(), "k_weight_table")) return bail(PTB_ECUDA);

    *out = h;
    return PTB_OK;
}

void ptb_destroy(ptb_model *h) {
    if (!h) return;
    cudaSetDevice(h->cfg.device);
    for (DevBuf *b : {&h->d_tab, &h->d_W, &h->d_time_own, &h->d_meta, &h->d_obs_own, &h->d_blk, &h->d_nblk, &h->d_orb,
                      &h->d_ldrec, &h->d_ldp, &h->d_istar, &h->d_flux, &h->d_partial, &h->d_isig2, &h->d_lnl, &h->d_xyc,
                      &h->d_tsw, &h->d_tsrec, &h->d_stage, &h->d_bmeta, &h->d_bobs, &h->d_cmeta, &h->d_cobs, &h->d_sort, &h->d_rec, &h->d_work, &h->d_tsgeo, &h->d_lpf, &h->d_basis, &h->d_blmeta, &h->d_cells, &h->d_dummy})
        b->release();
    for (auto *b : h->hr) { b->d_lit.release(); delete b; }
    h->hr.clear();
    h->h_stage.release();
    h->h_hrstat.release();
    for (auto &ev : h->ev_part) if (ev) cudaEventDestroy(ev);
    for (auto &g : h->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
    h->graphs.clear();
    if (h->cap_stream) cudaStreamDestroy(h->cap_stream);
    if (h->stage_ev) cudaEventDestroy(h->stage_ev);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    if (h->side_stream) cudaStreamDestroy(h->side_stream);
    for (auto &e : h->tev) if (e) cudaEventDestroy(e);
    h->tev.clear();
    delete h;
}

int ptb_get_tables(ptb_model *h, double *ze, double *zm, double *mu, double *weights, double *dk, double *dg) {
    if (!h) return PTB_EINVAL;
    if (int rc = set_device(h)) return rc;
    if (ze) memcpy(ze, h->ze.data(), h->nz * 8);
    if (zm) memcpy(zm, h->zm.data(), h->nz * 8);
    if (mu) memcpy(mu, h->mu.data(), h->nz * 8);
    if (dk) *dk = h->dk;
    if (dg) *dg = h->dg;
    if (weights) {
        const size_t n = (size_t)h->cfg.nk * h->cfg.ng * h->nz;
        CU(cudaMemcpy(weights, h->d_W.ptr, n * 8, is_device_ptr(weights) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost));
    }
    return PTB_OK;
}

int ptb_set_data(ptb_model *h, const double *time, int64_t npt, const int64_t *lcids, int64_t nlc, const int64_t *pbids,
                 int64_t npb, const int64_t *epids, int64_t nep, const int64_t *nsamples, const double *exptimes) {
    if (!h) return PTB_EINVAL;
    if (int rc = set_device(h)) return rc;
    if (!time || npt <= 0) return fail(h, PTB_EINVAL, "set_data: time array is empty");
    if (npt > 0x7fffffffLL) return fail(h, PTB_EINVAL, "set_data: npt=%lld exceeds the 2^31-1 points per dataset limit", (long long)npt);
    if (nlc < 1 || npb < 1 || nep < 1) return fail(h, PTB_ESHAPE, "set_data: nlc, npb, nep must be >= 1");
    if (!lcids && nlc != 1) return fail(h, PTB_ESHAPE, "set_data: lcids missing but nlc=%lld", (long long)nlc);
    if (npb > 256) return fail(h, PTB_EINVAL, "set_data: npb=%lld > 256 passbands (use the TSModel entry point for spectroscopy)", (long long)npb);

    // host views of the small per-light-curve arrays (they may live on the device)
    auto fetch64 = [&](const int64_t *p, int64_t n, std::vector<int64_t> &v, int64_t def) -> int {
        v.assign(n, def);
        if (!p) return PTB_OK;
        CU(cudaMemcpy(v.data(), p, n * 8, cudaMemcpyDefault));
        return PTB_OK;
    };
    std::vector<int64_t> hpb, hep, hns;
    if (int rc = fetch64(pbids, nlc, hpb, 0)) return rc;
    if (int rc = fetch64(epids, nlc, hep, 0)) return rc;
    if (int rc = fetch64(nsamples, nlc, hns, 1)) return rc;
    std::vector<double> het(nlc, 0.0);
    if (exptimes) CU(cudaMemcpy(het.data(), exptimes, nlc * 8, cudaMemcpyDefault));
    int nsmax = 1;
    for (int64_t i = 0; i < nlc; ++i) {
        if (hpb[i] < 0 || hpb[i] >= npb)
            return fail(h, PTB_EINVAL, "Passband indices (`pbids`) for %lld unique passbands should be given as integers between 0 and %lld.",
                        (long long)npb, (long long)npb - 1);
        if (hep[i] < 0 || hep[i] >= nep) return fail(h, PTB_EINVAL, "set_data: epids[%lld]=%lld outside [0,%lld)", (long long)i, (long long)hep[i], (long long)nep);
        if (hns[i] < 1 || hns[i] > 100000) return fail(h, PTB_EINVAL, "set_data: nsamples[%lld]=%lld must be in [1,100000]", (long long)i, (long long)hns[i]);
        nsmax = std::max<int>(nsmax, (int)hns[i]);
    }

    // light-curve ids -> int32 on the device (validated against [0,nlc))
    std::vector<int32_t> meta;
    const size_t n_lc32 = lcids ? (size_t)npt : 0;
    const size_t lc_pad = (n_lc32 + 3) & ~size_t(3);
    meta.resize(lc_pad + 3 * ((nlc + 3) & ~int64_t(3)) + 2 * nlc + 8, 0);
    if (lcids) {
        std::vector<int64_t> hl(npt);
        CU(cudaMemcpy(hl.data(), lcids, npt * 8, cudaMemcpyDefault));
        for (int64_t i = 0; i < npt; ++i) {
            if (hl[i] < 0 || hl[i] >= nlc)
                return fail(h, PTB_EINVAL, "set_data: lcids[%lld]=%lld outside [0,%lld)", (long long)i, (long long)hl[i], (long long)nlc);
            meta[i] = (int32_t)hl[i];
        }
    }
    const size_t nlc4 = (nlc + 3) & ~int64_t(3);
    const size_t o_pb = lc_pad, o_ep = o_pb + nlc4, o_ns = o_ep + nlc4, o_et = o_ns + nlc4;  // o_et: 16-byte aligned
    for (int64_t i = 0; i < nlc; ++i) {
        meta[o_pb + i] = (int32_t)hpb[i];
        meta[o_ep + i] = (int32_t)hep[i];
        meta[o_ns + i] = (int32_t)hns[i];
    }
    memcpy(&meta[o_et], het.data(), nlc * 8);
    CU(h->d_meta.reserve(meta.size() * 4));
    CU(cudaMemcpy(h->d_meta.ptr, meta.data(), meta.size() * 4, cudaMemcpyHostToDevice));
    int32_t *mb = h->d_meta.as<int32_t>();
    h->d_lcids = lcids ? mb : nullptr;
    h->d_pbids = mb + o_pb;
    h->d_epids = mb + o_ep;
    h->d_nsamples = mb + o_ns;
    h->d_exptimes = reinterpret_cast<double *>(mb + o_et);

    std::vector<double> htime;
    const double *ht = time;
    if (is_device_ptr(time)) {
        h->d_time = time;  // zero copy: the caller keeps the tensor alive (as with RoadRunnerModelCL buffers)
        htime.resize(npt);
        CU(cudaMemcpy(htime.data(), time, npt * 8, cudaMemcpyDeviceToHost));
        ht = htime.data();
    } else {
        CU(h->d_time_own.reserve(npt * 8 + 16));
        CU(cudaMemcpy(h->d_time_own.ptr, time, npt * 8, cudaMemcpyHostToDevice));
        h->d_time = h->d_time_own.as<double>();
    }
    // per block of `bs` points: time range and light curve (block classification of the points kernels)
    auto build_blocks = [&](int64_t bs, DevBuf &buf, double *&dmin, double *&dmax, int32_t *&dlc, int64_t &nb_out) -> int {
        const int64_t nb = (npt + bs - 1) / bs;
        std::vector<double> bm(2 * nb);
        std::vector<int32_t> bl(nb, 0);
        for (int64_t b = 0; b < nb; ++b) {
            const int64_t j0 = b * bs, j1 = std::min<int64_t>(npt, j0 + bs);
            double mn = ht[j0], mx = ht[j0];
            bool nanv = false;
            int32_t lc = lcids ? meta[j0] : 0;
            for (int64_t j = j0; j < j1; ++j) {
                nanv |= std::isnan(ht[j]);
                mn = std::min(mn, ht[j]);
                mx = std::max(mx, ht[j]);
                if (lcids && meta[j] != lc) lc = -1;
            }
            if (nanv) lc = -1;  // force the exact per-point path
            bm[b] = mn;
            bm[nb + b] = mx;
            bl[b] = lc;
        }
        CU(buf.reserve(nb * 20 + 64));
        CU(cudaMemcpy(buf.ptr, bm.data(), nb * 16, cudaMemcpyHostToDevice));
        dmin = buf.as<double>();
        dmax = dmin + nb;
        dlc = reinterpret_cast<int32_t *>(dmax + nb);
        CU(cudaMemcpy(dlc, bl.data(), nb * 4, cudaMemcpyHostToDevice));
        nb_out = nb;
        return PTB_OK;
    };
    if (int rc = build_blocks(PT_BLOCK, h->d_bmeta, h->d_bmin, h->d_bmax, h->d_blc, h->nblk64)) return rc;
    h->ncell = 0;
    if (nsmax > 1)
        if (int rc = build_blocks(SS_CELL, h->d_cmeta, h->d_cmin, h->d_cmax, h->d_clc, h->ncell)) return rc;
    h->npt = npt;
    h->nlc = nlc;
    h->npb = npb;
    h->nep = nep;
    h->ns_max = nsmax;
    h->h_nsamples = hns;
    h->h_exptimes = het;
    h->has_data = true;
    h->has_obs = false;  // observations and the baseline basis are tied to the time axis
    h->bl_nbasis = 0;
    h->data_gen++;
    return PTB_OK;
}

int ptb_set_obs(ptb_model *h, const double *obs, const int64_t *slices, const int64_t *nids, int64_t nsl, int64_t nblocks) {
    if (!h) return PTB_EINVAL;
    if (int rc = set_device(h)) return rc;
    if (!h->has_data) return fail(h, PTB_ESTATE, "set_obs: call set_data first");
    if (!obs) return fail(h, PTB_EINVAL, "set_obs: obs is null");
    const int64_t npt = h->npt;
    std::vector<int32_t> blk(npt, -1);
    std::vector<double> cnt;
    bool trivial = false;
    if (!slices) {
        nblocks = 1;
        cnt.assign(1, (double)npt);
        trivial = true;
    } else {
        if (nsl < 1 || nblocks < 1 || !nids) return fail(h, PTB_ESHAPE, "set_obs: slices given but nsl/nblocks/nids invalid");
        std::vector<int64_t> hs(2 * nsl), hn(nsl);
        CU(cudaMemcpy(hs.data(), slices, 16 * nsl, cudaMemcpyDefault));
        CU(cudaMemcpy(hn.data(), nids, 8 * nsl, cudaMemcpyDefault));
        cnt.assign(nblocks, 0.0);
        for (int64_t s = 0; s < nsl; ++s) {
            if (hn[s] < 0 || hn[s] >= nblocks) return fail(h, PTB_EINVAL, "set_obs: nids[%lld]=%lld outside [0,%lld)", (long long)s, (long long)hn[s], (long long)nblocks);
            if (hs[2 * s] < 0 || hs[2 * s + 1] > npt || hs[2 * s] > hs[2 * s + 1])
                return fail(h, PTB_EINVAL, "set_obs: slice %lld = [%lld,%lld) outside [0,%lld]", (long long)s, (long long)hs[2 * s], (long long)hs[2 * s + 1], (long long)npt);
            for (int64_t j = hs[2 * s]; j < hs[2 * s + 1]; ++j) {
                if (blk[j] >= 0) return fail(h, PTB_EINVAL, "set_obs: overlapping slices at point %lld are not supported", (long long)j);
                blk[j] = (int32_t)hn[s];
            }
            cnt[hn[s]] += (double)(hs[2 * s + 1] - hs[2 * s]);
        }
        trivial = (nblocks == 1 && cnt[0] == (double)npt);
    }
    if (!trivial) {
        CU(h->d_blk.reserve(npt * 4));
        CU(cudaMemcpy(h->d_blk.ptr, blk.data(), npt * 4, cudaMemcpyHostToDevice));
    }
    CU(h->d_nblk.reserve(nblocks * 8));
    CU(cudaMemcpy(h->d_nblk.ptr, cnt.data(), nblocks * 8, cudaMemcpyHostToDevice));
    std::vector<double> hobs;
    const double *ho = obs;
    if (is_device_ptr(obs)) {
        h->d_obs = obs;
        hobs.resize(npt);
        CU(cudaMemcpy(hobs.data(), obs, npt * 8, cudaMemcpyDeviceToHost));
        ho = hobs.data();
    } else {
        CU(h->d_obs_own.reserve(npt * 8 + 16));
        CU(cudaMemcpy(h->d_obs_own.ptr, obs, npt * 8, cudaMemcpyHostToDevice));
        h->d_obs = h->d_obs_own.as<double>();
    }
    // per block of `bs` points: sum of (obs-1)^2 and the block's noise id (likelihood fast path of the points kernels)
    auto build_obs_blocks = [&](int64_t bs, int64_t nb, DevBuf &buf, double *&dchi, int32_t *&dnoise) -> int {
        std::vector<double> bc(nb, 0.0);
        std::vector<int32_t> bn(nb, -1);
        for (int64_t b = 0; b < nb; ++b) {
            const int64_t j0 = b * bs, j1 = std::min<int64_t>(npt, j0 + bs);
            int32_t id = -1;
            bool mixed = false;
            double sum = 0.0;
            for (int64_t j = j0; j < j1; ++j) {
                const int32_t bj = trivial ? 0 : blk[j];
                if (bj < 0) continue;
                if (id < 0) id = bj;
                else if (id != bj) mixed = true;
                const double d = ho[j] - 1.0;
                sum += d * d;
            }
            bc[b] = sum;
            bn[b] = mixed ? -2 : id;
        }
        CU(buf.reserve(nb * 12 + 64));
        dchi = buf.as<double>();
        dnoise = reinterpret_cast<int32_t *>(dchi + nb);
        CU(cudaMemcpy(dchi, bc.data(), nb * 8, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(dnoise, bn.data(), nb * 4, cudaMemcpyHostToDevice));
        return PTB_OK;
    };
    if (int rc = build_obs_blocks(PT_BLOCK, h->nblk64, h->d_bobs, h->d_bchi, h->d_bnoise)) return rc;
    if (h->ncell > 0)
        if (int rc = build_obs_blocks(SS_CELL, h->ncell, h->d_cobs, h->d_cchi, h->d_cnoise)) return rc;
    h->blk_trivial = trivial;
    h->nblocks = nblocks;
    h->has_obs = true;
    h->data_gen++;
    return PTB_OK;
}

}  // extern "C"

namespace {

int check_model_args(ptb_model *h, const char *who, const ModelArgs &A, int64_t npb) {
    if (!h->has_data) return fail(h, PTB_ESTATE, "%s: call set_data first", who);
    if (A.npv < 1) return fail(h, PTB_ESHAPE, "%s: npv must be >= 1", who);
    if (A.npv > 0x7fffffffLL / 64) return fail(h, PTB_EINVAL, "%s: npv=%lld too large for one call", who, (long long)A.npv);
    if (!A.k || !A.ld || !A.t0 || !A.p || !A.a || !A.inc || !A.e || !A.w) return fail(h, PTB_EINVAL, "%s: null parameter array", who);
    if (A.kcols != 1 && A.kcols != npb)
        return fail(h, PTB_ESHAPE, "Radius ratios should be given either as an [npv, 1] or [npv, npb] array.");
    if (h->cfg.ldlaw == PTB_LD_PROFILES) {
        if (A.nld != h->nz) return fail(h, PTB_ESHAPE, "%s: tabulated profiles need %d mu nodes per passband, got %lld", who, h->nz, (long long)A.nld);
        if (!A.istar) return fail(h, PTB_EINVAL, "%s: istar is required with tabulated profiles", who);
    } else {
        static const int need[] = {0, 1, 2, 2, 4, 1, 2, 2, 2, 2, 2};
        if (A.nld < need[h->cfg.ldlaw]) return fail(h, PTB_ESHAPE, "%s: limb darkening law %d needs %d coefficients, got %lld", who, h->cfg.ldlaw, need[h->cfg.ldlaw], (long long)A.nld);
        if (A.nld < 1) return fail(h, PTB_ESHAPE, "%s: nld must be >= 1", who);
    }
    return PTB_OK;
}


// ---------------------------------------------------------------------------------------------
// CUDA-graph replay.  A call is eligible when the population is small enough to be launch bound and nothing
// is being timed.  The first call with a given signature runs normally (it sizes every buffer), the second
// is captured on a private stream while it is enqueued, and from then on the call is: refill the pinned
// argument block, ONE cudaGraphLaunch into the caller's stream, one event record.
// ---------------------------------------------------------------------------------------------
using GKey = std::vector<unsigned long long>;
constexpr int64_t GRAPH_MAX_NPV = 2048;
constexpr size_t GRAPH_CACHE = 6;

inline unsigned long long pkey(const void *p) {  // device pointers are baked into the graph, host ones are staged
    if (!p) return 0ull;
    return is_device_ptr(p) ? (unsigned long long)reinterpret_cast<uintptr_t>(p) : 1ull;
}

bool graph_eligible(const ptb_model *h, int64_t npv) {
    static const bool env_on = [] {
        const char *e = getenv("PTB_GRAPHS");
        return !(e && atoi(e) == 0);
    }();
    return env_on && h->graphs_enabled && !h->profiling && npv <= GRAPH_MAX_NPV;
}

ptb_model::GraphEntry *graph_find(ptb_model *h, const GKey &key) {
    for (auto &g : h->graphs)
        if (g.key == key) {
            g.used = ++h->graph_clock;
            return &g;
        }
    return nullptr;
}

// true when this signature has already run once; otherwise remembers it (a few recent ones: output tensors of
// the caller typically alternate between two or three addresses)
bool graph_seen(ptb_model *h, const GKey &key) {
    for (auto &k : h->seen_keys)
        if (k == key) return true;
    if (h->seen_keys.size() >= 8) h->seen_keys.erase(h->seen_keys.begin());
    h->seen_keys.push_back(key);
    return false;
}

int graph_fence(ptb_model *h, cudaStream_t st) {
    if (!h->stage_ev) CU(cudaEventCreateWithFlags(&h->stage_ev, cudaEventDisableTiming));
    CU(cudaEventRecord(h->stage_ev, st));
    h->stage_pending = true;
    return PTB_OK;
}

int graph_begin(ptb_model *h) {
    if (!h->cap_stream) CU(cudaStreamCreateWithFlags(&h->cap_stream, cudaStreamNonBlocking));
    if (h->stage_pending) {
        CU(cudaEventSynchronize(h->stage_ev));
        h->stage_pending = false;
    }
    CU(cudaStreamBeginCapture(h->cap_stream, cudaStreamCaptureModeRelaxed));
    h->capturing = true;
    return PTB_OK;
}

// Ends the capture; on success instantiates, stores and launches the graph on `st`.  `*launched` tells the caller
// whether the work has been queued (otherwise it must enqueue normally).
int graph_end(ptb_model *h, const GKey &key, cudaStream_t st, int rc_enqueue, int64_t nlaunches, int nchunks,
              unsigned long long gen0, bool *launched) {
    *launched = false;
    h->capturing = false;
    cudaGraph_t gr = nullptr;
    cudaError_t ce = cudaStreamEndCapture(h->cap_stream, &gr);
    if (ce != cudaSuccess || !gr || rc_enqueue != PTB_OK || gen0 != g_alloc_gen) {
        cudaGetLastError();
        if (gr) cudaGraphDestroy(gr);
        return rc_enqueue;   // nothing queued: the caller falls back to a normal enqueue (or reports the error)
    }
    cudaGraphExec_t exec = nullptr;
    ce = cudaGraphInstantiate(&exec, gr, 0);
    cudaGraphDestroy(gr);
    if (ce != cudaSuccess || !exec) {
        cudaGetLastError();
        return PTB_OK;
    }
    if (h->graphs.size() >= GRAPH_CACHE) {  // evict the least recently used entry
        size_t lru = 0;
        for (size_t i = 1; i < h->graphs.size(); ++i)
            if (h->graphs[i].used < h->graphs[lru].used) lru = i;
        cudaGraphExecDestroy(h->graphs[lru].exec);
        h->graphs.erase(h->graphs.begin() + lru);
    }
    ptb_model::GraphEntry g;
    g.key = key;
    g.exec = exec;
    g.launches = nlaunches;
    g.nchunks = nchunks;
    g.used = ++h->graph_clock;
    h->graphs.push_back(g);
    h->graph_captures++;
    CU(cudaGraphLaunch(exec, st));
    *launched = true;
    return graph_fence(h, st);
}

struct Staged {
    const double *k, *ld, *istar, *t0, *p, *a, *inc, *e, *w, *sigma;
};

int stage_model_args(ptb_model *h, const ModelArgs &A, int64_t npb, int64_t nep, const double *sigma, int64_t nsig,
                     cudaStream_t st, Staged &D, bool pack_only = false) {
    Stager S(h, st);
    const size_t npv = A.npv;
    auto rk = S.add(A.k, npv * A.kcols * 8);
    auto rld = S.add(A.ld, npv * npb * A.nld * 8);
    auto ris = S.add(A.istar, npv * npb * 8);
    auto rt0 = S.add(A.t0, npv * nep * 8);
    auto rp = S.add(A.p, npv * 8), ra = S.add(A.a, npv * 8), ri = S.add(A.inc, npv * 8), re = S.add(A.e, npv * 8),
         rw = S.add(A.w, npv * 8);
    auto rs = S.add(sigma, npv * nsig * 8);
    if (pack_only) return S.pack();  // graph replay: the captured copy node ships the block
    if (int rc = S.commit()) return rc;
    D.k = S.get<double>(rk);
    D.ld = S.get<double>(rld);
    D.istar = S.get<double>(ris);
    D.t0 = S.get<double>(rt0);
    D.p = S.get<double>(rp);
    D.a = S.get<double>(ra);
    D.inc = S.get<double>(ri);
    D.e = S.get<double>(re);
    D.w = S.get<double>(rw);
    D.sigma = S.get<double>(rs);
    return PTB_OK;
}

int launch_rr_setup(ptb_model *h, const ModelArgs &A, const Staged &D, cudaStream_t st) {
    const int64_t npv = A.npv, npb = h->npb;
    const int ng = h->cfg.ng, nz = h->nz, nk = h->cfg.nk;
    const int lds = (ng + 4 + 1) & ~1;
    const int nep_pad = (int)((h->nep + 1) & ~int64_t(1));
    const int rec_ld = ORB_STRIDE + nep_pad;
    const size_t recstride = (size_t)rec_ld + (size_t)npb * lds;
    if (recstride * 8 > 96 * 1024)
        return fail(h, PTB_EINVAL, "per-vector record of %zu bytes (npb=%lld passbands, nep=%lld epochs) exceeds 96 KB", recstride * 8,
                    (long long)npb, (long long)h->nep);
    h->recstride = (int)recstride;
    h->rec_ld = rec_ld;
    CU(h->d_rec.reserve((size_t)npv * recstride * 8));
    CU(h->d_ldp.reserve((size_t)npv * npb * nz * 8));
    CU(h->d_istar.reserve((size_t)npv * npb * 8));
    if (h->xyc_injected && h->xyc_npv != npv)
        return fail(h, PTB_ESHAPE, "injected xyc has npv=%lld but evaluate was called with npv=%lld", (long long)h->xyc_npv, (long long)npv);
    if (nk + 2 > SORT_MAXBINS) return fail(h, PTB_EINVAL, "nk=%d: at most %d table rows are supported", nk, SORT_MAXBINS - 2);

    // ---- orbit solve on the side stream (independent of the table contraction) ------------------------
    if (!h->side_stream) {
        CU(cudaStreamCreateWithFlags(&h->side_stream, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
    }
    CU(cudaEventRecord(h->ev_fork, st));
    CU(cudaStreamWaitEvent(h->side_stream, h->ev_fork, 0));
    OrbitParams O{};
    O.k = D.k; O.p = D.p; O.a = D.a; O.inc = D.inc; O.e = D.e; O.w = D.w;
    O.xyc_in = h->xyc_injected ? h->d_xyc.as<double>() : nullptr;
    O.t0 = D.t0; O.rec = h->d_rec.as<double>();
    O.npv = (int)npv; O.kcols = (int)A.kcols; O.nep = (int)h->nep; O.recstride = (int)recstride;
    O.eclipse = h->ecl_mode ? 1 : 0; O.rstar = h->ecl_rstar; O.rstar_v = h->ecl_mode ? h->ecl_rstar_v : nullptr;
    k_rr_orbit<<<(unsigned)((npv * 8 + 255) / 256), 256, 0, h->side_stream>>>(O);
    CU(cudaGetLastError());
    CU(cudaEventRecord(h->ev_join, h->side_stream));

    // ---- counting sort by table row + group descriptors (one cluster) ---------------------------------------
    // group size: as many vectors as fit ~32 KB of profiles, at most RR_GROUP
    const int grp = (int)std::max<int64_t>(1, std::min<int64_t>(RR_GROUP, (32 * 1024) / (npb * nz * 8)));
    const size_t ngroups_max = (size_t)((npv + grp - 1) / grp + nk + 1);  // every bin can end with a partial group
    // sort workspace: gdesc[ngroups_max] int4 | perm[npv] | ngroups
    CU(h->d_sort.reserve(ngroups_max * 16 + ((size_t)npv + 4) * 4));
    int4 *gdesc = h->d_sort.as<int4>();
    int *perm = reinterpret_cast<int *>(gdesc + ngroups_max);
    int *ngroups = perm + npv;
    SortParams SP{};
    SP.k = D.k; SP.a = D.a; SP.e = D.e; SP.perm = perm; SP.gdesc = gdesc; SP.ngroups = ngroups;
    SP.npv = (int)npv; SP.kcols = (int)A.kcols; SP.nk = nk; SP.grp = grp;
    SP.kmin = h->cfg.kmin; SP.kmax = h->cfg.kmax; SP.dk = h->dk;
    k_bin_sort<<<SORT_CTAS, SORT_THREADS, 0, st>>>(SP);  // one cluster
    h->launches += 2;
    CU(cudaGetLastError());

    LdmParams P{};
    P.k = D.k; P.ld = D.ld; P.istar = D.istar;
    P.W = h->d_W.as<double>(); P.ze = h->d_ze; P.mu = h->d_mu; P.gs = h->d_gs; P.ldmu200 = h->d_ldmu; P.ldz200 = h->d_ldz;
    P.perm = perm; P.ngroups = ngroups; P.gdesc = gdesc;
    P.rec = h->d_rec.as<double>(); P.ldp_out = h->keep_stages ? h->d_ldp.as<double>() : nullptr;
    P.recstride = (int)recstride; P.rec_ld = rec_ld;
    P.istar_out = h->keep_stages ? h->d_istar.as<double>() : nullptr;
    P.npv = (int)npv; P.kcols = (int)A.kcols; P.npb = (int)npb; P.nld = (int)A.nld; P.law = h->cfg.ldlaw;
    P.nk = nk; P.ng = ng; P.nz = nz; P.lds = lds; P.grp = grp;
    P.kmin = h->cfg.kmin; P.dk = h->dk;
    static const bool analytic[] = {true, true, true, true, false, false, false, false, false, true, false};
    P.numeric_istar = (h->cfg.ldlaw != PTB_LD_PROFILES && !analytic[h->cfg.ldlaw]) ? 1 : 0;
    const size_t smem = (2 * (size_t)ng * nz + (size_t)grp * npb * nz + grp * npb + (P.numeric_istar ? 8 * 200 : 0)) * 8 + 16;
    if (smem > 227 * 1024) return fail(h, PTB_EINVAL, "k_rr_ldm needs %zu bytes of shared memory (> 227 KB): too many passbands", smem);
    if (smem != h->ldm_smem) {
        CU(cudaFuncSetAttribute(k_rr_ldm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        h->ldm_smem = smem;
    }
    k_rr_ldm<<<(unsigned)ngroups_max, 256, smem, st>>>(P);
    h->launches++;
    CU(cudaGetLastError());
    CU(cudaStreamWaitEvent(st, h->ev_join, 0));  // records complete: orbit part (side stream) + ld part
    h->last_npv = npv;
    h->last_npb = npb;
    return PTB_OK;
}

template <int VEC, bool SINGLE, bool LNL, bool S1, typename T>
int launch_points_t(ptb_model *h, const PointsParams &P, size_t smem, cudaStream_t st) {
    // one sample per point: k_rr_points; supersampled data sets: k_rr_points_ss (ptb_ss_kernels.cuh)
    void (*kern)(PointsParams);
    if constexpr (S1) kern = k_rr_points<VEC, SINGLE, LNL, T>;
    else kern = k_rr_points_ss<VEC, SINGLE, LNL, T>;
    const int slot = (sizeof(T) == 4 ? 16 : 0) + (S1 ? 8 : 0) + (VEC == 2 ? 4 : 0) + (SINGLE ? 2 : 0) + (LNL ? 1 : 0);
    constexpr int threads = S1 ? PT_THREADS : SS_THREADS, warps = threads / 32;
    if (h->pt_occ[slot] == 0 || h->pt_occ_smem[slot] != smem) {
        CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int occ = 0;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem));
        if (occ < 1) return fail(h, PTB_EINVAL, "k_rr_points does not fit on an SM with %zu bytes of shared memory", smem);
        h->pt_occ[slot] = occ;
        h->pt_occ_smem[slot] = smem;
    }
    // persistent grid: one wave of CTAs; every warp pulls items from the work counter
    const long long nitems = (long long)P.npv * P.nchunks;
    const long long grid = std::min<long long>((nitems + warps - 1) / warps, (long long)h->sm_count * h->pt_occ[slot]);
    kern<<<(unsigned)grid, threads, smem, st>>>(P);
    h->launches++;
    CU(cudaGetLastError());
    return PTB_OK;
}

// flux != nullptr -> flux mode; else lnL mode (partials into h->d_partial)
int launch_points(ptb_model *h, int64_t npv, void *flux, const double *isig2, cudaStream_t st, int *nchunks_out,
                  int64_t row0 = 0) {   // rows [row0, row0 + npv) of the records; `flux` points at row row0
    const int ng = h->cfg.ng;
    const int lds = (ng + 4 + 1) & ~1;
    PointsParams P{};
    P.time = h->d_time; P.lcids = h->d_lcids; P.pbids = h->d_pbids; P.epids = h->d_epids; P.nsamples = h->d_nsamples;
    P.exptimes = h->d_exptimes; P.rec = h->d_rec.as<double>() + (size_t)row0 * h->recstride; P.recstride = h->recstride; P.rec_ld = h->rec_ld;
    P.flux = flux; P.obs = h->d_obs; P.blk = h->blk_trivial ? nullptr : h->d_blk.as<int32_t>(); P.isig2 = isig2;
    P.npt = h->npt; P.npv = (int)npv; P.nlc = (int)h->nlc; P.npb = (int)h->npb; P.nep = (int)h->nep; P.ng = ng; P.lds = lds;
    P.nblocks = (int)h->nblocks; P.ns_max = h->ns_max; P.dg = h->dg; P.inv_dg = 1.0 / h->dg;

    P.bmin = h->d_bmin; P.bmax = h->d_bmax; P.blc = h->d_blc; P.bchi = h->d_bchi; P.bnoise = h->d_bnoise;
    P.nblk64 = (int)h->nblk64;
    P.cmin = h->d_cmin; P.cmax = h->d_cmax; P.clc = h->d_clc; P.cchi = h->d_cchi; P.cnoise = h->d_cnoise; P.ncell = (int)h->ncell;
    if (!h->d_work.ptr || h->work_dirty) {
        // the last CTA of every launch re-arms the counters itself; after a CUDA error (a launch that died
        // mid-way would leave them poisoned) they are reset here before the next launch
        CU(h->d_work.reserve(64));
        CU(cudaMemsetAsync(h->d_work.ptr, 0, 64, st));
        h->work_dirty = false;
    }
    P.work = h->d_work.as<int>();
    const bool single = (h->nlc == 1);
    const bool lnl = (flux == nullptr);
    const bool f32 = h->cfg.precision == 1;
    const int tsize = f32 ? 4 : 8;
    const bool aligned = (h->npt % 2 == 0) && ((reinterpret_cast<uintptr_t>(h->d_time) & 15) == 0) &&
                         (lnl ? (reinterpret_cast<uintptr_t>(h->d_obs) & 15) == 0 : (reinterpret_cast<uintptr_t>(flux) & (f32 ? 7 : 15)) == 0);
    const int vec = aligned ? 2 : 1;
    // Items: rows are cut into chunks of whole 8-block groups so that every warp of the persistent grid
    // gets ~PT_ITEMS_PER_WARP items (a short tail), at least PT_MIN_ITEM_BLOCKS and at most PT_MAXBLK
    // blocks each.
    const long long nb = h->nblk64;
    static const double items_per_warp = [] {
        const char *e = getenv("PTB_ITEMS_PER_WARP");
        return (e && atof(e) > 0) ? atof(e) : 8.0;
    }();
    static const long long min_item_blocks = [] {
        const char *e = getenv("PTB_MIN_ITEM_BLOCKS");
        return (long long)((e && atoi(e) > 0) ? atoi(e) : 16);
    }();
    const long long workers = (long long)h->sm_count * 3 * PT_WARPS;
    // the supersampled kernel's items are ~10x heavier: four times as many keep its tail short (measured on C3:
    // 3.99 ms with 8 items per warp, 3.83 with 16, 3.81 with 32, 3.87 with 64; C2 loses with more than 8)
    const long long want = (long long)(workers * items_per_warp * (h->ns_max > 1 ? 4 : 1));
    long long nchunks = std::min<long long>(std::max<long long>(1, nb / min_item_blocks), std::max<long long>(1, (want + npv - 1) / npv));
    // a population too small to give every warp an item: cut finer (down to two blocks per item) -- the
    // single-vector call is latency bound and wants all the parallelism there is
    if (npv * nchunks < workers) nchunks = std::min<long long>(std::max<long long>(1, nb / 2), (workers + npv - 1) / npv);
    // the hit bitmap holds PT_MAXBLK bits: 64-point blocks, or (supersampled kernel) 16-point cells
    const long long maxblk = (h->ns_max == 1) ? PT_MAXBLK : PT_MAXBLK * SS_CELL / PT_BLOCK;
    nchunks = std::max<long long>(nchunks, (nb + maxblk - 1) / maxblk);
    long long bpc = (nb + nchunks - 1) / nchunks;
    bpc = std::min<long long>(bpc >= 8 ? (bpc + 7) / 8 * 8 : bpc, maxblk);
    nchunks = (nb + bpc - 1) / bpc;
    P.nchunks = (int)nchunks;
    P.blocks_per_chunk = (int)bpc;
    if (nchunks_out) *nchunks_out = (int)nchunks;
    if (lnl) {
        CU(h->d_partial.reserve((size_t)npv * nchunks * 8));
        P.partial = h->d_partial.as<double>();
    }
    P.ssc = std::min(h->ns_max, PT_SSC_MAX);
    P.frac_tab = ((long long)h->nlc * h->ns_max <= PT_FRAC_MAX) ? 1 : 0;
    const bool s1 = (h->ns_max == 1);
    const size_t nfrac = P.frac_tab ? (size_t)h->nlc * h->ns_max : 0;
    const size_t smem = s1 ? pt_shared_bytes((int)h->nlc, tsize) + pt_warp_bytes(h->recstride, tsize) * PT_WARPS
                           : ss_shared_bytes((int)h->nlc, (int)nfrac, tsize) + ss_warp_bytes(P.ssc, h->recstride, tsize) * SS_WARPS;
    if (smem > 220 * 1024)
        return fail(h, PTB_EINVAL, "npb=%lld passbands x nlc=%lld light curves need %zu bytes of shared memory (> 220 KB)",
                    (long long)h->npb, (long long)h->nlc, smem);

#define PTB_DISPATCH_T(V, S, L, T)                                          \
    do {                                                                    \
        if (s1) return launch_points_t<V, S, L, true, T>(h, P, smem, st);  \
        return launch_points_t<V, S, L, false, T>(h, P, smem, st);         \
    } while (0)
#define PTB_DISPATCH(V, S, L)                        \
    do {                                             \
        if (f32) PTB_DISPATCH_T(V, S, L, float);     \
        PTB_DISPATCH_T(V, S, L, double);             \
    } while (0)
    if (vec == 2) {
        if (single) { if (lnl) PTB_DISPATCH(2, true, true); else PTB_DISPATCH(2, true, false); }
        else        { if (lnl) PTB_DISPATCH(2, false, true); else PTB_DISPATCH(2, false, false); }
    } else {
        if (single) { if (lnl) PTB_DISPATCH(1, true, true); else PTB_DISPATCH(1, true, false); }
        else        { if (lnl) PTB_DISPATCH(1, false, true); else PTB_DISPATCH(1, false, false); }
    }
#undef PTB_DISPATCH_T
#undef PTB_DISPATCH
}

}  // namespace

extern "C" {

// stage + per-vector setup + points kernel (+ eclipse finish) on `st`
static int rr_evaluate_enqueue(ptb_model *h, const ModelArgs &A, void *dflux, size_t count, cudaStream_t st, bool eclipse) {
    Staged D{};
    if (int rc = stage_model_args(h, A, h->npb, h->nep, nullptr, 0, st, D)) return rc;
    mark(h, 0, st);
    if (int rc = launch_rr_setup(h, A, D, st)) return rc;
    mark(h, 1, st);
    mark(h, 2, st);
    if (int rc = launch_points(h, A.npv, dflux, nullptr, st, nullptr)) return rc;
    if (eclipse) {  // pi k^2 - A from the uniform-disk transit shape (model_eclipse.py:72-80)
        k_ecl_finish<<<(unsigned)((count / 2 + 256) / 256), 256, 0, st>>>(static_cast<double *>(dflux), D.k, h->npt, (long long)count);
        h->launches++;
        CU(cudaGetLastError());
    }
    mark(h, 3, st);
    return PTB_OK;
}

static int rr_evaluate_impl(ptb_model *h, int64_t npv, const double *k, int64_t kcols, const double *ld, int64_t nld,
                            const double *istar, const double *t0, const double *p, const double *a, const double *inc,
                            const double *e, const double *w, void *flux, void *stream, bool eclipse) {
    if (!h) return PTB_EINVAL;
    if (int rc = set_device(h)) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    ModelArgs A{npv, kcols, nld, k, ld, istar, t0, p, a, inc, e, w};
    if (int rc = check_model_args(h, eclipse ? "eclipse_evaluate" : "rr_evaluate", A, h->npb)) return rc;
    const size_t count = (size_t)npv * h->npt;
    const size_t esize = h->cfg.precision == 1 ? 4 : 8;  // fp32 mode: `flux` is a float array
    void *dflux = flux;
    const bool direct = flux && is_device_ptr(flux);
    if (!direct) {
        CU(h->d_flux.reserve(count * esize));
        dflux = h->d_flux.ptr;
    }
    bool queued = false;
    if (graph_eligible(h, npv)) {
        // the graph always writes the handle-owned flux buffer, so that one graph serves every output tensor the
        // caller comes with (a small device-to-device copy follows); the eager path writes in place
        void *gflux = dflux;
        if (direct) {
            CU(h->d_flux.reserve(count * esize));
            gflux = h->d_flux.ptr;
        }
        double rs = h->ecl_rstar;
        unsigned long long rsbits;
        memcpy(&rsbits, &rs, 8);
        const GKey key{eclipse ? 2ull : 0ull, (unsigned long long)npv, (unsigned long long)kcols, (unsigned long long)nld, pkey(k),
                       pkey(ld), pkey(istar), pkey(t0), pkey(p), pkey(a), pkey(inc), pkey(e), pkey(w),
                       (unsigned long long)reinterpret_cast<uintptr_t>(gflux), eclipse ? rsbits : 0ull,
                       h->xyc_injected ? 1ull : 0ull, g_alloc_gen.load(), h->data_gen};
        if (auto *g = graph_find(h, key)) {
            Staged D{};
            if (int rc = stage_model_args(h, A, h->npb, h->nep, nullptr, 0, st, D, true)) return rc;
            CU(cudaGraphLaunch(g->exec, st));
            if (int rc = graph_fence(h, st)) return rc;
            h->launches += g->launches;
            h->graph_replays++;
            h->last_npv = npv;
            h->last_npb = h->npb;
            queued = true;
        } else if (graph_seen(h, key)) {  // second call with this signature: capture while enqueueing
            const unsigned long long gen0 = g_alloc_gen.load();
            const int64_t l0 = h->launches;
            if (graph_begin(h) == PTB_OK) {
                const int rc = rr_evaluate_enqueue(h, A, gflux, count, h->cap_stream, eclipse);
                if (int rc2 = graph_end(h, key, st, rc, h->launches - l0, 1, gen0, &queued)) return rc2;
            }
        }
        if (queued && direct) CU(cudaMemcpyAsync(flux, gflux, count * esize, cudaMemcpyDeviceToDevice, st));
    }
    // Large managed host result in its steady state: the population is cut into parts and the delta transfer of
    // part i (PCIe bound, on the side stream) runs under the points kernel of part i+1 (HBM bound).
    ptb_model::HostBinding *B = (!queued && flux && !direct && !eclipse) ? find_binding(h, flux, count) : nullptr;
    if (B && B->valid && B->esize == esize && count * esize >= ((size_t)64 << 20) && !h->profiling) {
        const long long wordlen = 32LL * HD_BLOCK;                          // elements per bitmap word
        long long g = h->npt, b = wordlen;                                  // rows per part must keep parts word aligned
        while (b) { const long long t = g % b; g = b; b = t; }
        const long long rowstep = wordlen / g;
        const int nparts = 4;
        long long rows_per = ((npv + nparts - 1) / nparts + rowstep - 1) / rowstep * rowstep;
        if (rows_per > 0 && rows_per < npv) {
            Staged D{};
            if (int rc = stage_model_args(h, A, h->npb, h->nep, nullptr, 0, st, D)) return rc;
            if (int rc = launch_rr_setup(h, A, D, st)) return rc;
            const long long nwords = ((long long)count + wordlen - 1) / wordlen;
            const size_t lit_bytes = ((size_t)nwords * 4 + 15) & ~size_t(15);
            if (lit_bytes + 16 > B->d_lit.cap) return fail(h, PTB_ESTATE, "host result: bitmap lost");
            unsigned *lit = B->d_lit.as<unsigned>();
            unsigned long long *nwr = reinterpret_cast<unsigned long long *>(static_cast<char *>(B->d_lit.ptr) + lit_bytes);
            const int np = (int)((npv + rows_per - 1) / rows_per);
            while ((int)h->ev_part.size() < np + 1) {
                cudaEvent_t ev;
                CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
                h->ev_part.push_back(ev);
            }
            B->valid = false;
            CU(cudaMemsetAsync(nwr, 0, 8, st));
            for (int ip = 0; ip < np; ++ip) {
                const long long r0 = ip * rows_per, r1 = std::min<long long>(npv, r0 + rows_per);
                if (int rc = launch_points(h, r1 - r0, static_cast<char *>(dflux) + (size_t)r0 * h->npt * esize, nullptr, st, nullptr, r0)) return rc;
                CU(cudaEventRecord(h->ev_part[ip], st));
                CU(cudaStreamWaitEvent(h->side_stream, h->ev_part[ip], 0));
                const long long w0 = r0 * h->npt / wordlen;
                const long long w1 = (r1 == npv) ? nwords : r1 * h->npt / wordlen;
                const unsigned grid = (unsigned)std::min<long long>((w1 - w0 + 7) / 8, (long long)h->sm_count * 4);
                if (esize == 8)
                    k_host_delta<double><<<grid, 256, 0, h->side_stream>>>(static_cast<const double *>(dflux), static_cast<double *>(B->dev),
                                                                            lit, nwr, (long long)count, w0, w1, 0);
                else
                    k_host_delta<float><<<grid, 256, 0, h->side_stream>>>(static_cast<const float *>(dflux), static_cast<float *>(B->dev),
                                                                           lit, nwr, (long long)count, w0, w1, 0);
                h->launches++;
                CU(cudaGetLastError());
            }
            CU(cudaEventRecord(h->ev_part[np], h->side_stream));
            CU(cudaStreamWaitEvent(st, h->ev_part[np], 0));
            CU(h->h_hrstat.reserve(16));
            CU(cudaMemcpyAsync(h->h_hrstat.ptr, nwr, 8, cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            B->valid = true;
            h->hr_last_bytes = (int64_t)(*static_cast<unsigned long long *>(h->h_hrstat.ptr)) * HD_BLOCK * (int64_t)esize;
            h->hr_delta_calls++;
            h->last_flux_count = (int64_t)count;
            return PTB_OK;
        }
    }
    if (!queued)
        if (int rc = rr_evaluate_enqueue(h, A, dflux, count, st, eclipse)) return rc;
    h->last_flux_count = direct ? 0 : (int64_t)count;
    if (flux && !direct) return deliver_host(h, flux, dflux, count, esize, st);
    return PTB_OK;
}

int ptb_rr_evaluate(ptb_model *h, int64_t npv, const double *k, int64_t kcols, const double *ld, int64_t nld,
                    const double *istar, const double *t0, const double *p, const double *a, const double *inc,
                    const double *e, const double *w, void *flux, void *stream) {
    return rr_evaluate_impl(h, npv, k, kcols, ld, nld, istar, t0, p, a, inc, e, w, flux, stream, false);
}

int ptb_eclipse_evaluate(ptb_model *h, int64_t npv, const double *k, const double *t0, const double *p, const double *a,
                         const double *inc, const double *e, const double *w, double rstar, double *flux, void *stream) {
    if (!h) return PTB_EINVAL;
    if (int rc = set_device(h)) return rc;
    if (h->cfg.ldlaw != PTB_LD_UNIFORM || h->cfg.precision != 0)
        return fail(h, PTB_ESTATE, "eclipse_evaluate: the handle must be created with PTB_LD_UNIFORM in fp64 (no stellar limb darkening in an eclipse)");
    if (npv < 1) return fail(h, PTB_ESHAPE, "eclipse_evaluate: npv must be >= 1");
    // a uniform disk has no coefficients: one zero per (vector, passband) keeps the shared argument checks happy
    const size_t nz = (size_t)npv * h->npb;
    if (nz * 8 > h->d_dummy.cap) {
        CU(h->d_dummy.reserve(nz * 8));
        CU(cudaMemset(h->d_dummy.ptr, 0, h->d_dummy.cap));
    }
    h->ecl_mode = true;
    h->ecl_rstar = rstar;
    const int rc = rr_evaluate_impl(h, npv, k, 1, h->d_dummy.as<double>(), 1, nullptr, t0, p, a, inc, e, w, flux, stream, true);
    h->ecl_mode = false;
    return rc;
}


int ptb_es_evaluate(ptb_model *h, int64_t npv, int64_t npb, const double *fratio, const double *k, const double *t0,
                    const double *p, const double *a, const double *inc, const double *e, const double *w,
                    const double *rstar, double *flux, void *stream) {
    if (!h) return PTB_EINVAL;
    if (int rc = set_device(h)) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (h->cfg.ldlaw != PTB_LD_UNIFORM || h->cfg.precision != 0)
        return fail(h, PTB_ESTATE, "es_evaluate: the handle must be created with PTB_LD_UNIFORM in fp64");
    if (!h->has_data) return fail(h, PTB_ESTATE, "es_evaluate: call set_data first");
    if (h->nlc != 1 || h->nep != 1)
        return fail(h, PTB_ESTATE, "es_evaluate: eclipse spectroscopy uses one light curve and one epoch (model_ecspec.py:13-14), got nlc=%lld nep=%lld",
                    (long long)h->nlc, (long long)h->nep);
    if (npv < 1 || npb < 1) return fail(h, PTB_ESHAPE, "es_evaluate: npv and npb must be >= 1");
    if (!fratio || !k || !t0 || !p || !a || !inc || !e || !w || !rstar) return fail(h, PTB_EINVAL, "es_evaluate: null parameter array");
    if ((double)npv * (double)npb * (double)h->npt >= 9.0e18 || npv * ((npb + ES_CH - 1) / ES_CH) > 0x7fffffffLL)
        return fail(h, PTB_EINVAL, "es_evaluate: output too large");
    // a uniform disk has no coefficients: one zero per vector keeps the shared argument checks happy
    if ((size_t)npv * h->npb * 8 > h->d_dummy.cap) {
        CU(h->d_dummy.reserve((size_t)npv * h->npb * 8));
        CU(cudaMemset(h->d_dummy.ptr, 0, h->d_dummy.cap));
    }
    ModelArgs A{npv, 1, 1, k, h->d_dummy.as<double>(), nullptr, t0, p, a, inc, e, w};
    if (int rc = check_model_args(h, "es_evaluate", A, h->npb)) return rc;
    // every host argument of this call in ONE pinned block: the model arguments, the stellar radii and the flux ratios
    Stager S(h, st);
    const size_t n = (size_t)npv;
    auto rk = S.add(k, n * 8), rt0 = S.add(t0, n * 8), rp = S.add(p, n * 8), ra = S.add(a, n * 8), ri = S.add(inc, n * 8),
         re = S.add(e, n * 8), rw = S.add(w, n * 8), rr = S.add(rstar, n * 8), rf = S.add(fratio, n * npb * 8);
    if (int rc = S.commit()) return rc;
    Staged D{};
    D.k = S.get<double>(rk); D.ld = h->d_dummy.as<double>(); D.istar = nullptr; D.t0 = S.get<double>(rt0); D.p = S.get<double>(rp);
    D.a = S.get<double>(ra); D.inc = S.get<double>(ri); D.e = S.get<double>(re); D.w = S.get<double>(rw); D.sigma = nullptr;
    const double *d_rstar = S.get<double>(rr), *d_fratio = S.get<double>(rf);
    h->ecl_mode = true;
    h->ecl_rstar_v = d_rstar;
    mark(h, 0, st);
    int rc = launch_rr_setup(h, A, D, st);
    h->ecl_mode = false;
    h->ecl_rstar_v = nullptr;
    if (rc) return rc;
    mark(h, 1, st);
    // uniform-disk eclipse shape F[npv, npt], then the per-channel expansion
    CU(h->d_tsgeo.reserve(n * h->npt * 8 + 64));
    double *shape = h->d_tsgeo.as<double>();
    if (int rc2 = launch_points(h, npv, shape, nullptr, st, nullptr)) return rc2;
    const size_t count = n * npb * h->npt;
    double *dflux = flux;
    const bool direct = flux && is_device_ptr(flux);
    if (!direct) {
        CU(h->d_flux.reserve(count * 8));
        dflux = h->d_flux.as<double>();
    }
    const int nchunks = (int)((npb + ES_CH - 1) / ES_CH);
    mark(h, 2, st);
    k_es_expand<<<(unsigned)(npv * nchunks), 256, 0, st>>>(shape, d_fratio, D.k, dflux, h->npt, (int)npb, nchunks);
    h->launches++;
    CU(cudaGetLastError());
    mark(h, 3, st);
    h->last_npv = 0;  // RoadRunner stage taps do not describe this evaluation
    h->last_flux_count = direct ? 0 : (int64_t)count;
    if (flux && !direct) return deliver_host(h, flux, dflux, count, 8, st);
    return PTB_OK;
}

// stage + per-vector setup + fused likelihood kernels on `st`
static int rr_lnlike_enqueue(ptb_model *h, const ModelArgs &A, const double *sigma, LnlOut out, cudaStream_t st) {
    const int64_t npv = A.npv;
    Staged D{};
    if (int rc = stage_model_args(h, A, h->npb, h->nep, sigma, h->nblocks, st, D)) return rc;
    mark(h, 0, st);
    if (int rc = launch_rr_setup(h, A, D, st)) return rc;
    const long long nsig = (long long)npv * h->nblocks;
    CU(h->d_isig2.reserve(nsig * 8));
    k_inv_sigma2<<<(unsigned)((nsig + 255) / 256), 256, 0, st>>>(D.sigma, nsig, h->d_isig2.as<double>());
    h->launches++;
    mark(h, 1, st);
    int nchunks = 1;
    mark(h, 2, st);
    if (int rc = launch_points(h, npv, nullptr, h->d_isig2.as<double>(), st, &nchunks)) return rc;
    mark(h, 3, st);
    out.done = h->d_work.as<int>() + WORK_FINISH;
    out.err = h->d_work.as<int>() + WORK_GATHER_ERR;
    out.timeout_ns = GATHER_TIMEOUT_NS;
    // with arrival flags (fused all-gather) the kernel's last CTA also waits for every peer's shard of this step
    k_lnl_finish<<<(unsigned)((npv + 127) / 128), 128, 0, st>>>(h->d_partial.as<double>(), nchunks, D.sigma,
                                                                 h->d_nblk.as<double>(), (int)h->nblocks, (int)npv, out);
    h->launches++;
    CU(cudaGetLastError());
    return PTB_OK;
}

static int rr_lnlike_impl(ptb_model *h, int64_t npv, const double *k, int64_t kcols, const double *ld, int64_t nld,
                          const double *istar, const double *t0, const double *p, const double *a, const double *inc,
                          const double *e, const double *w, const double *sigma, double *lnl, double *const *peers,
                          uint64_t *const *peer_flags, uint64_t seq, int world, int rank, void *stream) {
    if (!h) return PTB_EINVAL;
    if (int rc = set_device(h)) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (!h->has_obs) return fail(h, PTB_ESTATE, "rr_lnlike: call set_obs first");
    if (!sigma || (!lnl && !peers)) return fail(h, PTB_EINVAL, "rr_lnlike: sigma / lnl is null");
    if (peers) {
        if (world < 1 || world > LNL_MAXPEERS || rank < 0 || rank >= world)
            return fail(h, PTB_EINVAL, "rr_lnlike_allgather: world=%d (max %d), rank=%d", world, LNL_MAXPEERS, rank);
        for (int r = 0; r < world; ++r)
            if (!peers[r] || (peer_flags && !peer_flags[r])) return fail(h, PTB_EINVAL, "rr_lnlike_allgather: peer buffer %d is null", r);
        if (peer_flags && seq == 0) return fail(h, PTB_EINVAL, "rr_lnlike_allgather: step numbers start at 1 (flags are zero-initialised)");
    }
    ModelArgs A{npv, kcols, nld, k, ld, istar, t0, p, a, inc, e, w};
    if (int rc = check_model_args(h, "rr_lnlike", A, h->npb)) return rc;
    LnlOut out{};
    const bool direct = peers || is_device_ptr(lnl);
    const bool use_graph = graph_eligible(h, npv) && !peer_flags;  // the step number changes every call
    bool via_own = false;   // the result lands in the handle's buffer first (host output, or graph mode: one graph
                            // for every output tensor) and is copied out afterwards
    if (peers) {
        out.nout = world;
        out.rank = rank;
        out.seq = seq;
        for (int r = 0; r < world; ++r) {
            out.ptr[r] = peers[r] + (size_t)rank * npv;
            out.flag[r] = peer_flags ? reinterpret_cast<unsigned long long *>(peer_flags[r]) : nullptr;
        }
    } else {
        out.nout = 1;
        out.ptr[0] = lnl;
        if (!direct || use_graph) {
            CU(h->d_lnl.reserve(npv * 8));
            out.ptr[0] = h->d_lnl.as<double>();
            via_own = true;
        }
    }
    bool queued = false;
    if (use_graph) {
        GKey key{1ull, (unsigned long long)npv, (unsigned long long)kcols, (unsigned long long)nld, pkey(k), pkey(ld), pkey(istar),
                 pkey(t0), pkey(p), pkey(a), pkey(inc), pkey(e), pkey(w), pkey(sigma), h->xyc_injected ? 1ull : 0ull, g_alloc_gen.load(),
                 h->data_gen, (unsigned long long)out.nout};
        for (int r = 0; r < out.nout; ++r) key.push_back((unsigned long long)reinterpret_cast<uintptr_t>(out.ptr[r]));
        if (auto *g = graph_find(h, key)) {
            Staged D{};
            if (int rc = stage_model_args(h, A, h->npb, h->nep, sigma, h->nblocks, st, D, true)) return rc;
            CU(cudaGraphLaunch(g->exec, st));
            if (int rc = graph_fence(h, st)) return rc;
            h->launches += g->launches;
            h->graph_replays++;
            h->last_npv = npv;
            h->last_npb = h->npb;
            queued = true;
        } else if (graph_seen(h, key)) {
            const unsigned long long gen0 = g_alloc_gen.load();
            const int64_t l0 = h->launches;
            if (graph_begin(h) == PTB_OK) {
                const int rc = rr_lnlike_enqueue(h, A, sigma, out, h->cap_stream);
                if (int rc2 = graph_end(h, key, st, rc, h->launches - l0, 1, gen0, &queued)) return rc2;
            }
        }
    }
    if (!queued)
        if (int rc = rr_lnlike_enqueue(h, A, sigma, out, st)) return rc;
    if (via_own) {
        CU(cudaMemcpyAsync(lnl, out.ptr[0], npv * 8, direct ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
        if (!direct) CU(cudaStreamSynchronize(st));
    }
    return PTB_OK;
}

int ptb_rr_lnlike(ptb_model *h, int64_t npv, const double *k, int64_t kcols, const double *ld, int64_t nld,
                  const double *istar, const double *t0, const double *p, const double *a, const double *inc,
                  const double *e, const double *w, const double *sigma, double *lnl, void *stream) {
    return rr_lnlike_impl(h, npv, k, kcols, ld, nld, istar, t0, p, a, inc, e, w, sigma, lnl, nullptr, nullptr, 0, 1, 0, stream);
}

int ptb_rr_lnlike_allgather(ptb_model *h, int64_t npv, const double *k, int64_t kcols, const double *ld, int64_t nld,
                            const double *istar, const double *t0, const double *p, const double *a, const double *inc,
                            const double *e, const double *w, const double *sigma, double *const *peer_bufs,
                            uint64_t *const *peer_flags, uint64_t seq, int32_t world, int32_t rank, void *stream) {
    if (!peer_bufs) return h ? fail(h, PTB_EINVAL, "rr_lnlike_allgather: peer_bufs is null") : PTB_EINVAL;
    return rr_lnlike_impl(h, npv, k, kcols, ld, nld, istar, t0, p, a, inc, e, w, sigma, nullptr, peer_bufs, peer_flags, seq, world, rank, stream);
}

int ptb_gather_status(ptb_model *h, int32_t *timed_out_rank) {
    if (!h || !timed_out_rank) return PTB_EINVAL;
    if (int rc = set_device(h)) return rc;
    *timed_out_rank = -1;
    if (!h->d_work.ptr) return PTB_OK;
    int v = 0;
    CU(cudaMemcpy(&v, h->d_work.as<int>() + WORK_GATHER_ERR, 4, cudaMemcpyDeviceToHost));
    if (v) {
        *timed_out_rank = v - 1;
        CU(cudaMemset(h->d_work.as<int>() + WORK_GATHER_ERR, 0, 4));
        return fail(h, PTB_ESTATE, "fused all-gather: rank %d never published its shard (waited %.0f s)", v - 1, GATHER_TIMEOUT_NS * 1e-9);
    }
    return PTB_OK;
}

int ptb_lnlike_normal(ptb_model *h, int64_t npv, const double *model, const double *sigma, double *lnl, void *stream) {
    if (!h) return PTB_EINVAL;
    if (int rc = set_device(h)) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (!h->has_obs) return fail(h, PTB_ESTATE, "lnlike_normal: call set_obs first");
    if (!model || !sigma || !lnl || npv < 1) return fail(h, PTB_EINVAL, "lnlike_normal: null argument or npv < 1");
    Stager S(h, st);
    auto rm = S.add(model, (size_t)npv * h->npt * 8);
    auto rs = S.add(sigma, (size_t)npv * h->nblocks * 8);
    if (int rc = S.commit()) return rc;
    const double *dm = S.get<double>(rm), *ds = S.get<double>(rs);
    const long long nsig = (long long)npv * h->nblocks;
    CU(h->d_isig2.reserve(nsig * 8));
    CU(h->d_partial.reserve(npv * 8));
    k_inv_sigma2<<<(unsigned)((nsig + 255) / 256), 256, 0, st>>>(ds, nsig, h->d_isig2.as<double>());
    k_lnl_model<<<(unsigned)npv, 256, 0, st>>>(dm, h->d_obs, h->blk_trivial ? nullptr : h->d_blk.as<int32_t>(),
                                               h->d_isig2.as<double>(), h->npt, (int)h->nblocks, h->d_partial.as<double>(),
                                               BaselineParams{});
    const bool direct = is_device_ptr(lnl);
    double *dl = lnl;
    if (!direct) {
        CU(h->d_lnl.reserve(npv * 8));
        dl = h->d_lnl.as<double>();
    }
    LnlOut out{};
    out.nout = 1;
    out.ptr[0] = dl;
    k_lnl_finish<<<(unsigned)((npv + 127) / 128), 128, 0, st>>>(h->d_partial.as<double>(), 1, ds, h->d_nblk.as<double>(),
                                                                 (int)h->nblocks, (int)npv, out);
    h->launches += 3;
    CU(cudaGetLastError());
    if (!direct) {
        CU(cudaMemcpyAsync(lnl, dl, npv * 8, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
    }
    return PTB_OK;
}

int ptb_get_stage(ptb_model *h, int32_t stage, double *out) {
    if (!h || !out) return PTB_EINVAL;
    if (int rc = set_device(h)) return rc;
    if (h->last_npv == 0) return fail(h, PTB_ESTATE, "get_stage: no evaluation has run yet");
    CU(cudaDeviceSynchronize());
    const int64_t npv = h->last_npv, npb = h->last_npb;
    const int ng = h->cfg.ng, nz = h->nz, lds = (ng + 4 + 1) & ~1;
    const cudaMemcpyKind kind = cudaMemcpyDefault;
    switch (stage) {
    case PTB_STAGE_LDP: CU(cudaMemcpy(out, h->d_ldp.ptr, (size_t)npv * npb * nz * 8, kind)); break;
    case PTB_STAGE_ISTAR: CU(cudaMemcpy(out, h->d_istar.ptr, (size_t)npv * npb * 8, kind)); break;
    case PTB_STAGE_LDM:
        for (int64_t pb = 0; pb < npb; ++pb)  // rows of passband pb: record pitch in, [npv][npb][ng] out
            CU(cudaMemcpy2D(out + pb * ng, (size_t)npb * ng * 8, h->d_rec.as<double>() + h->rec_ld + pb * lds,
                            (size_t)h->recstride * 8, ng * 8, npv, kind));
        break;
    case PTB_STAGE_XYC: CU(cudaMemcpy2D(out, 80, h->d_rec.ptr, (size_t)h->recstride * 8, 80, npv, kind)); break;
    case PTB_STAGE_BBOX:
        CU(cudaMemcpy2D(out, 16, h->d_rec.as<double>() + ORB_T1, (size_t)h->recstride * 8, 16, npv, kind));
        break;
    case PTB_STAGE_GOOD: {  // valid orbit/geometry (k_rr_orbit) and a non-NaN limb-darkening profile (k_rr_ldm)
        std::vector<double> g(npv), l(npv);
        CU(cudaMemcpy2D(g.data(), 8, h->d_rec.as<double>() + ORB_GOOD, (size_t)h->recstride * 8, 8, npv, cudaMemcpyDeviceToHost));
        CU(cudaMemcpy2D(l.data(), 8, h->d_rec.as<double>() + ORB_LDNAN, (size_t)h->recstride * 8, 8, npv, cudaMemcpyDeviceToHost));
        for (int64_t i = 0; i < npv; ++i) g[i] = (g[i] != 0.0 && !(l[i] != 0.0)) ? 1.0 : 0.0;
        CU(cudaMemcpy(out, g.data(), npv * 8, kind));
        break;
    }
    default: return fail(h, PTB_EINVAL, "get_stage: unknown stage %d", stage);
    }
    return PTB_OK;
}

int ptb_inject_xyc(ptb_model *h, const double *xyc, int64_t npv) {
    if (!h) return PTB_EINVAL;
    if (int rc = set_device(h)) return rc;
    h->data_gen++;
    if (!xyc) {
        h->xyc_injected = false;
        h->xyc_npv = 0;
        return PTB_OK;
    }
    if (npv < 1) return fail(h, PTB_ESHAPE, "inject_xyc: npv must be >= 1");
    CU(h->d_xyc.reserve(npv * 80));
    CU(cudaMemcpy(h->d_xyc.ptr, xyc, npv * 80, cudaMemcpyDefault));
    h->xyc_injected = true;
    h->xyc_npv = npv;
    return PTB_OK;
}

int ptb_rr_derivatives(ptb_model *h, int64_t npv, int64_t nb, int64_t pb, const double *b, double *dfdk, double *dfdb, void *stream) {
    if (!h) return PTB_EINVAL;
    if (int rc = set_device(h)) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (h->last_npv == 0) return fail(h, PTB_ESTATE, "rr_derivatives: no RoadRunner evaluation has run yet");
    if (npv != h->last_npv) return fail(h, PTB_ESHAPE, "rr_derivatives: npv=%lld but the last evaluation had %lld parameter vectors", (long long)npv, (long long)h->last_npv);
    if (nb < 1 || !b || (!dfdk && !dfdb)) return fail(h, PTB_EINVAL, "rr_derivatives: null argument or nb < 1");
    if (pb < 0 || pb >= h->last_npb) return fail(h, PTB_EINVAL, "rr_derivatives: passband %lld outside [0,%lld)", (long long)pb, (long long)h->last_npb);
    const size_t n = (size_t)npv * nb;
    Stager S(h, st);
    auto rb = S.add(b, n * 8);
    if (int rc = S.commit()) return rc;
    const bool hk = dfdk && !is_device_ptr(dfdk), hb = dfdb && !is_device_ptr(dfdb);
    CU(h->d_partial.reserve(2 * n * 8));
    double *dk = dfdk ? (hk ? h->d_partial.as<double>() : dfdk) : nullptr;
    double *db = dfdb ? (hb ? h->d_partial.as<double>() + n : dfdb) : nullptr;
    const int ng = h->cfg.ng, lds = (ng + 4 + 1) & ~1;
    k_rr_derivs<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(h->d_rec.as<double>(), h->recstride, h->rec_ld, lds, ng, (int)pb, h->dg,
                                                              S.get<double>(rb), (int)nb, (int)npv, dk, db);
    h->launches++;
    CU(cudaGetLastError());
    if (hk) CU(cudaMemcpyAsync(dfdk, dk, n * 8, cudaMemcpyDeviceToHost, st));
    if (hb) CU(cudaMemcpyAsync(dfdb, db, n * 8, cudaMemcpyDeviceToHost, st));
    if (hk || hb) CU(cudaStreamSynchronize(st));
    return PTB_OK;
}

int ptb_flux_device_ptr(ptb_model *h, void **ptr, int64_t *count) {
    if (!h || !ptr || !count) return PTB_EINVAL;
    *ptr = h->d_flux.ptr;
    *count = h->last_flux_count;
    return PTB_OK;
}

int ptb_bind_host_result(ptb_model *h, void *buf, int64_t count) {
    if (!h) return PTB_EINVAL;
    if (int rc = set_device(h)) return rc;
    if (!buf) return ptb_unbind_host_result(h, nullptr);
    if (count < 1) return fail(h, PTB_ESHAPE, "bind_host_result: count must be >= 1");
    void *dp = nullptr;
    if (cudaHostGetDevicePointer(&dp, buf, 0) != cudaSuccess || !dp) {
        cudaGetLastError();
        return fail(h, PTB_EINVAL, "bind_host_result: the buffer is not page-locked memory mapped into the device "
                                   "(allocate it with ptb_host_alloc)");
    }
    for (auto *b : h->hr)
        if (b->buf == buf) {  // re-binding: whatever the buffer held is no longer trusted
            b->dev = dp;
            b->count = count;
            b->valid = false;
            return PTB_OK;
        }
    if (h->hr.size() >= 8) return fail(h, PTB_ESTATE, "bind_host_result: at most 8 host result buffers can be bound (unbind one first)");
    auto *b = new ptb_model::HostBinding();
    b->buf = buf;
    b->dev = dp;
    b->count = count;
    h->hr.push_back(b);
    return PTB_OK;
}

int ptb_unbind_host_result(ptb_model *h, void *buf) {
    if (!h) return PTB_EINVAL;
    if (int rc = set_device(h)) return rc;
    for (size_t i = 0; i < h->hr.size();) {
        if (!buf || h->hr[i]->buf == buf) {
            CU(cudaDeviceSynchronize());  // a transfer into it may still be running
            h->hr[i]->d_lit.release();
            delete h->hr[i];
            h->hr.erase(h->hr.begin() + i);
        } else {
            ++i;
        }
    }
    return PTB_OK;
}

int ptb_host_result_stats(const ptb_model *h, int64_t *last_bytes, int64_t *delta_calls, int64_t *full_calls) {
    if (!h) return PTB_EINVAL;
    if (last_bytes) *last_bytes = h->hr_last_bytes;
    if (delta_calls) *delta_calls = h->hr_delta_calls;
    if (full_calls) *full_calls = h->hr_full_calls;
    return PTB_OK;
}

// Page-locked host memory for results.  Large buffers (>= 32 MB) are carved from 2 MB-aligned anonymous memory with
// transparent huge pages requested (madvise) and then registered with CUDA: the delta transfer writes 128-byte bursts
// scattered over the whole array, and every 4 KB page it touches costs an IOMMU / PCIe address translation -- 512x
// fewer with 2 MB pages.  First touch happens here, on the calling thread, so the memory lands on the NUMA node the
// caller is bound to (see pytransit_b200.distributed.bind_to_gpu_numa).  PTB_HOST_HUGEPAGES=0 restores cudaMallocHost.
namespace {
struct HostBlock { void *p; size_t bytes; };
std::vector<HostBlock> g_host_blocks;   // registered (mmap-backed) allocations; everything else is cudaMallocHost memory
std::mutex g_host_mu;
}  // namespace

int ptb_host_alloc(void **ptr, size_t bytes) {
    if (!ptr) return PTB_EINVAL;
    static const bool huge = [] {
        const char *e = getenv("PTB_HOST_HUGEPAGES");
        return !(e && atoi(e) == 0);
    }();
    if (huge && bytes >= ((size_t)32 << 20)) {
        const size_t two_mb = (size_t)2 << 20, len = (bytes + two_mb - 1) & ~(two_mb - 1);
        void *raw = mmap(nullptr, len + two_mb, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        if (raw != MAP_FAILED) {
            char *p = reinterpret_cast<char *>((reinterpret_cast<uintptr_t>(raw) + two_mb - 1) & ~(uintptr_t)(two_mb - 1));
            // give back the unaligned head and tail so that munmap(p, len) frees everything later
            if (p > static_cast<char *>(raw)) munmap(raw, p - static_cast<char *>(raw));
            char *end = static_cast<char *>(raw) + len + two_mb;
            if (end > p + len) munmap(p + len, end - (p + len));
            madvise(p, len, MADV_HUGEPAGE);
            for (size_t o = 0; o < len; o += 4096) p[o] = 0;   // first touch (NUMA placement, huge-page faults)
            if (cudaHostRegister(p, len, cudaHostRegisterPortable | cudaHostRegisterMapped) == cudaSuccess) {
                std::lock_guard<std::mutex> lk(g_host_mu);
                g_host_blocks.push_back({p, len});
                *ptr = p;
                return PTB_OK;
            }
            cudaGetLastError();
            munmap(p, len);
        }
    }
    cudaError_t e = cudaMallocHost(ptr, bytes ? bytes : 1);
    if (e != cudaSuccess) {
        cudaGetLastError();
        g_create_error = std::string("cudaMallocHost failed: ") + cudaGetErrorString(e);
        return PTB_ENOMEM;
    }
    return PTB_OK;
}

int ptb_host_free(void *ptr) {
    if (!ptr) return PTB_OK;
    {
        std::lock_guard<std::mutex> lk(g_host_mu);
        for (size_t i = 0; i < g_host_blocks.size(); ++i)
            if (g_host_blocks[i].p == ptr) {
                const size_t len = g_host_blocks[i].bytes;
                g_host_blocks.erase(g_host_blocks.begin() + i);
                const bool ok = cudaHostUnregister(ptr) == cudaSuccess;
                munmap(ptr, len);
                return ok ? PTB_OK : PTB_ECUDA;
            }
    }
    return cudaFreeHost(ptr) == cudaSuccess ? PTB_OK : PTB_ECUDA;
}

int64_t ptb_launch_count(const ptb_model *h) { return h ? h->launches : 0; }

int ptb_measure_fp64_peak(ptb_model *h, double *tflops) {
    if (!h || !tflops) return PTB_EINVAL;
    if (int rc = set_device(h)) return rc;
    CU(h->d_dummy.reserve(64));
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    const int iters = 4096, grid = h->sm_count * 8;
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        CU(cudaEventRecord(e0, 0));
        k_dfma_peak<<<grid, 256>>>(h->d_dummy.as<double>(), iters, 0.999999, 1e-9);
        CU(cudaEventRecord(e1, 0));
        CU(cudaEventSynchronize(e1));
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, e0, e1));
        const double tf = 2.0 * 64.0 * iters * 256.0 * grid / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;   // first launch: warm-up
    }
    h->launches += 5;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *tflops = best;
    return PTB_OK;
}

int ptb_set_graphs(ptb_model *h, int32_t enabled) {
    if (!h) return PTB_EINVAL;
    h->graphs_enabled = enabled != 0;
    return PTB_OK;
}

int ptb_graph_stats(const ptb_model *h, int64_t *replays, int64_t *captures) {
    if (!h) return PTB_EINVAL;
    if (replays) *replays = h->graph_replays;
    if (captures) *captures = h->graph_captures;
    return PTB_OK;
}

int ptb_set_profiling(ptb_model *h, int32_t enabled) {
    if (!h) return PTB_EINVAL;
    if (int rc = set_device(h)) return rc;
    if (enabled && h->tev.empty()) {
        h->tev.assign(ptb_model::TRING * 4, nullptr);
        for (auto &e : h->tev) CU(cudaEventCreate(&e));
    }
    h->profiling = enabled != 0;
    h->tcalls = 0;
    return PTB_OK;
}

int ptb_timing_summary(ptb_model *h, int64_t *ncalls, double *setup_ms_total, double *points_ms_total) {
    if (!h) return PTB_EINVAL;
    if (int rc = set_device(h)) return rc;
    if (h->tev.empty() || h->tcalls == 0) return fail(h, PTB_ESTATE, "timing: profiling is off or no call has been timed");
    const int64_t n = std::min<int64_t>(h->tcalls, ptb_model::TRING);
    double a = 0, b = 0;
    for (int64_t c = h->tcalls - n; c < h->tcalls; ++c) {
        cudaEvent_t *e = &h->tev[(c % ptb_model::TRING) * 4];
        CU(cudaEventSynchronize(e[3]));
        float x = 0, y = 0;
        CU(cudaEventElapsedTime(&x, e[0], e[1]));
        CU(cudaEventElapsedTime(&y, e[2], e[3]));
        a += x;
        b += y;
    }
    if (ncalls) *ncalls = n;
    if (setup_ms_total) *setup_ms_total = a;
    if (points_ms_total) *points_ms_total = b;
    return PTB_OK;
}

int ptb_last_timing(ptb_model *h, double *setup_ms, double *points_ms) {
    if (!h) return PTB_EINVAL;
    if (int rc = set_device(h)) return rc;
    if (h->tev.empty() || h->tcalls == 0) return fail(h, PTB_ESTATE, "last_timing: profiling is off or no call has been timed");
    cudaEvent_t *e = &h->tev[((h->tcalls - 1) % ptb_model::TRING) * 4];
    CU(cudaEventSynchronize(e[3]));
    float a = 0, b = 0;
    CU(cudaEventElapsedTime(&a, e[0], e[1]));
    CU(cudaEventElapsedTime(&b, e[2], e[3]));
    if (setup_ms) *setup_ms = a;
    if (points_ms) *points_ms = b;
    return PTB_OK;
}

int ptb_synchronize(ptb_model *h, void *stream) {
    if (!h) return PTB_EINVAL;
    if (int rc = set_device(h)) return rc;
    CU(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
    return PTB_OK;
}

}  // extern "C"

namespace {

static_assert(sizeof(LpfLayout) == sizeof(ptb_lpf_layout), "LpfLayout must mirror ptb_lpf_layout");

// pvp -> device arrays of the RoadRunner arguments (k_lpf_map); returns device pointers in A / sigma
int lpf_map(ptb_model *h, const char *who, const double *pvp, int64_t npv, const ptb_lpf_layout *lay, bool want_sigma,
            cudaStream_t st, ModelArgs &A, const double *&sigma, const double **pvp_dev = nullptr) {
    if (!h->has_data) return fail(h, PTB_ESTATE, "%s: call set_data first", who);
    if (!pvp || !lay || npv < 1) return fail(h, PTB_EINVAL, "%s: null argument or npv < 1", who);
    if (h->cfg.ldlaw == PTB_LD_PROFILES) return fail(h, PTB_ESTATE, "%s: needs a named limb-darkening law", who);
    const int64_t npb = h->npb;
    const ptb_lpf_layout &L = *lay;
    auto col = [&](int32_t i, int32_t n) { return i >= 0 && n >= 1 && (int64_t)i + n <= L.npar; };
    if (L.ntc != h->nep)
        return fail(h, PTB_ESTATE, "%s: %d transit-centre columns (ntc) for a dataset with %lld epochs", who, L.ntc, (long long)h->nep);
    if (L.npar < 5 || !col(L.i_tc, L.ntc) || !col(L.i_p, 1) || !col(L.i_rho, 1) || !col(L.i_b, 1))
        return fail(h, PTB_ESHAPE, "%s: orbit columns outside the %d-column parameter array", who, L.npar);
    if (!(L.nk2 == 1 || L.nk2 == npb) || !col(L.i_k2, L.nk2))
        return fail(h, PTB_ESHAPE, "%s: nk2=%d must be 1 or npb=%lld and lie inside the parameter array", who, L.nk2, (long long)npb);
    if (L.nldc < 1 || !col(L.i_ld, (int32_t)(npb * L.nldc)) || (L.ld_map && L.nldc != 2))
        return fail(h, PTB_ESHAPE, "%s: limb-darkening block (npb=%lld x nldc=%d at column %d) invalid", who, (long long)npb, L.nldc, L.i_ld);
    if ((L.i_secw >= 0) != (L.i_sesw >= 0) || (L.i_secw >= 0 && (!col(L.i_secw, 1) || !col(L.i_sesw, 1))))
        return fail(h, PTB_ESHAPE, "%s: secw / sesw columns invalid", who);
    if (want_sigma && (L.nloge != h->nblocks || !col(L.i_loge, L.nloge)))
        return fail(h, PTB_ESHAPE, "%s: %d log10-sigma columns given, the observations have %lld noise blocks", who, L.nloge, (long long)h->nblocks);

    Stager S(h, st);
    auto rp = S.add(pvp, (size_t)npv * L.npar * 8);
    if (int rc = S.commit()) return rc;
    const size_t n = (size_t)npv;
    const size_t nsig = want_sigma ? (size_t)L.nloge : 0;
    const size_t total = n * (L.nk2 + npb * L.nldc + L.ntc + 5 + nsig);
    CU(h->d_lpf.reserve(total * 8));
    double *b = h->d_lpf.as<double>();
    LpfMapParams P{};
    P.pvp = S.get<double>(rp);
    if (pvp_dev) *pvp_dev = P.pvp;
    P.k = b; b += n * L.nk2;
    P.ldc = b; b += n * npb * L.nldc;
    P.t0 = b; b += n * L.ntc;
    P.p = b; b += n;
    P.a = b; b += n;
    P.inc = b; b += n;
    P.e = b; b += n;
    P.w = b; b += n;
    P.sigma = want_sigma ? b : nullptr;
    P.npv = (int)npv; P.npb = (int)npb;
    memcpy(&P.L, lay, sizeof(LpfLayout));
    k_lpf_map<<<(unsigned)((npv + 127) / 128), 128, 0, st>>>(P);
    h->launches++;
    CU(cudaGetLastError());
    A = ModelArgs{npv, L.nk2, L.nldc, P.k, P.ldc, nullptr, P.t0, P.p, P.a, P.inc, P.e, P.w};
    sigma = P.sigma;
    return PTB_OK;
}

// baseline description for the kernels; fails when the layout asks for a baseline that was not registered
int baseline_params(ptb_model *h, const char *who, const double *pvp_dev, const ptb_lpf_layout *lay, BaselineParams &B) {
    B = BaselineParams{};
    if (lay->i_bl < 0) return PTB_OK;
    if (h->bl_nbasis == 0) return fail(h, PTB_ESTATE, "%s: the layout names baseline columns (i_bl=%d) but no baseline is registered (ptb_set_baseline)", who, lay->i_bl);
    if (h->cfg.precision != 0) return fail(h, PTB_ESTATE, "%s: the baseline needs an fp64 handle", who);
    B.pvp = pvp_dev;
    B.basis = h->d_basis.as<double>();
    B.lcids = h->d_lcids;
    B.cstart = h->d_blmeta.as<int32_t>();
    B.ncoef = B.cstart + h->nlc;
    B.npt = h->npt;
    B.npar = lay->npar;
    B.i_bl = lay->i_bl;
    return PTB_OK;
}

}  // namespace

extern "C" {

int ptb_lpf_transit_model(ptb_model *h, const double *pvp, int64_t npv, const ptb_lpf_layout *lay, void *flux, void *stream) {
    if (!h) return PTB_EINVAL;
    if (int rc = set_device(h)) return rc;
    ModelArgs A{};
    const double *sigma = nullptr;
    if (int rc = lpf_map(h, "lpf_transit_model", pvp, npv, lay, false, static_cast<cudaStream_t>(stream), A, sigma)) return rc;
    return ptb_rr_evaluate(h, npv, A.k, A.kcols, A.ld, A.nld, nullptr, A.t0, A.p, A.a, A.inc, A.e, A.w, flux, stream);
}

int ptb_lpf_lnlike(ptb_model *h, const double *pvp, int64_t npv, const ptb_lpf_layout *lay, double *lnl, void *stream) {
    if (!h) return PTB_EINVAL;
    if (int rc = set_device(h)) return rc;
    if (!h->has_obs) return fail(h, PTB_ESTATE, "lpf_lnlike: call set_obs first");
    if (!lnl) return fail(h, PTB_EINVAL, "lpf_lnlike: lnl is null");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    ModelArgs A{};
    const double *sigma = nullptr, *pvp_dev = nullptr;
    if (int rc = lpf_map(h, "lpf_lnlike", pvp, npv, lay, true, st, A, sigma, &pvp_dev)) return rc;
    if (lay->i_bl < 0)
        return ptb_rr_lnlike(h, npv, A.k, A.kcols, A.ld, A.nld, nullptr, A.t0, A.p, A.a, A.inc, A.e, A.w, sigma, lnl, stream);
    // with a baseline: transit flux on the device, then the likelihood of baseline * flux (lpf.py:445-475)
    BaselineParams B{};
    if (int rc = baseline_params(h, "lpf_lnlike", pvp_dev, lay, B)) return rc;
    const size_t count = (size_t)npv * h->npt;
    CU(h->d_flux.reserve(count * 8));
    if (int rc = ptb_rr_evaluate(h, npv, A.k, A.kcols, A.ld, A.nld, nullptr, A.t0, A.p, A.a, A.inc, A.e, A.w, h->d_flux.ptr, stream)) return rc;
    const long long nsig = (long long)npv * h->nblocks;
    CU(h->d_isig2.reserve(nsig * 8));
    CU(h->d_partial.reserve(npv * 8));
    k_inv_sigma2<<<(unsigned)((nsig + 255) / 256), 256, 0, st>>>(sigma, nsig, h->d_isig2.as<double>());
    k_lnl_model<<<(unsigned)npv, 256, 0, st>>>(h->d_flux.as<double>(), h->d_obs, h->blk_trivial ? nullptr : h->d_blk.as<int32_t>(),
                                               h->d_isig2.as<double>(), h->npt, (int)h->nblocks, h->d_partial.as<double>(), B);
    const bool direct = is_device_ptr(lnl);
    double *dl = lnl;
    if (!direct) {
        CU(h->d_lnl.reserve(npv * 8));
        dl = h->d_lnl.as<double>();
    }
    LnlOut out{};
    out.nout = 1;
    out.ptr[0] = dl;
    k_lnl_finish<<<(unsigned)((npv + 127) / 128), 128, 0, st>>>(h->d_partial.as<double>(), 1, sigma, h->d_nblk.as<double>(),
                                                                 (int)h->nblocks, (int)npv, out);
    h->launches += 3;
    CU(cudaGetLastError());
    if (!direct) {
        CU(cudaMemcpyAsync(lnl, dl, npv * 8, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
    }
    return PTB_OK;
}

int ptb_set_baseline(ptb_model *h, const double *basis, int64_t nbasis, const int64_t *cstart, const int64_t *ncoef) {
    if (!h) return PTB_EINVAL;
    if (int rc = set_device(h)) return rc;
    h->bl_nbasis = 0;
    h->data_gen++;
    if (!basis) return PTB_OK;
    if (!h->has_data) return fail(h, PTB_ESTATE, "set_baseline: call set_data first");
    if (nbasis < 1 || nbasis > 64 || !cstart || !ncoef) return fail(h, PTB_EINVAL, "set_baseline: 1..64 basis functions and cstart / ncoef per light curve expected");
    std::vector<int64_t> cs(h->nlc), nc(h->nlc);
    CU(cudaMemcpy(cs.data(), cstart, h->nlc * 8, cudaMemcpyDefault));
    CU(cudaMemcpy(nc.data(), ncoef, h->nlc * 8, cudaMemcpyDefault));
    std::vector<int32_t> meta(2 * h->nlc);
    for (int64_t i = 0; i < h->nlc; ++i) {
        if (nc[i] < 0 || nc[i] > nbasis || cs[i] < 0 || cs[i] > (1 << 20))
            return fail(h, PTB_EINVAL, "set_baseline: light curve %lld has ncoef=%lld (nbasis=%lld), cstart=%lld", (long long)i, (long long)nc[i], (long long)nbasis, (long long)cs[i]);
        meta[i] = (int32_t)cs[i];
        meta[h->nlc + i] = (int32_t)nc[i];
    }
    CU(h->d_blmeta.reserve(meta.size() * 4));
    CU(cudaMemcpy(h->d_blmeta.ptr, meta.data(), meta.size() * 4, cudaMemcpyHostToDevice));
    CU(h->d_basis.reserve((size_t)nbasis * h->npt * 8));
    CU(cudaMemcpy(h->d_basis.ptr, basis, (size_t)nbasis * h->npt * 8, cudaMemcpyDefault));
    h->bl_nbasis = nbasis;
    return PTB_OK;
}

int ptb_lpf_flux_model(ptb_model *h, const double *pvp, int64_t npv, const ptb_lpf_layout *lay, int32_t only_baseline,
                       double *flux, void *stream) {
    if (!h) return PTB_EINVAL;
    if (int rc = set_device(h)) return rc;
    if (!flux) return fail(h, PTB_EINVAL, "lpf_flux_model: flux is null");
    if (h->cfg.precision != 0) return fail(h, PTB_ESTATE, "lpf_flux_model: needs an fp64 handle");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    ModelArgs A{};
    const double *sigma = nullptr, *pvp_dev = nullptr;
    if (int rc = lpf_map(h, "lpf_flux_model", pvp, npv, lay, false, st, A, sigma, &pvp_dev)) return rc;
    BaselineParams B{};
    if (int rc = baseline_params(h, "lpf_flux_model", pvp_dev, lay, B)) return rc;
    if (only_baseline && !B.basis) return fail(h, PTB_ESTATE, "lpf_flux_model: no baseline registered");
    const size_t count = (size_t)npv * h->npt;
    const bool direct = is_device_ptr(flux);
    double *dflux = flux;
    if (!direct) {
        CU(h->d_flux.reserve(count * 8));
        dflux = h->d_flux.as<double>();
    }
    if (!only_baseline)
        if (int rc = ptb_rr_evaluate(h, npv, A.k, A.kcols, A.ld, A.nld, nullptr, A.t0, A.p, A.a, A.inc, A.e, A.w, dflux, stream)) return rc;
    if (B.basis) {
        if (npv > 65535) return fail(h, PTB_EINVAL, "lpf_flux_model: at most 65535 parameter vectors per call with a baseline");
        k_lpf_baseline<<<dim3((unsigned)((h->npt + 255) / 256), (unsigned)npv), 256, 0, st>>>(B, dflux, only_baseline ? 1 : 0);
        h->launches++;
        CU(cudaGetLastError());
    }
    h->last_flux_count = direct ? 0 : (int64_t)count;
    if (!direct) return deliver_host(h, flux, dflux, count, 8, st);
    return PTB_OK;
}

}  // extern "C"

#include "ptb_ts_host.inl"
