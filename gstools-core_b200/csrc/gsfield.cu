// gsfield.cu -- C ABI (include/gsfield.h) and host runtime of the B200 field summation.
//
// Replaces the bodies of field::summator / summator_incompr / summator_fourier
// (/root/reference/src/field.rs:37-65, 97-182, 219-249).  No CPU compute path exists here: every
// entry point either runs the CUDA kernels of gsf_kernels.cuh or returns an error code.
//
// Host runtime in one paragraph: a process-wide context holds, per device, three pipeline slots
// (stream + device chunk buffers + pinned staging).  Host-resident points are cut into chunks;
// chunk c goes through slot c%3 as  [gather->pinned] -> H2D -> kernel -> D2H -> [scatter<-pinned]
// so copies of neighbouring chunks overlap the kernel.  Pinned user memory skips the staging
// steps, device-resident memory skips the copies.  With several devices configured the points are
// sharded contiguously, one host thread per device, no collective.
#include <cuda_runtime.h>
#include <sched.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <deque>
#include <memory>
#include <condition_variable>
#include <functional>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/gsfield.h"
#include "gsf_kernels.cuh"
#include "gsf_grid_kernels.cuh"
#include "gsf_krige_kernels.cuh"
#include "gsf_variogram_kernels.cuh"

namespace {

using gsf::kThreads;
using gsf::SumArgs;

thread_local std::string g_err;

int fail(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

// GSF_TRACE=1: per-phase host timestamps of every call on stderr (latency work on small problems)
struct Trace {
    std::chrono::steady_clock::time_point t0, last;
    static bool enabled()
    {
        static const bool e = []() { const char *v = getenv("GSF_TRACE"); return v && v[0] == '1'; }();
        return e;
    }
    void begin()
    {
        if (!enabled()) return;
        t0 = last = std::chrono::steady_clock::now();
        fprintf(stderr, "[gsf-trace] ---- call\n");
    }
    void mark(const char *what)
    {
        if (!enabled()) return;
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[gsf-trace] %-28s +%7.1f us  (%8.1f)\n", what,
                std::chrono::duration<double, std::micro>(now - last).count(),
                std::chrono::duration<double, std::micro>(now - t0).count());
        last = now;
    }
};
thread_local Trace g_trace;

#define GSF_CUDA(call)                                                                     \
    do {                                                                                   \
        cudaError_t e_ = (call);                                                           \
        if (e_ != cudaSuccess)                                                             \
            return fail(e_ == cudaErrorMemoryAllocation ? GSF_ERR_ALLOC : GSF_ERR_CUDA,    \
                        "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__,  \
                        __LINE__);                                                         \
    } while (0)

// ---------------------------------------------------------------------------------------------
// kernel table
typedef void (*SumKernel)(SumArgs);

// Every (P, L) variant exists at the high degree; the throughput degree (gsf::kFastDeg) only for
// L = 1 -- it is chosen for large problems only, and those never split the modes over lanes.
template <int D, int NC, int DEG>
SumKernel pick_pl(int P, int L)
{
#define GSF_V(p, l) \
    if (P == p && L == l) return gsf::gsf_sum_kernel<D, NC, p, l, DEG>;
    GSF_V(4, 1) GSF_V(3, 1) GSF_V(2, 1) GSF_V(1, 1)
    if constexpr (DEG == gsf::kHiDeg) {
        GSF_V(2, 2) GSF_V(2, 4) GSF_V(2, 8) GSF_V(2, 16) GSF_V(2, 32)
        GSF_V(1, 2) GSF_V(1, 4) GSF_V(1, 8) GSF_V(1, 16) GSF_V(1, 32)
    }
#undef GSF_V
    return nullptr;
}

template <int D, int DEG>
SumKernel pick_hi(int P, int L)   // dim 4..8: fewer variants
{
    if (P == 1 && L == 1) return gsf::gsf_sum_kernel<D, 1, 1, 1, DEG>;
    if constexpr (DEG == gsf::kHiDeg) {
        if (P == 1 && L == 4) return gsf::gsf_sum_kernel<D, 1, 1, 4, DEG>;
        if (P == 1 && L == 32) return gsf::gsf_sum_kernel<D, 1, 1, 32, DEG>;
    }
    return nullptr;
}

template <int DEG>
SumKernel pick_kernel_deg(int dim, bool incompr, int P, int L)
{
    if (incompr) {
        if (dim == 2) return pick_pl<2, 2, DEG>(P, L);
        if (dim == 3) return pick_pl<3, 3, DEG>(P, L);
        return nullptr;
    }
    switch (dim) {
        case 1: return pick_pl<1, 1, DEG>(P, L);
        case 2: return pick_pl<2, 1, DEG>(P, L);
        case 3: return pick_pl<3, 1, DEG>(P, L);
        case 4: return pick_hi<4, DEG>(P, L);
        case 5: return pick_hi<5, DEG>(P, L);
        case 6: return pick_hi<6, DEG>(P, L);
        case 7: return pick_hi<7, DEG>(P, L);
        case 8: return pick_hi<8, DEG>(P, L);
    }
    return nullptr;
}

SumKernel pick_kernel(int dim, bool incompr, int P, int L, int deg = gsf::kHiDeg)
{
    if (deg == gsf::kFastDeg && gsf::kFastDeg != gsf::kHiDeg) return pick_kernel_deg<gsf::kFastDeg>(dim, incompr, P, L);
    return pick_kernel_deg<gsf::kHiDeg>(dim, incompr, P, L);
}

// ---------------------------------------------------------------------------------------------
constexpr int kSlots = 3;

struct Slot {
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_h2d = nullptr;    // staging-in buffer consumed
    cudaEvent_t ev_done = nullptr;   // D2H into staging-out finished
    double *d_pos = nullptr, *d_out = nullptr;
    double *h_pos = nullptr, *h_out = nullptr;
    size_t d_pos_cap = 0, d_out_cap = 0, h_pos_cap = 0, h_out_cap = 0;   // in doubles
    double *g_partial = nullptr, *g_counter = nullptr;                   // grid path: split-K workspace
    size_t g_partial_cap = 0, g_counter_cap = 0;
};

struct DeviceCtx {
    int dev = -1;
    int sm_count = 0;
    bool ready = false;
    Slot slot[kSlots];
    cudaEvent_t ev_modes = nullptr;  // records are ready
    cudaEvent_t ev_ws = nullptr;     // last consumer of the mode workspace
    bool ws_used = false;
    double *d_raw = nullptr, *d_rec = nullptr;
    size_t raw_cap = 0, rec_cap = 0;
    // host copy of the raw modes whose records currently sit in d_rec: a call with identical
    // modes (the same random field evaluated at new positions) skips the upload and the pre-pass
    std::vector<double> rec_src;
    int rec_kind = -1, rec_dim = 0;
    int64_t rec_n = -1;
    double rec_scale = 0.0;
    bool rec_valid = false;
    // raw modes of the last one-launch small call (see small_modes_repeat)
    std::vector<double> small_seen;
    int small_kind = -1, small_dim = 0;
    double small_scale = 0.0;
    double *g_axes = nullptr, *g_E0 = nullptr, *g_E1 = nullptr, *g_F = nullptr;   // grid-path tables
    size_t g_axes_cap = 0, g_E0_cap = 0, g_E1_cap = 0, g_F_cap = 0;
    double *ring_in = nullptr, *ring_out = nullptr;   // pinned staging rings of the chunk pipeline
    size_t ring_in_cap = 0, ring_out_cap = 0;
    std::vector<cudaEvent_t> ev_ring_in, ev_ring_out;   // per ring slot: H2D consumed it / D2H filled it
    int staging_threads = 0;             // host threads that staged pageable memory in the last call
    double *g_counter0 = nullptr;        // scratch of the device-side grid detection
    size_t g_counter0_cap = 0;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof;   // pool of timing events
    size_t prof_used = 0;
    cudaEvent_t prep_beg = nullptr, prep_end = nullptr;
    bool prep_timed = false;
    // last-call bookkeeping
    int64_t h2d_bytes = 0, d2h_bytes = 0;
    int launches = 0, chunks = 0;
    int status = GSF_OK;
    std::string err;
};

struct Context {
    std::mutex mu;
    std::vector<int> devices;            // configured device ids
    bool devices_explicit = false;
    std::vector<DeviceCtx *> dctx;       // indexed by device id
    int64_t chunk_points = 0;
    int force_p = 0, force_l = 0;
    int poly_degree = 0;                 // 0: choose_degree's rule; else forced (gsf_set_poly_degree / GSF_POLY_DEGREE)
    int profiling = 0;                   // 0 off, 1 per call, 2 accumulate over calls (gsf_set_profiling)
    int grid_detect = -1;                // -1: env GSF_GRID_DETECT (default on), 0 off, 1 on
    gsf_stats last{};
    std::vector<int> last_devs;
};

Context &ctx()
{
    static Context c;
    return c;
}

int device_count_raw()
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int ensure_cap(double **p, size_t *cap, size_t need, bool pinned)
{
    if (need == 0) need = 1;
    if (*cap >= need) return GSF_OK;
    if (*p) {
        if (pinned) cudaFreeHost(*p); else cudaFree(*p);
        *p = nullptr;
        *cap = 0;
    }
    size_t want = need + need / 8;
    cudaError_t e = pinned ? cudaMallocHost((void **)p, want * sizeof(double))
                           : cudaMalloc((void **)p, want * sizeof(double));
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(GSF_ERR_ALLOC, "%s of %zu bytes failed: %s", pinned ? "cudaMallocHost" : "cudaMalloc",
                    want * sizeof(double), cudaGetErrorString(e));
    }
    *cap = want;
    return GSF_OK;
}

int get_device_ctx(int dev, DeviceCtx **out)
{
    Context &c = ctx();
    if (dev < 0) return fail(GSF_ERR_ARG, "bad device id %d", dev);
    if ((size_t)dev >= c.dctx.size()) c.dctx.resize(dev + 1, nullptr);
    if (!c.dctx[dev]) c.dctx[dev] = new DeviceCtx();
    DeviceCtx *d = c.dctx[dev];
    if (!d->ready) {
        GSF_CUDA(cudaSetDevice(dev));
        d->dev = dev;
        cudaDeviceProp prop;
        GSF_CUDA(cudaGetDeviceProperties(&prop, dev));
        d->sm_count = prop.multiProcessorCount;
        // the library carries sm_100a SASS only (no PTX): any other architecture -- older or newer --
        // would fail at the first launch with "no kernel image"; say so here instead
        cudaFuncAttributes fa;
        if (prop.major != 10 || cudaFuncGetAttributes(&fa, gsf::gsf_prep_modes) != cudaSuccess) {
            cudaGetLastError();
            return fail(GSF_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a (B200) only",
                        dev, prop.major, prop.minor);
        }
        for (int s = 0; s < kSlots; ++s) {
            GSF_CUDA(cudaStreamCreateWithFlags(&d->slot[s].stream, cudaStreamNonBlocking));
            GSF_CUDA(cudaEventCreateWithFlags(&d->slot[s].ev_h2d, cudaEventDisableTiming));
            GSF_CUDA(cudaEventCreateWithFlags(&d->slot[s].ev_done, cudaEventDisableTiming));
        }
        GSF_CUDA(cudaEventCreateWithFlags(&d->ev_modes, cudaEventDisableTiming));
        GSF_CUDA(cudaEventCreateWithFlags(&d->ev_ws, cudaEventDisableTiming));
        GSF_CUDA(cudaEventCreate(&d->prep_beg));
        GSF_CUDA(cudaEventCreate(&d->prep_end));
        d->ready = true;
    }
    *out = d;
    return GSF_OK;
}

void free_device_ctx(DeviceCtx *d)
{
    if (!d) return;
    if (d->ready) {
        cudaSetDevice(d->dev);
        cudaDeviceSynchronize();
        for (int s = 0; s < kSlots; ++s) {
            Slot &sl = d->slot[s];
            if (sl.d_pos) cudaFree(sl.d_pos);
            if (sl.d_out) cudaFree(sl.d_out);
            if (sl.h_pos) cudaFreeHost(sl.h_pos);
            if (sl.h_out) cudaFreeHost(sl.h_out);
            if (sl.g_partial) cudaFree(sl.g_partial);
            if (sl.g_counter) cudaFree(sl.g_counter);
            if (sl.ev_h2d) cudaEventDestroy(sl.ev_h2d);
            if (sl.ev_done) cudaEventDestroy(sl.ev_done);
            if (sl.stream) cudaStreamDestroy(sl.stream);
        }
        if (d->d_raw) cudaFree(d->d_raw);
        if (d->d_rec) cudaFree(d->d_rec);
        if (d->g_axes) cudaFree(d->g_axes);
        if (d->g_E0) cudaFree(d->g_E0);
        if (d->g_E1) cudaFree(d->g_E1);
        if (d->g_F) cudaFree(d->g_F);
        if (d->g_counter0) cudaFree(d->g_counter0);
        if (d->ring_in) cudaFreeHost(d->ring_in);
        if (d->ring_out) cudaFreeHost(d->ring_out);
        for (cudaEvent_t e : d->ev_ring_in) cudaEventDestroy(e);
        for (cudaEvent_t e : d->ev_ring_out) cudaEventDestroy(e);
        for (auto &pr : d->prof) {
            cudaEventDestroy(pr.first);
            cudaEventDestroy(pr.second);
        }
        cudaEventDestroy(d->ev_modes);
        cudaEventDestroy(d->ev_ws);
        cudaEventDestroy(d->prep_beg);
        cudaEventDestroy(d->prep_end);
    }
    delete d;
}

int ensure_ring_events(DeviceCtx &d, int depth)
{
    while ((int)d.ev_ring_in.size() < depth) {
        cudaEvent_t e;
        GSF_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        d.ev_ring_in.push_back(e);
    }
    while ((int)d.ev_ring_out.size() < depth) {
        cudaEvent_t e;
        GSF_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        d.ev_ring_out.push_back(e);
    }
    return GSF_OK;
}

// memory kind of a user pointer: 0 pageable host, 1 pinned host, 2 device/managed
int classify(const void *p, int *kind, int *device)
{
    cudaPointerAttributes at;
    cudaError_t e = cudaPointerGetAttributes(&at, p);
    if (e != cudaSuccess) {
        cudaGetLastError();
        *kind = 0;
        *device = -1;
        return GSF_OK;
    }
    switch (at.type) {
        case cudaMemoryTypeHost: *kind = 1; *device = -1; break;
        case cudaMemoryTypeDevice:
        case cudaMemoryTypeManaged: *kind = 2; *device = at.device; break;
        default: *kind = 0; *device = -1; break;
    }
    return GSF_OK;
}

// ---------------------------------------------------------------------------------------------
struct Problem {
    int kind;            // gsf::Kind
    int dim;
    int64_t N, M;
    const double *sf; int64_t sfs;
    const double *k;  int64_t ks0, ks1;
    const double *z1; int64_t z1s;
    const double *z2; int64_t z2s;
    const double *pos; int64_t ps0, ps1;
    double *out; int64_t os0, os1;
    int threads_hint = 0;
    int peer_owner = -1;                 // >= 0: device arrays live on this device, peers may read them
    bool zero_copy = false;              // pos/out are mapped pinned host memory read over PCIe
    int deg = gsf::kHiDeg;               // degree of the cosine polynomial, ONE per call (choose_degree)
    double scale = 1.0;                  // out = scale * sum + offset[a]   
    double offset[3] = {0.0, 0.0, 0.0};
    // memory kind of the mode arrays (classify): filled once per call by run_host_call, -1 = not yet known
    int mem_k = -1, mem_k_dev = -1, mem_z1 = -1, mem_z2 = -1, mem_sf = -1;
    int64_t mode_group = 1;              // structured-grid path: consecutive modes sharing all but the last wave-vector component
    int nc() const { return kind == gsf::kIncompr ? dim : 1; }
    int rec() const { return gsf::rec_doubles(dim, nc()); }
};

int validate(const Problem &p)
{
    if (p.N < 0 || p.M < 0) return fail(GSF_ERR_SHAPE, "negative size: n_modes=%lld n_points=%lld",
                                        (long long)p.N, (long long)p.M);
    if (p.kind == gsf::kIncompr) {
        if (p.dim != 2 && p.dim != 3)
            return fail(GSF_ERR_DIM, "Only two- and three-dimensional problems are supported. (dim=%d)", p.dim);
        if (p.N == 0) return fail(GSF_ERR_EMPTY_MODES, "summate_incompr needs at least one mode");
    } else if (p.dim < 1) {
        // any dim >= 1 works, as in the reference: 1..GSF_MAX_DIM on the tuned templates, larger
        // ones on gsf_sum_kernel_anyd
        return fail(GSF_ERR_DIM, "dim=%d: need at least one spatial dimension", p.dim);
    }
    if (p.N > 0 && (!p.k || !p.z1 || !p.z2 || (p.kind == gsf::kFourier && !p.sf)))
        return fail(GSF_ERR_ARG, "NULL mode array");
    if (p.M > 0 && (!p.pos || !p.out)) return fail(GSF_ERR_ARG, "NULL pos/out");
    return GSF_OK;
}

// Choose points-per-thread P and lanes-per-point L for a launch of m_launch points.
// Table from tools/variant_sweep.py on B200 (N = 1000, all kinds/dims):
//   * below ~750k points P = 1 wins (most CTAs per SM => best balance across the 148 SMs);
//   * above, P = 3 amortises the LDS/loop instructions that steal issue cycles from the FP64
//     pipe, with the last half resident wave in short P = 1 tiles (launch_sum);
//   * fewer than 2 CTAs per SM: split the modes over L lanes of a point group so that every SM
//     gets work (keep >= 16 modes per lane).
void choose_variant(const DeviceCtx &d, const Problem &p, int64_t m_launch, bool pipelined, int *P, int *L)
{
    Context &c = ctx();
    const bool inc = p.kind == gsf::kIncompr;
    if (c.force_p > 0 && c.force_l > 0 && pick_kernel(p.dim, inc, c.force_p, c.force_l, p.deg)) {
        *P = c.force_p;
        *L = c.force_l;
        return;
    }
    const int64_t ctas1 = (m_launch + kThreads - 1) / kThreads;   // P = 1, L = 1
    if (p.dim > gsf::kMaxTemplateDim) {   // gsf_sum_kernel_anyd: one point per thread
        *P = 1;
        *L = 1;
        return;
    }
    if (p.dim > 3) {   // dims 4..8 ship P = 1 and L in {1, 4, 32}
        *P = 1;
        *L = ctas1 >= 2 * d.sm_count ? 1 : (ctas1 * 4 >= 2 * d.sm_count || p.N < 512 ? 4 : 32);
        if (p.N < 16 * *L) *L = 1;
        return;
    }
    int bestP = 1, bestL = 1;
    if (p.zero_copy && p.N < 4000 && m_launch >= 750000) {   // see run_host_call: short tiles start the
        *P = 1;                                              // machine sooner (instead of P = 3)
        *L = 1;
        return;
    }
    // a pipelined chunk is followed by the next chunk's kernel on another stream, which fills the
    // SMs as this one drains: long tiles pay off from much smaller launches
    if (m_launch >= (pipelined ? 120000 : 750000)) {
        // (3-D incompressible: P = 4 measured +1.2 % over P = 3 with the degree-5 polynomial,
        //  tools/variant_p_probe.py; every other kind is best at 3)
        bestP = inc && p.dim == 3 && p.deg == gsf::kFastDeg ? 4 : 3;
    } else if (ctas1 < 2 * d.sm_count && !pipelined) {
        // (pipelined chunks never split the modes: neighbouring chunks' kernels fill the machine,
        //  and L = 1 keeps every point's summation order independent of chunking / device count)
        while (bestL < 32 && ctas1 * bestL < 2 * d.sm_count && p.N >= 32 * bestL) bestL *= 2;
    }
    if (p.deg != gsf::kHiDeg) bestL = 1;   // the throughput degree ships L = 1 variants only
    *P = bestP;
    *L = bestL;
}

// Degree of the cosine polynomial for one call (gsf_kernels.cuh): the throughput degree when the
// FP64 pipe is what the caller waits for, the high degree for small problems where launch latency
// dominates and the extra DFMA is free.  Decided ONCE per call from the whole problem -- never per
// chunk or per device shard -- so that results do not depend on chunking or on the device count.
int choose_degree(const Problem &p)
{
    Context &c = ctx();
    static const int env_deg = []() { const char *e = getenv("GSF_POLY_DEGREE"); return e && *e ? atoi(e) : 0; }();
    const int forced = c.poly_degree ? c.poly_degree : env_deg;
    const bool lanes_forced = c.force_p > 0 && c.force_l > 1;
    if (forced == gsf::kHiDeg || forced == gsf::kFastDeg)
        return lanes_forced || p.dim > gsf::kMaxTemplateDim ? gsf::kHiDeg : forced;
    if (lanes_forced || p.dim > gsf::kMaxTemplateDim) return gsf::kHiDeg;
    const double pm = (double)p.N * (double)p.M;
    return pm >= 134217728.0 && p.M >= 65536 ? gsf::kFastDeg : gsf::kHiDeg;   // 2^27 point*modes ~ 0.1 ms of kernel
}

int launch_sum(DeviceCtx &d, const Problem &p, const double *kpos, int64_t ps0, int64_t ps1, double *kout,
               int64_t os0, int64_t os1, int64_t m, cudaStream_t st, int P, int L)
{
    const bool anyd = p.dim > gsf::kMaxTemplateDim;
    if (anyd) { P = 1; L = 1; }
    SumKernel fn = anyd ? gsf::gsf_sum_kernel_anyd<gsf::kHiDeg> : pick_kernel(p.dim, p.kind == gsf::kIncompr, P, L, p.deg);
    if (!fn) return fail(GSF_ERR_ARG, "no kernel variant dim=%d P=%d L=%d degree=%d", p.dim, P, L, p.deg);
    SumArgs a;
    a.dim = p.dim;
    a.rec = d.d_rec;
    a.n_modes = p.N;
    a.pos = kpos; a.ps0 = ps0; a.ps1 = ps1;
    a.n_points = m;
    a.out = kout; a.os0 = os0; a.os1 = os1;
    for (int c = 0; c < 3; ++c) a.offset[c] = p.offset[c];
    gsf::poly_constants(anyd ? gsf::kHiDeg : p.deg, a.coef);
    // long tiles (P points per thread) first, then short tiles (1 point per thread) for the last
    // `tail` resident waves so that the machine drains in small steps (see gsf_sum_kernel)
    const int64_t tile = (int64_t)P * (kThreads / L);
    int64_t n_big = (m + tile - 1) / tile, n_small = 0;
    if (P > gsf::kTailP && L == 1) {
        static const double tail_waves = []() {
            const char *e = getenv("GSF_TAIL_WAVES");
            return e && *e ? atof(e) : 0.5;
        }();
        const int64_t resident = (int64_t)d.sm_count * 8;                  // ~CTAs in flight
        int64_t tail_pts = (int64_t)(tail_waves * (double)resident * (double)tile);
        tail_pts = std::min(tail_pts, m);
        n_big = (m - tail_pts) / tile;
        const int64_t small_tile = (int64_t)kThreads * gsf::kTailP;
        n_small = (m - n_big * tile + small_tile - 1) / small_tile;
    }
    a.n_big = n_big;
    const int64_t grid = n_big + n_small;
    if (grid > 0x7fffffffLL) return fail(GSF_ERR_SHAPE, "chunk of %lld points is too large for one launch", (long long)m);
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (ctx().profiling) {
        if (d.prof_used == d.prof.size()) {
            cudaEvent_t a0, a1;
            GSF_CUDA(cudaEventCreate(&a0));
            GSF_CUDA(cudaEventCreate(&a1));
            d.prof.emplace_back(a0, a1);
        }
        e0 = d.prof[d.prof_used].first;
        e1 = d.prof[d.prof_used].second;
        d.prof_used++;
        GSF_CUDA(cudaEventRecord(e0, st));
    }
    fn<<<(unsigned)grid, kThreads, 0, st>>>(a);
    GSF_CUDA(cudaGetLastError());
    if (e1) GSF_CUDA(cudaEventRecord(e1, st));
    d.launches++;
    return GSF_OK;
}

// Upload (if host) and pre-process the modes on `st`.  Host arrays are gathered into a contiguous
// temporary and copied with a pageable cudaMemcpyAsync (staged by the driver before it returns).
// `amp_factor`: what the consumer of the records wants folded into the amplitudes besides p.scale
// (gsf::amp_factor(p.deg) for the point x mode kernels, 1 for the structured-grid tables).
int prepare_modes(DeviceCtx &d, const Problem &p, cudaStream_t st, double amp_factor)
{
    const int64_t N = p.N;
    const double scale = amp_factor == 1.0 ? p.scale : p.scale * amp_factor;
    int rc;
    {
        const size_t cap_before = d.rec_cap;
        if ((rc = ensure_cap(&d.d_rec, &d.rec_cap, (size_t)N * p.rec(), false))) return rc;
        if (d.rec_cap != cap_before) d.rec_valid = false;
    }
    if (N == 0) return GSF_OK;
    int kk = p.mem_k, kd = p.mem_k_dev, tmp;
    if (kk < 0) classify(p.k, &kk, &kd);
    int k1 = p.mem_z1, k2 = p.mem_z2, k3 = kk;
    if (k1 < 0) classify(p.z1, &k1, &tmp);
    if (k2 < 0) classify(p.z2, &k2, &tmp);
    if (p.kind == gsf::kFourier) {
        k3 = p.mem_sf;
        if (k3 < 0) classify(p.sf, &k3, &tmp);
    }
    const bool all_dev = kk == 2 && k1 == 2 && k2 == 2 && k3 == 2;
    const bool all_host = kk != 2 && k1 != 2 && k2 != 2 && k3 != 2;
    if (!all_dev && !all_host)
        return fail(GSF_ERR_ARG, "mode arrays must be all host-resident or all device-resident");

    gsf::PrepArgs a;
    a.n_modes = N;
    a.dim = p.dim;
    a.incompr = p.kind == gsf::kIncompr;
    a.scale = scale;
    a.rec = d.d_rec;
    if (all_dev) {
        if (kd != d.dev && kd != p.peer_owner)
            return fail(GSF_ERR_ARG, "mode arrays live on device %d, work runs on device %d", kd, d.dev);
        a.k = p.k; a.ks0 = p.ks0; a.ks1 = p.ks1;
        a.z1 = p.z1; a.z1s = p.z1s;
        a.z2 = p.z2; a.z2s = p.z2s;
        a.sf = p.kind == gsf::kFourier ? p.sf : nullptr; a.sfs = p.sfs;
        d.rec_valid = false;
    } else {
        const int rows = p.dim + 2 + (p.kind == gsf::kFourier ? 1 : 0);
        if ((rc = ensure_cap(&d.d_raw, &d.raw_cap, (size_t)rows * N, false))) return rc;
        static thread_local std::vector<double> h;
        h.resize((size_t)rows * N);
        for (int dd = 0; dd < p.dim; ++dd)
            for (int64_t i = 0; i < N; ++i) h[(size_t)dd * N + i] = p.k[dd * p.ks0 + i * p.ks1];
        for (int64_t i = 0; i < N; ++i) h[(size_t)p.dim * N + i] = p.z1[i * p.z1s];
        for (int64_t i = 0; i < N; ++i) h[(size_t)(p.dim + 1) * N + i] = p.z2[i * p.z2s];
        if (p.kind == gsf::kFourier)
            for (int64_t i = 0; i < N; ++i) h[(size_t)(p.dim + 2) * N + i] = p.sf[i * p.sfs];
        if (d.rec_valid && d.rec_kind == p.kind && d.rec_dim == p.dim && d.rec_n == N && d.rec_scale == scale &&
            d.rec_src.size() == h.size() && memcmp(d.rec_src.data(), h.data(), h.size() * sizeof(double)) == 0)
            return GSF_OK;   // d_rec already holds exactly these modes
        d.rec_src = h;
        d.rec_kind = p.kind; d.rec_dim = p.dim; d.rec_n = N; d.rec_scale = scale;
        d.rec_valid = true;
        GSF_CUDA(cudaMemcpyAsync(d.d_raw, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, st));
        d.h2d_bytes += (int64_t)(h.size() * sizeof(double));
        a.k = d.d_raw; a.ks0 = N; a.ks1 = 1;
        a.z1 = d.d_raw + (size_t)p.dim * N; a.z1s = 1;
        a.z2 = d.d_raw + (size_t)(p.dim + 1) * N; a.z2s = 1;
        a.sf = p.kind == gsf::kFourier ? d.d_raw + (size_t)(p.dim + 2) * N : nullptr; a.sfs = 1;
    }
    d.prep_timed = ctx().profiling != 0;
    if (d.prep_timed) GSF_CUDA(cudaEventRecord(d.prep_beg, st));
    gsf::gsf_prep_modes<<<(unsigned)((N + 127) / 128), 128, 0, st>>>(a);
    GSF_CUDA(cudaGetLastError());
    if (d.prep_timed) GSF_CUDA(cudaEventRecord(d.prep_end, st));
    d.launches++;
    return GSF_OK;
}

void reset_call_counters(DeviceCtx &d)
{
    d.h2d_bytes = d.d2h_bytes = 0;
    d.launches = d.chunks = 0;
    d.staging_threads = 0;
    if (ctx().profiling != 2) d.prof_used = 0;
    d.prep_timed = false;
    d.status = GSF_OK;
    d.err.clear();
}

// On an early (error) return work may still be in flight on the pipeline streams while the caller
// goes on to free or reuse its buffers: drain the device's streams and drop cached state.
struct DrainOnError {
    DeviceCtx &d;
    bool armed = true;
    explicit DrainOnError(DeviceCtx &dc) : d(dc) {}
    ~DrainOnError()
    {
        if (!armed) return;
        const std::string keep = g_err;
        for (int s = 0; s < kSlots; ++s)
            if (d.slot[s].stream) cudaStreamSynchronize(d.slot[s].stream);
        cudaGetLastError();
        d.rec_valid = false;
        d.ws_used = false;
        for (int s = 0; s < kSlots; ++s) d.slot[s].g_counter_cap = 0;   // tile tickets may be dirty: re-zero
        g_err = keep;
    }
};

#include "gsf_host_staging.inc"

// ---------------------------------------------------------------------------------------------
// One-launch path for small host-resident problems (gsf_small_kernel): raw modes inside the kernel
// parameters, positions gathered into a pinned buffer the kernel reads in place, result written to a
// pinned buffer and scattered on return.  One launch, one stream sync, no copy-engine work.
constexpr int kSmallCapA = 480, kSmallCapB = 1536;

template <int CAP>
int launch_small(DeviceCtx &d, const Problem &p, const double *hpos, double *hout, int64_t os0, int64_t os1, int L,
                 cudaStream_t st)
{
    typedef void (*Fn)(const gsf::SmallArgs<CAP>);
    Fn fn = nullptr;
    const bool inc = p.kind == gsf::kIncompr;
    if (!inc && p.dim == 1) fn = gsf::gsf_small_kernel<1, 1, CAP>;
    else if (!inc && p.dim == 2) fn = gsf::gsf_small_kernel<2, 1, CAP>;
    else if (!inc && p.dim == 3) fn = gsf::gsf_small_kernel<3, 1, CAP>;
    else if (inc && p.dim == 2) fn = gsf::gsf_small_kernel<2, 2, CAP>;
    else if (inc && p.dim == 3) fn = gsf::gsf_small_kernel<3, 3, CAP>;
    if (!fn) return fail(GSF_ERR_ARG, "no small-problem kernel for dim=%d", p.dim);
    static thread_local gsf::SmallArgs<CAP> s;   // 4 / 12 KB: kept off the stack, rebuilt every call
    const int64_t N = p.N;
    s.a = SumArgs{};
    s.a.dim = p.dim;
    s.a.n_modes = N;
    s.a.pos = hpos; s.a.ps0 = p.M; s.a.ps1 = 1;
    s.a.n_points = p.M;
    s.a.out = hout; s.a.os0 = os0; s.a.os1 = os1;
    for (int c = 0; c < 3; ++c) s.a.offset[c] = p.offset[c];
    gsf::poly_constants(gsf::kHiDeg, s.a.coef);
    s.scale = p.scale * gsf::amp_factor(gsf::kHiDeg);
    s.lanes = L;
    s.has_sf = p.kind == gsf::kFourier;
    double *r = s.raw;
    for (int a = 0; a < p.dim; ++a)
        for (int64_t i = 0; i < N; ++i) *r++ = p.k[a * p.ks0 + i * p.ks1];
    for (int64_t i = 0; i < N; ++i) *r++ = p.z1[i * p.z1s];
    for (int64_t i = 0; i < N; ++i) *r++ = p.z2[i * p.z2s];
    if (s.has_sf)
        for (int64_t i = 0; i < N; ++i) *r++ = p.sf[i * p.sfs];
    const int64_t per_cta = kThreads / L;
    const int64_t grid = (p.M + per_cta - 1) / per_cta;
    cudaEvent_t e1 = nullptr;
    if (ctx().profiling) {
        if (d.prof_used == d.prof.size()) {
            cudaEvent_t a0, a1;
            GSF_CUDA(cudaEventCreate(&a0));
            GSF_CUDA(cudaEventCreate(&a1));
            d.prof.emplace_back(a0, a1);
        }
        GSF_CUDA(cudaEventRecord(d.prof[d.prof_used].first, st));
        e1 = d.prof[d.prof_used].second;
        d.prof_used++;
    }
    fn<<<(unsigned)grid, kThreads, 0, st>>>(s);
    GSF_CUDA(cudaGetLastError());
    if (e1) GSF_CUDA(cudaEventRecord(e1, st));
    d.launches++;
    d.h2d_bytes += (int64_t)((r - s.raw) * sizeof(double));
    return GSF_OK;
}

// raw-mode doubles the fused kernel needs for problem p; 0 if p does not qualify
int64_t small_fused_rows(const Problem &p)
{
    static const bool enabled = []() { const char *e = getenv("GSF_SMALL_FUSED"); return !(e && e[0] == '0'); }();
    const Context &c = ctx();
    if (!enabled || p.N < 1 || p.N > gsf::kModeBlock || p.dim > 3 || p.deg != gsf::kHiDeg) return 0;
    if (c.force_p > 0 || c.force_l > 0) return 0;                 // a forced (P, L) variant means the general kernels
    if (p.mem_k != 0 && p.mem_k != 1) return 0;                   // host-resident modes only
    if ((p.mem_z1 != 0 && p.mem_z1 != 1) || (p.mem_z2 != 0 && p.mem_z2 != 1)) return 0;
    if (p.kind == gsf::kFourier && p.mem_sf != 0 && p.mem_sf != 1) return 0;
    const int64_t rows = (p.dim + 2 + (p.kind == gsf::kFourier ? 1 : 0)) * p.N;
    return rows <= kSmallCapB ? rows : 0;
}

// The one-launch kernel rebuilds the mode records in every CTA, which is the cheapest thing to do
// for modes seen once (ensembles: fresh z1/z2 per call).  A caller that evaluates the SAME modes
// again (one field, new positions) is better served by records cached on the device and the plain
// kernel (gsf_prep_modes only once).  So: modes identical to the cached records, or seen for the
// second time in a row, take the cached-record path; anything else stays on the one-launch kernel.
bool small_modes_repeat(DeviceCtx &d, const Problem &p)
{
    const int rows = p.dim + 2 + (p.kind == gsf::kFourier ? 1 : 0);
    const int64_t N = p.N;
    static thread_local std::vector<double> h;
    h.resize((size_t)rows * N);
    for (int dd = 0; dd < p.dim; ++dd)
        for (int64_t i = 0; i < N; ++i) h[(size_t)dd * N + i] = p.k[dd * p.ks0 + i * p.ks1];
    for (int64_t i = 0; i < N; ++i) h[(size_t)p.dim * N + i] = p.z1[i * p.z1s];
    for (int64_t i = 0; i < N; ++i) h[(size_t)(p.dim + 1) * N + i] = p.z2[i * p.z2s];
    if (p.kind == gsf::kFourier)
        for (int64_t i = 0; i < N; ++i) h[(size_t)(p.dim + 2) * N + i] = p.sf[i * p.sfs];
    const double af = gsf::amp_factor(p.deg);
    const double scale = af == 1.0 ? p.scale : p.scale * af;
    const size_t bytes = h.size() * sizeof(double);
    if (d.rec_valid && d.rec_kind == p.kind && d.rec_dim == p.dim && d.rec_n == N && d.rec_scale == scale &&
        d.rec_src.size() == h.size() && memcmp(d.rec_src.data(), h.data(), bytes) == 0)
        return true;
    if (d.small_kind == p.kind && d.small_dim == p.dim && d.small_scale == scale && d.small_seen.size() == h.size() &&
        memcmp(d.small_seen.data(), h.data(), bytes) == 0)
        return true;
    d.small_seen = h;
    d.small_kind = p.kind;
    d.small_dim = p.dim;
    d.small_scale = scale;
    return false;
}

// One launch over device-visible memory (q.pos / q.out are device or mapped host pointers), then wait:
// run_shard's single-chunk case without the chunk schedule, the staging crew and its bookkeeping
// (small calls count microseconds).
int run_single_launch(DeviceCtx &d, const Problem &q, int *P_used, int *L_used)
{
    GSF_CUDA(cudaSetDevice(d.dev));
    reset_call_counters(d);
    DrainOnError guard(d);
    cudaStream_t s0 = d.slot[0].stream;
    int rc;
    if (d.ws_used) GSF_CUDA(cudaStreamWaitEvent(s0, d.ev_ws, 0));
    if ((rc = prepare_modes(d, q, s0, gsf::amp_factor(q.deg)))) return rc;
    g_trace.mark("single: prepare modes");
    choose_variant(d, q, q.M, false, P_used, L_used);
    if ((rc = launch_sum(d, q, q.pos, q.ps0, q.ps1, q.out, q.os0, q.os1, q.M, s0, *P_used, *L_used))) return rc;
    g_trace.mark("single: launched");
    GSF_CUDA(cudaStreamSynchronize(s0));
    g_trace.mark("single: synced");
    d.chunks = 1;
    d.ws_used = false;
    guard.armed = false;
    return GSF_OK;
}

int run_small_fused(DeviceCtx &d, const Problem &p, int64_t rows, int *L_used)
{
    GSF_CUDA(cudaSetDevice(d.dev));
    reset_call_counters(d);
    Slot &sl = d.slot[0];
    const int nc = p.nc();
    int rc;
    if ((rc = ensure_cap(&sl.h_pos, &sl.h_pos_cap, (size_t)p.dim * p.M, true))) return rc;
    if ((rc = ensure_cap(&sl.h_out, &sl.h_out_cap, (size_t)nc * p.M, true))) return rc;
    OutLayout lay;
    lay.aos = nc > 1 && p.os0 == 1 && p.os1 == nc;
    lay.direct = false;
    gather_pos_part(p, 0, p.M, 0, p.M, sl.h_pos, false);   // streaming stores: the GPU read this buffer last (3 us less than memcpy, tools/micro/launch_floor.cu)
    g_trace.mark("small: gather pos");
    // lanes per point: as choose_variant -- widen until every SM has two CTAs, >= 32 modes per lane
    int L = 1;
    const int64_t ctas1 = (p.M + kThreads - 1) / kThreads;
    while (L < 32 && ctas1 * L < 2 * d.sm_count && p.N >= 32 * L) L *= 2;
    *L_used = L;
    // (cudaMallocHost memory: under UVA the device address equals the host address)
    const int64_t os0 = nc == 1 ? 0 : (lay.aos ? 1 : p.M), os1 = nc > 1 && lay.aos ? nc : 1;
    rc = rows <= kSmallCapA ? launch_small<kSmallCapA>(d, p, sl.h_pos, sl.h_out, os0, os1, L, sl.stream)
                            : launch_small<kSmallCapB>(d, p, sl.h_pos, sl.h_out, os0, os1, L, sl.stream);
    g_trace.mark("small: launched");
    const cudaError_t e = cudaStreamSynchronize(sl.stream);   // also on a failed launch: nothing may stay in flight
    if (rc) return rc;
    GSF_CUDA(e);
    g_trace.mark("small: synced");
    scatter_out_part(p, lay, 0, p.M, 0, p.M, sl.h_out, true);
    g_trace.mark("small: scatter out");
    d.h2d_bytes += (int64_t)p.dim * p.M * 8;
    d.d2h_bytes += (int64_t)nc * p.M * 8;
    d.chunks = 1;
    return GSF_OK;
}

// Chunk sizes for streaming m points through the pipeline slots (see run_shard).
std::vector<int64_t> chunk_schedule(int64_t m, bool single_launch, int64_t forced_chunk)
{
    std::vector<int64_t> sizes;
    if (m <= 0) return sizes;
    if (single_launch) {
        sizes.push_back(m);
    } else if (forced_chunk > 0) {
        const int64_t chunk = std::min((forced_chunk + 1023) / 1024 * 1024, m);
        for (int64_t done = 0; done < m; done += chunk) sizes.push_back(std::min(chunk, m - done));
    } else {
        // knobs (tools/chunk_schedule_probe.py): first chunk, growth factor of the ramps, cap as a
        // fraction of the shard and in absolute points, number of down-ramp chunks
        static const struct Knobs {
            int64_t lo, cap_abs; int growth, cap_div, down_max;
            static int64_t env(const char *n, int64_t dflt) { const char *e = getenv(n); return e && *e ? atoll(e) : dflt; }
            Knobs() : lo(env("GSF_CHUNK_FIRST", 1 << 15)), cap_abs(env("GSF_CHUNK_CAP", 1 << 20)),
                      growth((int)env("GSF_CHUNK_GROWTH", 2)), cap_div((int)env("GSF_CHUNK_CAP_DIV", 6)),
                      down_max((int)env("GSF_CHUNK_DOWN", 64)) {}
        } kn;
        // first chunk: a quarter of a mid-sized problem, so that its kernel starts after a short copy
        // (tools/e2e_size_sweep.py: 188 -> 164 us at 6e4 points x 1000 modes), never more than kn.lo
        const int64_t lo = std::max<int64_t>(1024, (m <= kn.lo ? kn.lo : std::min<int64_t>(kn.lo, std::max<int64_t>(8192, m / 4))) / 1024 * 1024);
        const int growth = std::max(2, kn.growth);
        int64_t cap = std::min<int64_t>(kn.cap_abs, std::max<int64_t>(1 << 16, m / std::max(1, kn.cap_div)));
        cap = cap / 1024 * 1024;
        std::vector<int64_t> up, down;
        int64_t used = 0;
        for (int64_t c = lo; c < cap && used + c + c / 2 <= m / 2; c *= growth) { up.push_back(c); used += c; }
        for (int64_t c = lo; c < cap && used + c + c / 2 <= m * 3 / 4 && (int)down.size() < kn.down_max; c *= growth) {
            down.push_back(c);
            used += c;
        }
        int64_t rest = m - used;
        sizes = up;
        while (rest > 0) {
            int64_t c = std::min(cap, rest);
            if (rest - c > 0 && rest - c < lo) c = rest;          // do not leave a tiny remainder
            sizes.push_back(c);
            rest -= c;
        }
        for (size_t i = down.size(); i-- > 0;) sizes.push_back(down[i]);
    }
    return sizes;
}

// Process points [j_beg, j_end) of the problem on device d (host- or device-resident pos/out).
//
// Host-resident points stream through the device in chunks.  Chunk c runs on stream c % 3 with that
// stream's device buffers:  H2D -> kernel -> D2H, so the copies of neighbouring chunks overlap the
// kernel on the two copy engines.  Pageable memory goes through pinned rings (`ring_in` / `ring_out`
// slots, deeper than the stream count so that staging can run ahead of the GPU) filled and drained
// by the Stager crew (gsf_host_staging.inc); this thread only waits for "chunk c staged", issues,
// and publishes event completions.  Pinned user memory skips the rings, device memory the copies.
int run_shard(DeviceCtx &d, const Problem &p, int64_t j_beg, int64_t j_end, int pos_kind, int out_kind,
              int *P_used, int *L_used, int host_threads)
{
    GSF_CUDA(cudaSetDevice(d.dev));
    reset_call_counters(d);
    DrainOnError guard(d);
    const int nc = p.nc();
    const int64_t m_shard = j_end - j_beg;
    int rc;

    cudaStream_t s0 = d.slot[0].stream;
    if (d.ws_used) GSF_CUDA(cudaStreamWaitEvent(s0, d.ev_ws, 0));
    if ((rc = prepare_modes(d, p, s0, gsf::amp_factor(p.deg)))) return rc;
    g_trace.mark("shard: prepare modes");

    const bool pos_dev = pos_kind == 2, out_dev = out_kind == 2;
    // Chunk schedule.  Device-resident both ways => one launch.  Otherwise the first chunks are small
    // and double in size (the first kernel starts after a ~0.8 MB copy instead of waiting for
    // megabytes), the middle runs at `cap` points per chunk (long kernels, P = 3 tiles), and the last
    // chunks halve again so that little D2H is left exposed after the final kernel.
    // gsf_set_chunk_points(n) forces a fixed size.
    std::vector<int64_t> sizes = chunk_schedule(m_shard, pos_dev && out_dev, ctx().chunk_points);
    const int64_t n_chunks = (int64_t)sizes.size();
    int64_t chunk = 0;
    std::vector<int64_t> starts;
    int64_t j_run = j_beg;
    for (int64_t c : sizes) {
        starts.push_back(j_run);
        j_run += c;
        chunk = std::max(chunk, c);
    }
    if (n_chunks > 1) {   // (a single launch runs on s0 itself)
        GSF_CUDA(cudaEventRecord(d.ev_modes, s0));
        for (int s = 1; s < kSlots; ++s) GSF_CUDA(cudaStreamWaitEvent(d.slot[s].stream, d.ev_modes, 0));
    }

    // GSF_PAGEABLE_DIRECT=1: hand pageable memory straight to cudaMemcpyAsync (the driver stages it)
    static const bool pageable_direct = []() { const char *e = getenv("GSF_PAGEABLE_DIRECT"); return e && e[0] == '1'; }();
    const bool direct_in = (pos_kind == 1 || (pageable_direct && pos_kind == 0)) && p.ps1 == 1;
    OutLayout lay;
    lay.aos = nc > 1 && p.os0 == 1 && p.os1 == nc;
    lay.direct = out_kind == 1 && (nc == 1 ? p.os1 == 1 : (lay.aos || p.os1 == 1));

    int P = 1, L = 1;
    const bool pipelined = n_chunks > 1;
    choose_variant(d, p, std::min(chunk, m_shard), pipelined, &P, &L);
    *P_used = P;
    *L_used = L;

    // ---- staging crew
    const bool stage_in = !pos_dev && !direct_in, stage_out = !out_dev && !lay.direct;
    std::shared_ptr<Stager> sg;
    struct CrewGuard {   // on any exit: no worker may still be copying from / into the caller's arrays
        std::shared_ptr<Stager> *s;
        ~CrewGuard() { if (*s) (*s)->shut_down(); }
    } crew_guard{&sg};
    if (stage_in || stage_out) {
        sg = std::make_shared<Stager>();
        sg->p = p;
        sg->lay = lay;
        sg->stage_in = stage_in;
        sg->stage_out = stage_out;
        sg->plan(sizes, j_beg);
        static const int ring_depth = []() { const char *e = getenv("GSF_RING_DEPTH"); return e && *e ? std::max(kSlots, std::min(kMaxRing, atoi(e))) : 5; }();
        const int depth = (int)std::min<int64_t>(n_chunks, n_chunks <= kSlots ? kSlots : ring_depth);
        if (stage_in) {
            sg->ring_in = depth;
            sg->slot_in = (size_t)p.dim * chunk;
            if ((rc = ensure_cap(&d.ring_in, &d.ring_in_cap, sg->slot_in * depth, true))) return rc;
            sg->h_in = d.ring_in;
        }
        if (stage_out) {
            sg->ring_out = depth;
            sg->slot_out = (size_t)nc * chunk;
            if ((rc = ensure_cap(&d.ring_out, &d.ring_out_cap, sg->slot_out * depth, true))) return rc;
            sg->h_out = d.ring_out;
        }
        if ((rc = ensure_ring_events(d, depth))) return rc;
        // workers beyond this thread: none for small jobs (a wake-up costs more than the copy)
        const int64_t staged_bytes = ((stage_in ? p.dim : 0) + (stage_out ? nc : 0)) * m_shard * 8;
        const int helpers = staged_bytes >= (1 << 20) ? std::min<int64_t>(host_threads - 1, sg->total_parts - 1) : 0;
        d.staging_threads = 1 + std::max(0, helpers);
        if (helpers > 0) host_pool().submit(sg, helpers);
    }
    int64_t in_polled = 0, out_polled = 0, issued = 0;
    // publish event completions to the crew (never blocks)
    auto poll = [&]() {
        if (stage_in)
            while (in_polled < issued && cudaEventQuery(d.ev_ring_in[(size_t)(in_polled % sg->ring_in)]) == cudaSuccess)
                sg->in_released.store(++in_polled, std::memory_order_release);
        if (stage_out)
            while (out_polled < issued && cudaEventQuery(d.ev_ring_out[(size_t)(out_polled % sg->ring_out)]) == cudaSuccess)
                sg->out_ready.store(++out_polled, std::memory_order_release);
    };
    // wait for a crew condition, helping with the work while waiting
    auto wait_for = [&](const std::function<bool()> &ready, int64_t help_upto) {
        int idle = 0;
        while (!ready()) {
            poll();
            if (sg->step(help_upto)) { idle = 0; continue; }
            if (++idle < 4000) cpu_relax(); else std::this_thread::yield();
        }
    };

    for (int64_t c = 0; c < n_chunks; ++c) {
        const int si = (int)(c % kSlots);
        Slot &sl = d.slot[si];
        cudaStream_t st = sl.stream;
        const int64_t cnt = sizes[(size_t)c];
        const int64_t jc = starts[(size_t)c];

        // ---- input
        const double *kpos;
        int64_t kps0, kps1;
        if (pos_dev) {
            kpos = p.pos + jc * p.ps1;
            kps0 = p.ps0;
            kps1 = p.ps1;
        } else {
            if ((rc = ensure_cap(&sl.d_pos, &sl.d_pos_cap, (size_t)p.dim * cnt, false))) return rc;
            if (direct_in) {
                for (int a = 0; a < p.dim; ++a)
                    GSF_CUDA(cudaMemcpyAsync(sl.d_pos + (size_t)a * cnt, p.pos + a * p.ps0 + jc,
                                             (size_t)cnt * sizeof(double), cudaMemcpyHostToDevice, st));
            } else {
                wait_for([&]() { return sg->in_staged(c); }, c);
                GSF_CUDA(cudaMemcpyAsync(sl.d_pos, sg->in_slot(c), (size_t)p.dim * cnt * sizeof(double),
                                         cudaMemcpyHostToDevice, st));
                GSF_CUDA(cudaEventRecord(d.ev_ring_in[(size_t)(c % sg->ring_in)], st));
            }
            d.h2d_bytes += (int64_t)p.dim * cnt * (int64_t)sizeof(double);
            kpos = sl.d_pos;
            kps0 = cnt;
            kps1 = 1;
        }

        // ---- kernel
        double *kout;
        int64_t kos0, kos1;
        if (out_dev) {
            kout = p.out + jc * p.os1;
            kos0 = p.os0;
            kos1 = p.os1;
        } else {
            if ((rc = ensure_cap(&sl.d_out, &sl.d_out_cap, (size_t)nc * cnt, false))) return rc;
            kout = sl.d_out;
            if (lay.aos) { kos0 = 1; kos1 = nc; } else { kos0 = cnt; kos1 = 1; }
        }
        int Pc = P, Lc = L;
        if (cnt != chunk) choose_variant(d, p, cnt, pipelined, &Pc, &Lc);
        if ((rc = launch_sum(d, p, kpos, kps0, kps1, kout, kos0, kos1, cnt, st, Pc, Lc))) return rc;

        // ---- output
        if (!out_dev) {
            if (lay.direct) {
                if (nc == 1 || lay.aos) {
                    GSF_CUDA(cudaMemcpyAsync(p.out + jc * p.os1, sl.d_out, (size_t)nc * cnt * sizeof(double),
                                             cudaMemcpyDeviceToHost, st));
                } else {
                    for (int a = 0; a < nc; ++a)
                        GSF_CUDA(cudaMemcpyAsync(p.out + a * p.os0 + jc, sl.d_out + (size_t)a * cnt,
                                                 (size_t)cnt * sizeof(double), cudaMemcpyDeviceToHost, st));
                }
            } else {
                // the ring slot must have been drained of the chunk that used it last
                if (c >= sg->ring_out) wait_for([&]() { return sg->out_scattered(c - sg->ring_out); }, c + 1);
                GSF_CUDA(cudaMemcpyAsync(sg->out_slot(c), sl.d_out, (size_t)nc * cnt * sizeof(double),
                                         cudaMemcpyDeviceToHost, st));
                GSF_CUDA(cudaEventRecord(d.ev_ring_out[(size_t)(c % sg->ring_out)], st));
            }
            d.d2h_bytes += (int64_t)nc * cnt * (int64_t)sizeof(double);
        }
        d.chunks++;
        issued = c + 1;
        if (sg) poll();
    }

    // drain: scatter the remaining output chunks as their copies land
    if (stage_out) {
        int64_t c_done = 0;
        wait_for([&]() {
            while (c_done < n_chunks && sg->out_scattered(c_done)) ++c_done;
            return c_done == n_chunks;
        }, n_chunks);
    }
    g_trace.mark("shard: chunks issued");
    for (int s = 0; s < (n_chunks > 1 ? kSlots : 1); ++s) GSF_CUDA(cudaStreamSynchronize(d.slot[s].stream));
    g_trace.mark("shard: streams synced");
    d.ws_used = false;   // everything that read the workspace has finished
    guard.armed = false;
    return GSF_OK;
}

// Contiguous point range of shard g of G: boundaries rounded down to multiples of 1024 points
// (so every shard but the last starts and ends on a chunk/tile boundary), last shard takes the rest.
void shard_bounds(int64_t m, int G, int g, int64_t *j0, int64_t *j1)
{
    *j0 = m * g / G / 1024 * 1024;
    *j1 = g + 1 == G ? m : m * (g + 1) / G / 1024 * 1024;
}

#include "gsf_grid_host.inc"

// Caching allocator for pinned host memory handed to callers (result arrays): D2H then lands
// directly in the array the caller sees, with no staging copy.  cudaMallocHost costs ~0.5 ms/MB,
// so freed blocks are kept (up to GSF_PINNED_CACHE_MB, default 2048) and reused.
class PinnedPool {
  public:
    int alloc(size_t bytes, void **out)
    {
        const size_t sz = (bytes + 65535) / 65536 * 65536;
        {
            std::lock_guard<std::mutex> l(mu_);
            // Page-locked memory cannot be swapped: callers that keep many results alive (ensembles)
            // must not pin the host down.  Beyond the cap the request is refused and the caller
            // falls back to ordinary memory (the Python module: np.empty + staged D2H).
            if (live_bytes_ + sz > live_cap())
                return fail(GSF_ERR_ALLOC, "pinned pool: %zu bytes live, request of %zu exceeds GSF_PINNED_LIVE_MB",
                            live_bytes_, sz);
            auto it = free_.lower_bound(sz);
            if (it != free_.end() && it->first <= sz + sz / 4) {
                *out = it->second;
                live_[it->second] = it->first;
                live_bytes_ += it->first;
                cached_ -= it->first;
                free_.erase(it);
                return GSF_OK;
            }
        }
        void *p = nullptr;
        cudaError_t e = cudaMallocHost(&p, sz);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return fail(e == cudaErrorMemoryAllocation ? GSF_ERR_ALLOC : GSF_ERR_NO_DEVICE,
                        "cudaMallocHost(%zu) failed: %s", sz, cudaGetErrorString(e));
        }
        std::lock_guard<std::mutex> l(mu_);
        live_[p] = sz;
        live_bytes_ += sz;
        *out = p;
        return GSF_OK;
    }
    int release(void *p)
    {
        size_t sz = 0;
        {
            std::lock_guard<std::mutex> l(mu_);
            auto it = live_.find(p);
            if (it == live_.end()) return fail(GSF_ERR_ARG, "gsf_host_free: unknown pointer");
            sz = it->second;
            live_.erase(it);
            live_bytes_ -= sz;
            if (cached_ + sz <= cap()) {
                free_.emplace(sz, p);
                cached_ += sz;
                return GSF_OK;
            }
        }
        cudaFreeHost(p);
        return GSF_OK;
    }
    void trim()
    {
        std::lock_guard<std::mutex> l(mu_);
        for (auto &kv : free_) cudaFreeHost(kv.second);
        free_.clear();
        cached_ = 0;
    }

  private:
    static size_t cap()
    {
        const char *e = getenv("GSF_PINNED_CACHE_MB");
        return (size_t)(e && *e ? atoll(e) : 2048) << 20;
    }
    // live (handed-out) pinned bytes: GSF_PINNED_LIVE_MB, default min(4 GiB, physical RAM / 8)
    static size_t live_cap()
    {
        static const size_t v = []() {
            const char *e = getenv("GSF_PINNED_LIVE_MB");
            if (e && *e) return (size_t)atoll(e) << 20;
            const long pages = sysconf(_SC_PHYS_PAGES), psz = sysconf(_SC_PAGE_SIZE);
            size_t ram8 = pages > 0 && psz > 0 ? (size_t)pages * (size_t)psz / 8 : (size_t)1 << 30;
            return std::min<size_t>((size_t)4 << 30, ram8);
        }();
        return v;
    }
    size_t live_bytes_ = 0;
    std::mutex mu_;
    std::multimap<size_t, void *> free_;
    std::unordered_map<void *, size_t> live_;
    size_t cached_ = 0;
};

PinnedPool &pinned_pool()
{
    static PinnedPool *p = new PinnedPool();
    return *p;
}

std::vector<int> default_devices()
{
    std::vector<int> v;
    const char *e = getenv("GSF_DEVICES");
    if (e && *e) {
        const char *s = e;
        while (*s) {
            char *end;
            long id = strtol(s, &end, 10);
            if (end == s) break;
            v.push_back((int)id);
            s = *end == ',' ? end + 1 : end;
        }
    }
    if (v.empty()) v.push_back(0);
    return v;
}

// Let kernels running on `dev` load/store memory that lives on `owner` (NVLink peer mapping).
bool enable_peer(int dev, int owner)
{
    if (dev == owner) return true;
    int can = 0;
    if (cudaDeviceCanAccessPeer(&can, dev, owner) != cudaSuccess || !can) {
        cudaGetLastError();
        return false;
    }
    if (cudaSetDevice(dev) != cudaSuccess) return false;
    cudaError_t e = cudaDeviceEnablePeerAccess(owner, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) {
        cudaGetLastError();
        return true;
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return true;
}

void collect_stats(const Problem &p, const std::vector<DeviceCtx *> &used, double total_ms, int P, int L,
                   int pos_kind, int out_kind, int grid_path)
{
    Context &c = ctx();
    gsf_stats s{};
    s.total_ms = total_ms;
    s.point_modes = p.N * p.M;
    s.points_per_thread = P;
    s.lanes_per_point = L;
    s.pos_memory = pos_kind;
    s.out_memory = out_kind;
    s.grid_path = grid_path;
    s.poly_degree = grid_path ? 0 : (p.dim > gsf::kMaxTemplateDim ? gsf::kHiDeg : p.deg);
    s.fp64_slots = grid_path ? 0 : p.dim + gsf::cos_slots(s.poly_degree) + p.nc();
    s.n_devices = (int)used.size();
    s.mode_group = grid_path ? (int32_t)std::min<int64_t>(p.mode_group, 0x7fffffff) : 0;
    c.last_devs.clear();
    for (DeviceCtx *d : used) {
        s.h2d_bytes += d->h2d_bytes;
        s.d2h_bytes += d->d2h_bytes;
        s.kernel_launches += d->launches;
        s.n_chunks += d->chunks;
        s.staging_threads = std::max(s.staging_threads, d->staging_threads);
        c.last_devs.push_back(d->dev);
    }
    s.kernel_ms = -1.0;   // resolved lazily in gsf_get_last_stats (needs event sync)
    s.prep_ms = -1.0;
    c.last = s;
}

bool grid_detection_enabled()
{
    Context &c = ctx();
    if (c.grid_detect >= 0) return c.grid_detect != 0;
    const char *e = getenv("GSF_GRID_DETECT");
    return !(e && e[0] == '0');
}

// `grid` != nullptr: the caller gave axis vectors (explicit structured-grid request, p.pos unused).
int run_host_call(Problem p, const GridSpec *grid)
{
    Context &c = ctx();
    std::lock_guard<std::mutex> lock(c.mu);
    if (grid) {
        if (grid->dim != 2 && grid->dim != 3)
            return fail(GSF_ERR_DIM, "structured-grid path supports dim 2 and 3 (dim=%d)", grid->dim);
        p.dim = grid->dim;
        p.M = grid->points();
        for (int a = 0; a < grid->dim; ++a)
            if (grid->n[a] < 0 || (grid->n[a] > 0 && !grid->axis[a])) return fail(GSF_ERR_SHAPE, "bad grid axis %d", a);
        p.pos = p.out;           // placeholder so validate() sees a non-NULL pointer (pos is unused)
    }
    int rc = validate(p);
    if (rc) return rc;
    const int ndev_visible = device_count_raw();
    if (ndev_visible <= 0)
        return fail(GSF_ERR_NO_DEVICE, "no CUDA device available; gsfield has no CPU path");
    if (p.M == 0) {
        c.last = gsf_stats{};
        return GSF_OK;
    }
    const auto t0 = std::chrono::steady_clock::now();
    g_trace.begin();

    int pos_kind = 0, pos_dev = -1, out_kind, out_dev;
    if (!grid) classify(p.pos, &pos_kind, &pos_dev);
    classify(p.out, &out_kind, &out_dev);
    if (grid) {
        for (int a = 0; a < grid->dim; ++a) {
            int k, dv;
            classify(grid->axis[a], &k, &dv);
            if (k == 2) return fail(GSF_ERR_ARG, "grid axis vectors must be host-resident");
        }
    }

    p.deg = choose_degree(p);

    std::vector<int> devs = c.devices_explicit ? c.devices : default_devices();
    for (int id : devs)
        if (id < 0 || id >= ndev_visible) return fail(GSF_ERR_ARG, "configured device %d not visible (%d devices)", id, ndev_visible);
    // Device-resident arguments may still be in production on one of the caller's streams, and the
    // library works on its own non-blocking streams, which nothing orders after those: a blocking
    // entry point therefore waits for all work queued on the owning device(s) first.  (The
    // stream-ordered gsf_summate_on_stream takes the caller's stream instead and never syncs.)
    {
        int seen[8], n_seen = 0;
        auto sync_owner = [&](int dv) {
            if (dv < 0) return;
            for (int i = 0; i < n_seen; ++i)
                if (seen[i] == dv) return;
            if (n_seen < 8) seen[n_seen++] = dv;
            int prev = 0;
            cudaGetDevice(&prev);
            if (cudaSetDevice(dv) == cudaSuccess) cudaDeviceSynchronize();
            cudaSetDevice(prev);
            cudaGetLastError();
        };
        if (pos_kind == 2) sync_owner(pos_dev);
        if (out_kind == 2) sync_owner(out_dev);
        if (p.N > 0) {   // every mode array is classified here, once per call (prepare_modes reuses it)
            int dv = -1;
            classify(p.k, &p.mem_k, &p.mem_k_dev);
            if (p.mem_k == 2) sync_owner(p.mem_k_dev);
            classify(p.z1, &p.mem_z1, &dv);
            if (p.mem_z1 == 2) sync_owner(dv);
            classify(p.z2, &p.mem_z2, &dv);
            if (p.mem_z2 == 2) sync_owner(dv);
            if (p.kind == gsf::kFourier) {
                classify(p.sf, &p.mem_sf, &dv);
                if (p.mem_sf == 2) sync_owner(dv);
            }
        }
    }
    if (pos_kind == 2 || out_kind == 2) {
        // device-resident data: the work runs where the data lives ...
        const int dv = pos_kind == 2 ? pos_dev : out_dev;
        if (pos_kind == 2 && out_kind == 2 && pos_dev != out_dev)
            return fail(GSF_ERR_ARG, "pos on device %d but out on device %d", pos_dev, out_dev);
        // ... unless several devices were configured explicitly and both pos and out are device
        // resident: then every configured device takes a contiguous point shard and reads its
        // positions / writes its results directly in the owner's memory over NVLink (peer
        // mapping) -- no staging copies, no collective.
        bool ok = c.devices_explicit && devs.size() > 1 && pos_kind == 2 && out_kind == 2 && !grid;
        if (ok) {
            bool has_owner = false;
            for (int id : devs) has_owner |= id == dv;
            ok = has_owner;
            for (int id : devs) ok = ok && enable_peer(id, dv);
        }
        if (ok) {
            p.peer_owner = dv;   // peers may read the owner's arrays
        } else {
            devs.assign(1, dv);
        }
    }
    // do not spread tiny problems: at least 2^16 points per device
    int G = (int)devs.size();
    const bool one_device_process = G == 1 && pos_kind != 2 && out_kind != 2;
    G = (int)std::max<int64_t>(1, std::min<int64_t>(G, p.M / 65536));
    devs.resize(G);

    std::vector<DeviceCtx *> used(G);
    for (int g = 0; g < G; ++g)
        if ((rc = get_device_ctx(devs[g], &used[g]))) return rc;
    if (one_device_process) bind_rank_to_gpu_node(devs[0]);   // one process per GPU: stay on the GPU's NUMA node

    g_trace.mark("classify + device ctx");
    // ---- structured grid?  explicit request, or exact auto-detection on host-resident positions
    GridSpec detected;
    const GridSpec *gs = grid;
    // (staged bytes per point: positions always; the result only when it cannot be copied out directly)
    const int staged_bpp = 8 * p.dim + (out_kind == 0 ? 8 * p.nc() : 0);
    const int threads1 = staging_threads(p.threads_hint, 1, p.N, staged_bpp);
    // Worth it only when the general kernel would take longer than the grid path's fixed cost
    // (~5 extra launches, tools/latency_sweep.py: break-even near 2e7 point*modes) and when the
    // per-axis tables stay small next to HBM.
    // (the exact check is one long memory-bound pass: unlike the per-chunk staging copies it
    //  does profit from many threads: one per 2 MB of positions, up to half this rank's cores)
    const int64_t pos_mb = grid ? 0 : (int64_t)p.dim * p.M * 8 / (2 << 20);
    //  (inputs of hundreds of MB -- C4, C5 -- may take three quarters of them: with several GPUs the
    //   check, not the GEMM, is what the caller waits for)
    const int core_share = pos_mb >= 128 ? cores_per_pipeline(1) * 3 / 4 : cores_per_pipeline(1) / 2;
    const int detect_threads = std::max<int>(threads1, (int)std::min<int64_t>(std::min<int64_t>(32, std::max<int64_t>(1, pos_mb)),
                                             (int64_t)std::max(1, core_share)));
    // Speculation: the candidate structure (axis lengths from the first change of each coordinate)
    // costs microseconds; the EXACT check of all points costs a pass over the whole array.  The
    // grid kernels need only the candidate axes, so they start at once while pool workers verify;
    // a mismatch cancels the grid path and the general kernel recomputes everything.
    std::shared_ptr<GridVerify> verify;
    if (!gs && pos_kind != 2 && p.N >= 32 && (double)p.M * (double)p.N >= 2e7 && grid_detection_enabled() &&
        grid_candidate_host(p, &detected) && detected.rows() >= 16) {
        const double table_bytes = 16.0 * (double)(p.N + 16) *
                                   ((double)detected.n[0] + (detected.dim == 3 ? (double)detected.n[1] : 0.0) +
                                    (double)detected.n_last() * p.nc());
        if (table_bytes <= 8e9) {
            gs = &detected;
            verify = std::make_shared<GridVerify>(p, detected);
            const int vt = p.threads_hint > 0 ? std::min(detect_threads, p.threads_hint) : detect_threads;
            if (vt > 1) host_pool().submit(verify, vt - 1);
            else if (!verify->finish()) { gs = nullptr; verify.reset(); }   // one thread: check first
        }
    }
    struct VerifyJoin {   // no exit path may leave workers reading the caller's positions
        std::shared_ptr<GridVerify> *v;
        ~VerifyJoin() { if (*v) (*v)->finish(); }
    } verify_join{&verify};
    const std::atomic<int> *cancel = verify ? &verify->bad : nullptr;

    // device-resident positions and result (synchronous API): detect on the device
    bool axes_on_device = false;
    if (!gs && pos_kind == 2 && out_kind == 2 && G == 1 && p.peer_owner < 0 && p.N >= 32 &&
        (double)p.M * (double)p.N >= 2e7 && grid_detection_enabled()) {
        bool found = false;
        if ((rc = detect_grid_device(*used[0], p, &detected, &found))) return rc;
        if (found) {
            const double table_bytes = 16.0 * (double)(p.N + 16) *
                                       ((double)detected.n[0] + (detected.dim == 3 ? (double)detected.n[1] : 0.0) +
                                        (double)detected.n_last() * p.nc());
            if (table_bytes <= 8e9) {
                gs = &detected;
                axes_on_device = true;
            }
        }
    }

    int P = 0, L = 0;
    if (gs) {
        p.mode_group = detect_mode_group(p);   // tensor-structured modes shrink the GEMM's contraction
        const int64_t R = gs->rows();
        if (G == 1) {
            rc = run_grid(*used[0], p, *gs, 0, R, out_kind, threads1, axes_on_device, cancel);
        } else {
            std::vector<std::thread> th;
            for (int g = 0; g < G; ++g) {
                const int64_t r0 = R * g / G / 32 * 32;
                const int64_t r1 = g + 1 == G ? R : R * (g + 1) / G / 32 * 32;
                th.emplace_back([&, g, r0, r1]() {
                    DeviceCtx &d = *used[g];
                    int r = run_grid(d, p, *gs, r0, r1, out_kind, 1, false, cancel);
                    d.status = r;
                    if (r) d.err = g_err;
                });
            }
            for (auto &t : th) t.join();
            for (int g = 0; g < G; ++g)
                if (used[g]->status && !rc) {
                    rc = used[g]->status;
                    g_err = used[g]->err;
                }
        }
        if (verify) {
            const bool is_grid = verify->finish();
            verify.reset();
            g_trace.mark("grid: verification joined");
            if (!is_grid) {   // speculation failed: not a grid after all
                if (rc) return rc;
                gs = nullptr;
            }
        }
    }
    if (gs) {
        // structured-grid path done
    } else if (G == 1) {
        // Zero-copy: pinned host positions and result are mapped into the device address space
        // (UVA), so ONE full-size launch can read/write them over PCIe directly -- no chunking, no
        // copy-engine traffic.  The algorithmic traffic is 8*(dim+nc) bytes per point against
        // ~15 N FP64 slots, so the link is lightly used once N is large enough.
        // The first resident wave has to wait for its own positions to cross the link, so short
        // tiles (P = 1: ~5 MB in the first wave instead of ~12 MB) start the machine sooner unless
        // the per-point work is long (N >= 4000).  Measured on C2: 1.02 ms vs 1.06 ms for the
        // chunked copy pipeline (tools/e2e_probe.py).  GSF_ZERO_COPY=0 restores the pipeline.
        // Several ranks on one host (one process per GPU): the copy-engine pipeline moves the same bytes
        // with fewer, larger requests than SM-issued mapped loads -- 1.45 vs 1.59 ms per C2 call with
        // eight ranks saturating the host's memory system (profiles/c2_weak_r2.md) -- so mapped access of
        // LARGE pinned arrays is the single-rank default only.  (Small calls keep their mapped buffers.)
        static const int zero_copy = []() { const char *e = getenv("GSF_ZERO_COPY"); return e && *e ? atoi(e) : 1; }();
        static const bool zero_copy_large = []() {
            const char *e = getenv("GSF_ZERO_COPY");
            return e && *e ? atoi(e) != 0 : local_world_size() <= 1;
        }();
        bool zc = false;
        if (zero_copy_large && pos_kind == 1 && out_kind == 1 && p.ps1 == 1) {
            void *dpos = nullptr, *dout = nullptr;
            cudaSetDevice(used[0]->dev);
            if (cudaHostGetDevicePointer(&dpos, const_cast<double *>(p.pos), 0) == cudaSuccess &&
                cudaHostGetDevicePointer(&dout, p.out, 0) == cudaSuccess) {
                Problem q = p;
                q.pos = static_cast<const double *>(dpos);
                q.out = static_cast<double *>(dout);
                q.zero_copy = true;
                rc = run_shard(*used[0], q, 0, p.M, 2, 2, &P, &L, threads1);
                used[0]->h2d_bytes += (int64_t)p.dim * p.M * 8;
                used[0]->d2h_bytes += (int64_t)p.nc() * p.M * 8;
                zc = true;
            } else {
                cudaGetLastError();
            }
        }
        // Small host-resident problems (one chunk anyway): stage the positions into the pinned
        // ring on the host and let the kernel read / write the pinned buffers in place -- saves the
        // explicit H2D and D2H copies and their launch latencies (tools/latency_sweep.py).
        // (threshold swept with tools/e2e_size_sweep.py, GSF_SMALL_KB: the one-launch paths win up to ~800 KB of positions -- 60 vs 77 us at 2e4 points, 78 vs 97 us at 3e4 -- and lose beyond 1 MB)
        static const int64_t small_bytes = []() { const char *e = getenv("GSF_SMALL_KB"); return (int64_t)(e && *e ? atoll(e) : 800) * 1024; }();
        const bool small_host = !zc && zero_copy && pos_kind != 2 && out_kind != 2 && (int64_t)p.dim * p.M * 8 <= small_bytes;
        int64_t fused_rows = small_host ? small_fused_rows(p) : 0;
        static const bool promote = []() { const char *e = getenv("GSF_SMALL_PROMOTE"); return !(e && e[0] == '0'); }();
        if (fused_rows > 0 && promote && small_modes_repeat(*used[0], p)) fused_rows = 0;
        if (fused_rows > 0) {
            P = 1;
            if ((rc = run_small_fused(*used[0], p, fused_rows, &L))) return rc;
            zc = true;
        }
        if (!zc && small_host) {
            DeviceCtx &d0 = *used[0];
            Slot &sl = d0.slot[0];
            const int nc = p.nc();
            OutLayout lay;
            lay.aos = nc > 1 && p.os0 == 1 && p.os1 == nc;
            lay.direct = false;
            cudaSetDevice(d0.dev);
            if (!(rc = ensure_cap(&sl.h_pos, &sl.h_pos_cap, (size_t)p.dim * p.M, true)) &&
                !(rc = ensure_cap(&sl.h_out, &sl.h_out_cap, (size_t)nc * p.M, true))) {
                // (cudaMallocHost memory: under UVA the device address equals the host address)
                gather_pos_part(p, 0, p.M, 0, p.M, sl.h_pos, false);   // streaming stores: the GPU read this buffer last (3 us less than memcpy, tools/micro/launch_floor.cu)
                g_trace.mark("small: gather pos");
                Problem q = p;
                q.pos = sl.h_pos; q.ps0 = p.M; q.ps1 = 1;
                q.out = sl.h_out;
                if (nc == 1) { q.os0 = 0; q.os1 = 1; }
                else if (lay.aos) { q.os0 = 1; q.os1 = nc; }
                else { q.os0 = p.M; q.os1 = 1; }
                q.zero_copy = true;
                rc = run_single_launch(d0, q, &P, &L);
                if (!rc) scatter_out_part(p, lay, 0, p.M, 0, p.M, sl.h_out, true);
                g_trace.mark("small: scatter out");
                d0.h2d_bytes += (int64_t)p.dim * p.M * 8;
                d0.d2h_bytes += (int64_t)nc * p.M * 8;
                zc = true;
            }
            if (rc) return rc;
        }
        if (!zc) rc = run_shard(*used[0], p, 0, p.M, pos_kind, out_kind, &P, &L, threads1);
    } else {
        std::vector<std::thread> th;
        std::vector<int> Ps(G), Ls(G);
        const int threads_g = staging_threads(p.threads_hint, G, p.N, staged_bpp);   // per device: issuing thread + its share of the crew
        for (int g = 0; g < G; ++g) {
            int64_t j0, j1;
            shard_bounds(p.M, G, g, &j0, &j1);
            th.emplace_back([&, g, j0, j1]() {
                DeviceCtx &d = *used[g];
                int r = run_shard(d, p, j0, j1, pos_kind, out_kind, &Ps[g], &Ls[g], threads_g);
                d.status = r;
                if (r) d.err = g_err;   // g_err is thread-local to the worker
            });
        }
        for (auto &t : th) t.join();
        for (int g = 0; g < G; ++g)
            if (used[g]->status && !rc) {
                rc = used[g]->status;
                g_err = used[g]->err;
            }
        P = Ps[0];
        L = Ls[0];
    }
    if (rc) return rc;
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    collect_stats(p, used, ms, P, L, pos_kind, out_kind, gs ? (grid ? 2 : 1) : 0);
    g_trace.mark("done");
    return GSF_OK;
}

#include "gsf_krige_host.inc"
#include "gsf_variogram_host.inc"

}  // namespace

// =============================================================================================
extern "C" {

int gsf_abi_version(void) { return GSF_ABI_VERSION; }

int gsf_device_count(void) { return device_count_raw(); }

const char *gsf_last_error(void) { return g_err.c_str(); }

int gsf_summate(int dim, int64_t n_modes, int64_t n_points, const double *cov_samples, int64_t cov_s0,
                int64_t cov_s1, const double *z1, int64_t z1_s, const double *z2, int64_t z2_s,
                const double *pos, int64_t pos_s0, int64_t pos_s1, double *out, int num_threads)
{
    Problem p{gsf::kScalar, dim, n_modes, n_points, nullptr, 0, cov_samples, cov_s0, cov_s1, z1, z1_s,
              z2, z2_s, pos, pos_s0, pos_s1, out, 0, 1};
    p.threads_hint = num_threads;
    return run_host_call(p, nullptr);
}

int gsf_summate_incompr(int dim, int64_t n_modes, int64_t n_points, const double *cov_samples,
                        int64_t cov_s0, int64_t cov_s1, const double *z1, int64_t z1_s, const double *z2,
                        int64_t z2_s, const double *pos, int64_t pos_s0, int64_t pos_s1, double *out,
                        int64_t out_s0, int64_t out_s1, int num_threads)
{
    Problem p{gsf::kIncompr, dim, n_modes, n_points, nullptr, 0, cov_samples, cov_s0, cov_s1, z1, z1_s,
              z2, z2_s, pos, pos_s0, pos_s1, out, out_s0, out_s1};
    p.threads_hint = num_threads;
    return run_host_call(p, nullptr);
}

int gsf_summate_fourier(int dim, int64_t n_modes, int64_t n_points, const double *spectrum_factor,
                        int64_t sf_s, const double *modes, int64_t modes_s0, int64_t modes_s1,
                        const double *z1, int64_t z1_s, const double *z2, int64_t z2_s, const double *pos,
                        int64_t pos_s0, int64_t pos_s1, double *out, int num_threads)
{
    Problem p{gsf::kFourier, dim, n_modes, n_points, spectrum_factor, sf_s, modes, modes_s0, modes_s1,
              z1, z1_s, z2, z2_s, pos, pos_s0, pos_s1, out, 0, 1};
    p.threads_hint = num_threads;
    return run_host_call(p, nullptr);
}

int gsf_summate_on_stream(int kind, int dim, int64_t n_modes, int64_t n_points, const double *spectrum_factor,
                          int64_t sf_s, const double *cov_samples, int64_t cov_s0, int64_t cov_s1,
                          const double *z1, int64_t z1_s, const double *z2, int64_t z2_s, const double *pos,
                          int64_t pos_s0, int64_t pos_s1, double *out, int64_t out_s0, int64_t out_s1,
                          void *cuda_stream)
{
    if (kind < 0 || kind > 2) return fail(GSF_ERR_ARG, "kind must be 0, 1 or 2");
    Problem p{kind, dim, n_modes, n_points, spectrum_factor, sf_s, cov_samples, cov_s0, cov_s1, z1, z1_s,
              z2, z2_s, pos, pos_s0, pos_s1, out, out_s0, out_s1};
    if (kind != gsf::kIncompr) { p.os0 = 0; p.os1 = 1; }
    Context &c = ctx();
    std::lock_guard<std::mutex> lock(c.mu);
    int rc = validate(p);
    if (rc) return rc;
    if (device_count_raw() <= 0)
        return fail(GSF_ERR_NO_DEVICE, "no CUDA device available; gsfield has no CPU path");
    if (p.M == 0) return GSF_OK;
    int pk, pd, ok, od;
    classify(p.pos, &pk, &pd);
    classify(p.out, &ok, &od);
    if (pk != 2 || ok != 2 || pd != od)
        return fail(GSF_ERR_ARG, "gsf_summate_on_stream needs pos and out in device memory of one device");
    DeviceCtx *d;
    if ((rc = get_device_ctx(pd, &d))) return rc;
    struct DeviceRestore {   // every return below leaves the caller's current device untouched
        int prev = -1;
        DeviceRestore() { if (cudaGetDevice(&prev) != cudaSuccess) prev = -1; }
        ~DeviceRestore() { if (prev >= 0) cudaSetDevice(prev); }
    } restore;
    struct CacheGuard {      // a failed call must not leave "records valid" behind: they may never have been produced
        DeviceCtx *d;
        bool armed = true;
        ~CacheGuard() { if (armed) { d->rec_valid = false; d->ws_used = false; } }
    } cache_guard{d};
    GSF_CUDA(cudaSetDevice(pd));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    reset_call_counters(*d);
    p.deg = choose_degree(p);
    const auto t0 = std::chrono::steady_clock::now();
    if (d->ws_used) GSF_CUDA(cudaStreamWaitEvent(st, d->ev_ws, 0));
    if ((rc = prepare_modes(*d, p, st, gsf::amp_factor(p.deg)))) return rc;
    int P, L;
    choose_variant(*d, p, p.M, false, &P, &L);
    if ((rc = launch_sum(*d, p, p.pos, p.ps0, p.ps1, p.out, p.os0, p.os1, p.M, st, P, L))) return rc;
    GSF_CUDA(cudaEventRecord(d->ev_ws, st));
    d->ws_used = true;
    d->chunks = 1;
    cache_guard.armed = false;
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    std::vector<DeviceCtx *> used(1, d);
    collect_stats(p, used, ms, P, L, 2, 2, 0);
    return GSF_OK;
}

int gsf_krige(int64_t n_cond, int64_t n_points, const double *krig_mat, int64_t mat_s0, int64_t mat_s1,
              const double *krig_vecs, int64_t vecs_s0, int64_t vecs_s1, const double *cond, int64_t cond_s,
              double *field, double *error, int num_threads)
{
    (void)num_threads;
    KrigeProblem k{n_cond, n_points, krig_mat, mat_s0, mat_s1, krig_vecs, vecs_s0, vecs_s1, cond, cond_s, field, error};
    return run_krige(k);
}

int gsf_variogram_structured(int64_t n0, int64_t n1, const double *f, int64_t f_s0, int64_t f_s1,
                             const uint8_t *mask, int64_t mask_s0, int64_t mask_s1, char estimator_type,
                             double *variogram, int num_threads)
{
    (void)num_threads;
    VarioStructProblem v{n0, n1, f, f_s0, f_s1, mask, mask_s0, mask_s1, estimator_type == 'c', variogram};
    return run_variogram_struct(v);
}

int gsf_variogram_unstructured(int dim, int64_t n_fields, int64_t n_points, int64_t n_bins, const double *f,
                               int64_t f_s0, int64_t f_s1, const double *bin_edges, int64_t edges_s,
                               const double *pos, int64_t pos_s0, int64_t pos_s1, char estimator_type,
                               char distance_type, double *variogram, uint64_t *counts, int num_threads)
{
    (void)num_threads;
    VarioProblem v{};
    v.mode = distance_type == 'e' ? gsf::kVarEuclid : gsf::kVarHaversine;   // src/variogram.rs:74-87
    v.dim = dim; v.nf = n_fields; v.M = n_points; v.nb = n_bins; v.nd = 0;
    v.f = f; v.fs0 = f_s0; v.fs1 = f_s1;
    v.edges = bin_edges; v.es = edges_s;
    v.pos = pos; v.ps0 = pos_s0; v.ps1 = pos_s1;
    v.cressie = estimator_type == 'c';                                       // src/variogram.rs:24-38
    v.variogram = variogram; v.counts = counts;
    return run_variogram_pairs(v);
}

int gsf_variogram_directional(int dim, int64_t n_fields, int64_t n_points, int64_t n_bins, int64_t n_dirs,
                              const double *f, int64_t f_s0, int64_t f_s1, const double *bin_edges,
                              int64_t edges_s, const double *pos, int64_t pos_s0, int64_t pos_s1,
                              const double *direction, int64_t dir_s0, int64_t dir_s1, double angles_tol,
                              double bandwidth, int separate_dirs, char estimator_type, double *variogram,
                              uint64_t *counts, int num_threads)
{
    (void)num_threads;
    VarioProblem v{};
    v.mode = gsf::kVarDirectional;
    v.dim = dim; v.nf = n_fields; v.M = n_points; v.nb = n_bins; v.nd = n_dirs;
    v.f = f; v.fs0 = f_s0; v.fs1 = f_s1;
    v.edges = bin_edges; v.es = edges_s;
    v.pos = pos; v.ps0 = pos_s0; v.ps1 = pos_s1;
    v.dir = direction; v.ds0 = dir_s0; v.ds1 = dir_s1;
    v.angles_tol = angles_tol; v.bandwidth = bandwidth; v.separate = separate_dirs != 0;
    v.cressie = estimator_type == 'c';
    v.variogram = variogram; v.counts = counts;
    return run_variogram_pairs(v);
}

int gsf_debug_variogram_thresholds(double edge, double angles_tol, double *sqrt_thr, double *acos_thr)
{
    if (sqrt_thr) *sqrt_thr = sqrt_threshold(edge);
    if (acos_thr) *acos_thr = acos_threshold(angles_tol);
    return GSF_OK;
}

int gsf_summate_ex(const gsf_request *r)
{
    if (!r || r->struct_size != (int32_t)sizeof(gsf_request))
        return fail(GSF_ERR_ARG, "gsf_summate_ex: NULL request or struct_size mismatch (ABI %d)", GSF_ABI_VERSION);
    if (r->kind < 0 || r->kind > 2) return fail(GSF_ERR_ARG, "kind must be 0, 1 or 2");
    Problem p{r->kind, r->dim, r->n_modes, r->n_points, r->spectrum_factor, r->sf_s, r->modes, r->modes_s0,
              r->modes_s1, r->z1, r->z1_s, r->z2, r->z2_s, r->pos, r->pos_s0, r->pos_s1, r->out, r->out_s0,
              r->out_s1};
    if (r->kind != gsf::kIncompr) { p.os0 = 0; p.os1 = r->out_s1 ? r->out_s1 : 1; }
    p.threads_hint = r->num_threads;
    p.scale = r->scale;
    for (int a = 0; a < 3; ++a) p.offset[a] = r->offset[a];
    if (r->n_axes > 0) {
        if (r->n_axes != r->dim) return fail(GSF_ERR_SHAPE, "n_axes (%d) must equal dim (%d)", r->n_axes, r->dim);
        GridSpec g;
        g.dim = r->dim;
        for (int a = 0; a < r->dim && a < 3; ++a) {
            g.axis[a] = r->axis[a];
            g.n[a] = r->axis_n[a];
            g.s[a] = r->axis_s[a] ? r->axis_s[a] : 1;
        }
        return run_host_call(p, &g);
    }
    return run_host_call(p, nullptr);
}

/* Host-logic introspection (used by the CPU test-suite; no device needed). */
int gsf_debug_chunk_schedule(int64_t n_points, int64_t forced_chunk, int64_t *sizes, int max_sizes)
{
    if (n_points < 0 || !sizes || max_sizes < 1) return fail(GSF_ERR_ARG, "bad arguments");
    const std::vector<int64_t> v = chunk_schedule(n_points, false, forced_chunk);
    const int n = (int)std::min<size_t>(v.size(), (size_t)max_sizes);
    for (int i = 0; i < n; ++i) sizes[i] = v[(size_t)i];
    return (int)v.size() > max_sizes ? -(int)v.size() : n;
}

int64_t gsf_debug_mode_group(int dim, int64_t n_modes, const double *modes, int64_t modes_s0, int64_t modes_s1)
{
    if (!modes || dim < 1 || n_modes < 0) return 1;
    Problem p{};
    p.dim = dim; p.N = n_modes; p.k = modes; p.ks0 = modes_s0; p.ks1 = modes_s1;
    p.mem_k = 0;   // host-resident by contract of this entry point
    return detect_mode_group(p);
}

int gsf_debug_parse_cpulist(const char *list, int *cpus, int max_cpus)
{
    if (!list || (max_cpus > 0 && !cpus)) return fail(GSF_ERR_ARG, "bad arguments");
    cpu_set_t set;
    const int n = parse_cpulist(list, &set);
    int w = 0;
    for (int c = 0; c < CPU_SETSIZE && w < max_cpus; ++c)
        if (CPU_ISSET(c, &set)) cpus[w++] = c;
    return n;
}

int gsf_debug_detect_grid(int dim, int64_t n_points, const double *pos, int64_t pos_s0, int64_t pos_s1,
                          int64_t *axis_n)
{
    if (!pos || !axis_n) return fail(GSF_ERR_ARG, "bad arguments");
    Problem p{};
    p.dim = dim; p.M = n_points; p.pos = pos; p.ps0 = pos_s0; p.ps1 = pos_s1;
    GridSpec g;
    if (!detect_grid_host(p, &g, staging_threads(0, 1))) return 0;
    for (int a = 0; a < 3; ++a) axis_n[a] = a < dim ? g.n[a] : 1;
    return 1;
}

int gsf_set_grid_detection(int enabled)
{
    Context &c = ctx();
    std::lock_guard<std::mutex> lock(c.mu);
    c.grid_detect = enabled < 0 ? -1 : (enabled != 0);
    return GSF_OK;
}

int gsf_host_alloc(int64_t bytes, void **ptr)
{
    if (bytes <= 0 || !ptr) return fail(GSF_ERR_ARG, "gsf_host_alloc: bad arguments");
    return pinned_pool().alloc((size_t)bytes, ptr);
}

int gsf_host_free(void *ptr)
{
    if (!ptr) return GSF_OK;
    return pinned_pool().release(ptr);
}

int gsf_shard_bounds(int64_t n_points, int n_shards, int shard, int64_t *begin, int64_t *end)
{
    if (n_points < 0 || n_shards < 1 || shard < 0 || shard >= n_shards || !begin || !end)
        return fail(GSF_ERR_ARG, "bad shard request: n_points=%lld n_shards=%d shard=%d", (long long)n_points,
                    n_shards, shard);
    shard_bounds(n_points, n_shards, shard, begin, end);
    return GSF_OK;
}

int gsf_set_devices(const int *device_ids, int n)
{
    Context &c = ctx();
    std::lock_guard<std::mutex> lock(c.mu);
    if (n <= 0 || !device_ids) {
        c.devices.clear();
        c.devices_explicit = false;
        return GSF_OK;
    }
    const int vis = device_count_raw();
    for (int i = 0; i < n; ++i)
        if (device_ids[i] < 0 || device_ids[i] >= vis)
            return fail(GSF_ERR_ARG, "device %d not visible (%d devices)", device_ids[i], vis);
    c.devices.assign(device_ids, device_ids + n);
    c.devices_explicit = true;
    return GSF_OK;
}

int gsf_set_chunk_points(int64_t chunk_points)
{
    Context &c = ctx();
    std::lock_guard<std::mutex> lock(c.mu);
    if (chunk_points < 0) return fail(GSF_ERR_ARG, "chunk_points < 0");
    c.chunk_points = chunk_points;
    return GSF_OK;
}

int gsf_set_variant(int points_per_thread, int lanes_per_point)
{
    Context &c = ctx();
    std::lock_guard<std::mutex> lock(c.mu);
    if (points_per_thread == 0 && lanes_per_point == 0) {
        c.force_p = c.force_l = 0;
        return GSF_OK;
    }
    if (!pick_kernel(3, false, points_per_thread, lanes_per_point))
        return fail(GSF_ERR_ARG, "no kernel variant P=%d L=%d", points_per_thread, lanes_per_point);
    c.force_p = points_per_thread;
    c.force_l = lanes_per_point;
    return GSF_OK;
}

int gsf_set_poly_degree(int degree)
{
    Context &c = ctx();
    std::lock_guard<std::mutex> lock(c.mu);
    if (degree != 0 && degree != gsf::kHiDeg && degree != gsf::kFastDeg)
        return fail(GSF_ERR_ARG, "polynomial degree must be 0 (automatic), %d or %d", gsf::kFastDeg, gsf::kHiDeg);
    c.poly_degree = degree;
    return GSF_OK;
}

int gsf_host_register(void *ptr, int64_t bytes)
{
    if (!ptr || bytes <= 0) return fail(GSF_ERR_ARG, "gsf_host_register: bad arguments");
    if (device_count_raw() <= 0) return fail(GSF_ERR_NO_DEVICE, "no CUDA device available");
    cudaError_t e = cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterDefault);
    if (e == cudaErrorHostMemoryAlreadyRegistered) {
        cudaGetLastError();
        return GSF_OK;
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(e == cudaErrorMemoryAllocation ? GSF_ERR_ALLOC : GSF_ERR_CUDA, "cudaHostRegister(%lld bytes) failed: %s",
                    (long long)bytes, cudaGetErrorString(e));
    }
    return GSF_OK;
}

int gsf_host_unregister(void *ptr)
{
    if (!ptr) return GSF_OK;
    cudaError_t e = cudaHostUnregister(ptr);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(GSF_ERR_CUDA, "cudaHostUnregister failed: %s", cudaGetErrorString(e));
    }
    return GSF_OK;
}

int gsf_set_profiling(int enabled)
{
    Context &c = ctx();
    std::lock_guard<std::mutex> lock(c.mu);
    c.profiling = enabled < 0 ? 0 : (enabled > 2 ? 2 : enabled);
    for (DeviceCtx *d : c.dctx)
        if (d) d->prof_used = 0;
    return GSF_OK;
}

int gsf_get_last_stats(gsf_stats *out)
{
    if (!out) return fail(GSF_ERR_ARG, "NULL stats");
    Context &c = ctx();
    std::lock_guard<std::mutex> lock(c.mu);
    gsf_stats s = c.last;
    double kmax = -1.0, pmax = -1.0;
    for (int dev : c.last_devs) {
        DeviceCtx *d = c.dctx[dev];
        if (!d || d->prof_used == 0) continue;
        cudaSetDevice(dev);
        double sum = 0.0;
        for (size_t i = 0; i < d->prof_used; ++i) {
            float ms = 0.f;
            if (cudaEventSynchronize(d->prof[i].second) != cudaSuccess ||
                cudaEventElapsedTime(&ms, d->prof[i].first, d->prof[i].second) != cudaSuccess) {
                cudaGetLastError();
                sum = -1.0;
                break;
            }
            sum += ms;
        }
        kmax = std::max(kmax, sum);
        if (d->prep_timed) {
            float ms = 0.f;
            if (cudaEventSynchronize(d->prep_end) == cudaSuccess &&
                cudaEventElapsedTime(&ms, d->prep_beg, d->prep_end) == cudaSuccess)
                pmax = std::max(pmax, (double)ms);
            else
                cudaGetLastError();
        }
    }
    s.kernel_ms = kmax;
    s.prep_ms = pmax;
    *out = s;
    return GSF_OK;
}

int gsf_shutdown(void)
{
    Context &c = ctx();
    std::lock_guard<std::mutex> lock(c.mu);
    for (DeviceCtx *d : c.dctx) free_device_ctx(d);
    c.dctx.clear();
    pinned_pool().trim();
    c.last = gsf_stats{};
    c.last_devs.clear();
    return GSF_OK;
}

static int fp64_peak(int device, double min_ms, bool tensor, double *per_s, double *elapsed_ms)
{
    Context &c = ctx();
    std::lock_guard<std::mutex> lock(c.mu);
    if (device_count_raw() <= 0) return fail(GSF_ERR_NO_DEVICE, "no CUDA device available");
    DeviceCtx *d;
    int rc = get_device_ctx(device, &d);
    if (rc) return rc;
    GSF_CUDA(cudaSetDevice(device));
    double *sink;
    GSF_CUDA(cudaMalloc((void **)&sink, sizeof(double)));
    cudaEvent_t e0, e1;
    GSF_CUDA(cudaEventCreate(&e0));
    GSF_CUDA(cudaEventCreate(&e1));
    const int threads = 256;
    const int blocks = d->sm_count * 8;   // 2048 threads per SM
    cudaStream_t st = d->slot[0].stream;
    int iters = tensor ? 200 : 2000;
    float ms = 0.f;
    for (int attempt = 0; attempt < 14; ++attempt) {
        GSF_CUDA(cudaEventRecord(e0, st));
        if (tensor)
            gsf::gsf_dmma_peak_kernel<<<blocks, threads, 0, st>>>(sink, iters, 0.999999, 1e-9);
        else
            gsf::gsf_dfma_peak_kernel<<<blocks, threads, 0, st>>>(sink, iters, -1e-12, 1e-9);
        GSF_CUDA(cudaGetLastError());
        GSF_CUDA(cudaEventRecord(e1, st));
        GSF_CUDA(cudaEventSynchronize(e1));
        GSF_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (ms >= min_ms || iters > (1 << 28)) break;
        const double scale = ms > 0.05 ? std::min(16.0, 1.25 * min_ms / ms) : 16.0;
        iters = (int)std::min<double>((double)iters * std::max(scale, 1.5), 1 << 29);
    }
    double n;
    if (tensor)   // thread-level FMA: 256 per warp-level DMMA
        n = (double)blocks * (threads / 32) * (double)iters * gsf::kDmmaChains * gsf::kDmmaUnroll * 256.0;
    else
        n = (double)blocks * threads * (double)iters * gsf::kPeakChains * gsf::kPeakUnroll;
    if (per_s) *per_s = n / (ms * 1e-3);
    if (elapsed_ms) *elapsed_ms = ms;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    return GSF_OK;
}

int gsf_dfma_peak(int device, double min_ms, double *dfma_per_s, double *elapsed_ms)
{
    return fp64_peak(device, min_ms, false, dfma_per_s, elapsed_ms);
}

int gsf_dmma_peak(int device, double min_ms, double *fma_per_s, double *elapsed_ms)
{
    return fp64_peak(device, min_ms, true, fma_per_s, elapsed_ms);
}

}  // extern "C"
