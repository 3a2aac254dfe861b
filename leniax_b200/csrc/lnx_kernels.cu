// leniax_b200: C ABI (include/leniax_b200.h) and host-side launch code.  The resident 128x128 kernels compile in their own
// translation units (lnx_tu_tm.cu, lnx_tu_generic.cu; interfaces in lnx_internal.h); this one holds the device code of
//   lnx_tiled.cuh / lnx_tiled64.cuh / lnx_tiled2k.cuh   multi-pass engines for worlds that are not 128x128
//   lnx_conv.cuh                                        direct-convolution potential (fft=False)
// There is no CPU fallback anywhere in this library.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

#include "lnx_internal.h"
#include "lnx_resident_common.cuh"
#include "lnx_tiled.cuh"
#include "lnx_tiled64.cuh"
#include "lnx_tiled64h.cuh"
#include "lnx_tiled2k.cuh"
#include "lnx_conv.cuh"


namespace lnx {
// What qd.update_individuals reads of a scan (leniax/qd.py:168-186): per world N and, for every scalar statistic, the mean of
// rows [ns - window, ns) with ns = max(int(N), window) (clamped to the T rows that exist).  planes: one [n_sols][T][n_init]
// array per statistic; out [n_sols][n_init][1 + LNX_NB_STATS].  One thread per (statistic, world); threads of a warp are
// consecutive initialisations, so every row read is coalesced.
struct SummArgs {
    const float* planes[ST_COUNT];
    const float* n_alive;
    float* out;
    int n_sols, T, n_init, window;
};
__global__ void __launch_bounds__(128) lnx_summarize_kernel(SummArgs P) {
    const int i = blockIdx.x * 128 + threadIdx.x, s = blockIdx.y, k = blockIdx.z;
    if (i >= P.n_init) return;
    const float N = P.n_alive[(size_t)s * P.n_init + i];
    const int w = P.window < P.T ? P.window : P.T;
    int ns = (int)N;
    ns = ns < w ? w : (ns > P.T ? P.T : ns);
    const int lo = ns - P.window > 0 ? ns - P.window : 0;
    const float* col = P.planes[k] + (size_t)s * P.T * P.n_init + i;
    float acc = 0.f;
    for (int t = lo; t < ns; ++t) acc += col[(size_t)t * P.n_init];
    float* o = P.out + ((size_t)s * P.n_init + i) * (1 + ST_COUNT);
    o[1 + k] = acc / (float)(ns - lo);
    if (k == 0) o[0] = N;
}
}  // namespace lnx

// =====================================================================================================================
// C ABI
// =====================================================================================================================
using namespace lnx;

static thread_local char g_err[512] = "";
int lnx_fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
#define fail lnx_fail

struct lnx_plan {
    lnx_desc d;
    int device;
    int sm_count;
    bool tiled;            // false: resident 128x128 kernels, true: multi-pass tiled engine
    bool stats_only;       // world size is not a power of two: the plan serves lnx_compute_stats only (the scans need powers of two;
                           // the direct-convolution update lnx_update_conv takes any 2-D size and needs no plan)
    lnx::tiled::Geom g;    // tiled engine geometry
};


// ---------------------------------------------------------------------------------------------------------------------
// tiled engine: host side
// ---------------------------------------------------------------------------------------------------------------------
namespace th {
using namespace lnx::tiled;

static int ilog2i(int n) {
    int l = 0;
    while ((1 << l) < n) ++l;
    return l;
}
// geometry of a world of ANY size for the stand-alone statistics (lnx_compute_stats): one leading index per slab
static bool make_geom_any(int nd, const int32_t* dims, Geom* g, const char** why) {
    memset(g, 0, sizeof(*g));
    if (nd != 2 && nd != 3) {
        *why = "only 2-D and 3-D worlds are supported";
        return false;
    }
    for (int d = 0; d < nd; ++d) {
        if (dims[d] < 1 || dims[d] > NMAX) {
            *why = "every world dimension must be in [1, 4096]";
            return false;
        }
        g->dims[d] = dims[d];
    }
    g->nd = nd;
    g->L = dims[0];
    g->A1 = nd == 3 ? dims[1] : 1;
    g->A2 = dims[nd - 1];
    g->rows = g->L * g->A1;
    g->cells = (long long)g->rows * g->A2;
    g->slab_rows = g->A1;
    g->n_slabs = g->L;
    g->any_size = 1;
    return true;
}
static bool make_geom(int nd, const int32_t* dims, Geom* g, const char** why) {
    memset(g, 0, sizeof(*g));
    if (nd != 2 && nd != 3) {
        *why = "only 2-D and 3-D worlds are supported";
        return false;
    }
    for (int d = 0; d < nd; ++d) {
        const int n = dims[d];
        if (n < 8 || n > NMAX || (n & (n - 1))) {
            *why = "every world dimension must be a power of two in [8, 4096]";
            return false;
        }
        g->dims[d] = n;
    }
    g->nd = nd;
    g->L = dims[0];
    g->A1 = nd == 3 ? dims[1] : 1;
    g->A2 = dims[nd - 1];
    g->logL = ilog2i(g->L);
    g->logA1 = ilog2i(g->A1);
    g->logA2 = ilog2i(g->A2);
    g->half = g->A2 / 2 + 1;
    g->rows = g->L * g->A1;
    g->cells = (long long)g->rows * g->A2;
    g->spec = (long long)g->rows * g->half;
    if (nd == 3) {
        g->slab_rows = g->A1;
    } else {
        int r = 4096 / g->A2;
        if (r < 2) r = 2;
        if (r > g->rows) r = g->rows;
        g->slab_rows = r;
    }
    g->n_slabs = g->rows / g->slab_rows;
    int tc = 8192 / g->L;
    if (tc < 4) tc = 4;
    if (tc > 64) tc = 64;
    g->tc = tc;
    return true;
}
// lead kernel of the 64^3 engine: 6 CTAs of 64 threads per SM (166 registers, no spills; 8 CTAs / 128 registers measured slower)
static void launch_lead64(const PassBArgs& b, unsigned images, cudaStream_t st) {
    // (round 2: 32 / 96 / 192 threads per CTA measured 38.6 / 33.8 / 42.1 ms against 32.3 ms for 64 on the same box: the CTA width,
    // i.e. the contiguous bytes a CTA touches per plane, is not what holds the pass at 65 % of the HBM rate; profiles/r2_e_lead_tpb.txt)
    const dim3 grid(lnx::t64::COLS / lnx::t64::LEAD_TPB, b.C, images);
    lnx::t64::lead_kernel<6><<<grid, lnx::t64::LEAD_TPB, 0, st>>>(b);
}
static bool is_sq2k(const Geom& g) { return g.nd == 2 && g.dims[0] == 2048 && g.dims[1] == 2048; }
// one thread per slab of partial sums while there are few worlds (a single 2048^2 world has 1024 slabs); 128 threads otherwise
static unsigned pass_d_threads(const Geom& g, long long worlds) {
    if (worlds >= 64) return 128;
    unsigned t = 128;
    while ((int)t < g.n_slabs && t < 32 * PASS_D_MAX_WARPS) t <<= 1;
    return t;
}
static bool line2k_plan(const Geom& g, int C, int K) { return is_sq2k(g) && C == 1 && K == 1; }

// Executable graphs of the step loops.  One per (device, step-loop kind, steps per graph), kept for the life of the process: a scan
// captures its step(s) again (host-only work), lets cudaGraphExecUpdate write the new kernel arguments into the cached executable and
// launches that - no instantiation per scan (0.3 ms for two nodes, and the first instantiations of a 32-node graph in a process cost
// tens of ms when earlier launches are still in flight: profiles/r2_t2k_graph_ab.txt).  An update only affects LATER launches of the
// executable; the ones already enqueued keep the arguments they were enqueued with.  Executables an update refuses (another topology:
// cannot happen for one kind) are replaced; the old one is destroyed by a later call once the event behind its last launch completed.
struct CachedExec {
    int device, kind, n;
    cudaGraphExec_t exec;
};
struct RetiredGraph {
    cudaGraphExec_t exec;
    cudaEvent_t done;
};
static std::mutex g_graph_mu;  // the cache, the retired list, and update + launches of a cached executable as one unit
static std::vector<CachedExec> g_exec_cache;
static std::vector<RetiredGraph> g_retired;
static void sweep_retired_graphs() {  // (g_graph_mu held)
    for (size_t i = 0; i < g_retired.size();) {
        if (cudaEventQuery(g_retired[i].done) == cudaSuccess) {
            cudaGraphExecDestroy(g_retired[i].exec);
            cudaEventDestroy(g_retired[i].done);
            g_retired[i] = g_retired.back();
            g_retired.pop_back();
        } else {
            ++i;
        }
    }
    (void)cudaGetLastError();  // cudaErrorNotReady from the queries is not an error
}
enum GraphKind { GRAPH_T2K_PAIRS = 1, GRAPH_T2K_REAL_ROWS = 2, GRAPH_GENERIC_PASSES = 3, GRAPH_T2K_PAIRS_PDL = 4 };
// One time step whose launches have step-independent arguments (the step index lives in the world's carry), captured into a CUDA
// graph and replayed on `st`: the passes of small / single worlds take 10-30 us each and issuing them one by one cost the host
// 14-20 us per launch (the loop was launch-bound).
constexpr int GRAPH_MIN_STEPS = 8;  // shorter loops (lnx_update = one step) are launched directly
// Steps per graph.  A graph launch is a full dependency on the previous one, and crossing it costs more than an edge between two
// kernel nodes of one graph: config D 10.08 -> 9.49 ms at 16 steps per graph (tools/ab_config_d.py).  LNX_GRAPH_UNROLL=1: one step
// per graph (A/B; read per scan, not once per process: the A/B tool alternates the variants inside one process).
static int graph_unroll() {
    const char* e = getenv("LNX_GRAPH_UNROLL");
    const int v = e ? atoi(e) : 16;
    return v < 1 ? 1 : v > 64 ? 64 : v;
}
// enqueue_step(stream, u): u = position of the step inside its graph (0: no kernel of this loop precedes it in the capture)
template <class EnqueueStep>
static int replay_steps(EnqueueStep enqueue_step, int steps, int kind, cudaStream_t st) {
    if (steps < GRAPH_MIN_STEPS) {
        for (int t = 0; t < steps; ++t) enqueue_step(st, 0);
        const cudaError_t e = cudaGetLastError();
        return e == cudaSuccess ? LNX_OK : fail(LNX_ERR_CUDA, "tiled engine: launch failed: %s", cudaGetErrorString(e));
    }
    const int unroll = steps >= 2 * graph_unroll() ? graph_unroll() : 1;
    int device = 0;
    cudaStream_t cap = nullptr;
    cudaError_t e = cudaGetDevice(&device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking);
    // the main graph (unroll steps, launched steps / unroll times), then one of steps % unroll steps
    for (int part = 0; part < 2 && e == cudaSuccess; ++part) {
        const int n = part == 0 ? unroll : steps % unroll, reps = part == 0 ? steps / unroll : 1;
        if (n == 0) break;
        cudaGraph_t graph = nullptr;
        e = cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal);
        if (e == cudaSuccess) {
            for (int u = 0; u < n; ++u) enqueue_step(cap, u);
            e = cudaStreamEndCapture(cap, &graph);
        }
        if (e == cudaSuccess) {
            std::lock_guard<std::mutex> lk(g_graph_mu);
            sweep_retired_graphs();
            size_t slot = 0;
            while (slot < g_exec_cache.size() &&
                   !(g_exec_cache[slot].device == device && g_exec_cache[slot].kind == kind && g_exec_cache[slot].n == n))
                ++slot;
            cudaGraphExec_t exec = nullptr;
            if (slot < g_exec_cache.size()) {
                exec = g_exec_cache[slot].exec;
                cudaGraphExecUpdateResultInfo info;
                if (cudaGraphExecUpdate(exec, graph, &info) != cudaSuccess) {
                    (void)cudaGetLastError();
                    cudaEvent_t done = nullptr;  // earlier launches of the old executable were enqueued on some stream of this device
                    if (cudaEventCreateWithFlags(&done, cudaEventDisableTiming) == cudaSuccess && cudaEventRecord(done, nullptr) == cudaSuccess)
                        g_retired.push_back({exec, done});  // (legacy default stream: behind every blocking stream; a leak otherwise)
                    exec = nullptr;
                    g_exec_cache[slot] = g_exec_cache.back();
                    g_exec_cache.pop_back();
                }
            }
            if (!exec) {
                e = cudaGraphInstantiate(&exec, graph, 0);
                if (e == cudaSuccess) g_exec_cache.push_back({device, kind, n, exec});
            }
            for (int t = 0; e == cudaSuccess && t < reps; ++t) e = cudaGraphLaunch(exec, st);
        }
        if (graph) cudaGraphDestroy(graph);  // (an executable graph does not refer to the graph it was instantiated / updated from)
    }
    if (cap) cudaStreamDestroy(cap);
    if (e != cudaSuccess) return fail(LNX_ERR_CUDA, "tiled engine: graph capture / launch failed: %s", cudaGetErrorString(e));
    return LNX_OK;
}
// worlds per launch of the 64^3 line engine.  Default: all of them.  LNX_T64_BATCH=n runs every n worlds through ALL their steps
// before the next n start (so that a batch stays L2-resident: 3.2 MB per world); measured SLOWER on B200 (256 worlds x 64 steps:
// 30.2 ms unbatched, 42 / 46 / 55 ms at 64 / 39 / 13 worlds per batch, profiles/r2_e_batch_sweep.txt): a wave of the plane kernel
// takes 43-47 us per step whether its planes come from L2 or HBM - the per-warp dependency chain, not the memory level, sets the
// pace - so smaller launches only lower the number of warps in flight.  Kept for A/B runs.
static int t64_batch(long long worlds) {
    static const int forced = [] {
        const char* e = getenv("LNX_T64_BATCH");
        return e ? atoi(e) : 0;
    }();
    return forced > 0 ? forced : (int)worlds;
}
// worlds per window of the 64^3 whole-scan kernel: a window runs ALL its steps before the next one starts, so that its working set
// (3.2 MB per world) stays in L2.  Measured: DRAM reads fall 7x (52.8 -> 7.5 GB for 256 worlds x 64 steps at 24 worlds per window) but
// the scan gets SLOWER (27.5 / 23.9 / 23.0 ms at 24 / 48 / all worlds per window): the kernel is bound by instruction fetch and
// latency, not by HBM, and a small window starves it of independent work.  Default: one window; LNX_T64_WINDOW=n for A/B runs.
static int t64_window(int worlds) {
    static const int v = [] {
        const char* e = getenv("LNX_T64_WINDOW");
        const int n = e ? atoi(e) : 0;
        return n > 0 ? n : 0;
    }();
    return v > 0 && v < worlds ? v : worlds;
}
// LNX_T64_STREAMS=2: the two halves of the world batch on two streams, so that the SMs could interleave one half's memory-latency-bound
// lead pass with the other half's issue-bound plane pass.  Measured: 30.9 ms against 30.0 ms on one stream (256 worlds x 64 steps,
// profiles/r2_e_two_streams.txt) - every launch already fills the machine, the kernels do not overlap.  Off by default, kept for A/B runs.
static bool t64_two_streams() {
    static const bool on = [] {
        const char* e = getenv("LNX_T64_STREAMS");
        return e && atoi(e) == 2;
    }();
    return on;
}
// per host thread and device: the second stream of the 64^3 engine and its fork / join events
struct SideStream {
    cudaStream_t stream = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
};
static SideStream* side_stream(int dev) {
    static thread_local SideStream tl[64];
    SideStream& s = tl[dev];
    if (!s.stream) {
        cudaError_t e = cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming);
        if (e != cudaSuccess) {
            fail(LNX_ERR_CUDA, "64^3 engine: side stream setup failed: %s", cudaGetErrorString(e));
            s = SideStream();
            return nullptr;
        }
    }
    return &s;
}
static bool is_cube64(const Geom& g) { return g.nd == 3 && g.dims[0] == 64 && g.dims[1] == 64 && g.dims[2] == 64; }
static size_t tw_bytes(int logn) { return ((size_t)1 << (logn - 1)) * sizeof(float2); }  // shared-memory twiddle table of one pass
static int log_inner(const Geom& g) { return g.logA2 > g.logA1 ? g.logA2 : g.logA1; }
static size_t smem_a(const Geom& g) {
    return ((size_t)(g.slab_rows / 2) * g.A2 + (g.nd == 3 ? (size_t)g.A1 * g.half : 0)) * sizeof(float2) + tw_bytes(log_inner(g));
}
static size_t smem_b(const Geom& g, bool two_buf = true) { return (two_buf ? 2 : 1) * (size_t)g.L * g.tc * sizeof(float2) + tw_bytes(g.logL); }
static size_t smem_c(const Geom& g, int C) {
    return ((size_t)g.slab_rows * g.half + (size_t)(g.slab_rows / 2) * g.A2) * sizeof(float2) + (size_t)C * g.slab_rows * g.A2 * sizeof(float) +
           tw_bytes(log_inner(g));
}
constexpr size_t SMEM_LIMIT = 220 * 1024;  // dynamic part; pass C also has ~1 KB of static shared memory

static float2* g_tw2k[64] = {nullptr};  // (cos, sin)(2 pi i / 2048), i < 2048: twiddles of the four-step engine (lnx_tiled2k.cuh)
static float2* g_tw[64] = {nullptr};  // library-owned twiddle master table per device: (cos, sin)(2 pi k / NMAX)
static std::mutex g_tiled_init_mu;
static int ensure_tiled_init(int dev) {
    std::lock_guard<std::mutex> lk(g_tiled_init_mu);  // plans may be created from several host threads
    if (g_tw[dev]) return LNX_OK;
    std::vector<float2> host(NMAX / 2), host2k(2048);
    for (int k = 0; k < NMAX / 2; ++k) {
        const double a = 2.0 * 3.14159265358979323846 * k / NMAX;
        host[k] = make_float2((float)cos(a), (float)sin(a));
    }
    for (int k = 0; k < 2048; ++k) {
        const double a = 2.0 * 3.14159265358979323846 * k / 2048;
        host2k[k] = make_float2((float)cos(a), (float)sin(a));
    }
    float2 *d = nullptr, *d2 = nullptr;
    cudaError_t e = cudaMalloc(&d, host.size() * sizeof(float2));
    if (e == cudaSuccess) e = cudaMemcpy(d, host.data(), host.size() * sizeof(float2), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMalloc(&d2, host2k.size() * sizeof(float2));
    if (e == cudaSuccess) e = cudaMemcpy(d2, host2k.data(), host2k.size() * sizeof(float2), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(pass_a_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(pass_b_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(pass_c_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(lnx::t2k::rows_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lnx::t2k::ROWS_SMEM);
    if (e == cudaSuccess) e = lnx::t2k::set_rows_inv_attributes();
    if (e != cudaSuccess) {
        if (d) cudaFree(d);
        if (d2) cudaFree(d2);
        fail(LNX_ERR_CUDA, "tiled engine setup failed: %s", cudaGetErrorString(e));
        return -1;
    }
    g_tw2k[dev] = d2;
    g_tw[dev] = d;
    return LNX_OK;
}
struct Workspace {   // carve-up of the caller's scratch for one lnx_run_scan call
    float* state;
    float2* spec;
    float2* pot;
    float* partials;
    WorldCarry* carry;
    size_t bytes;
};
static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }
static Workspace carve(const Geom& g, int C, int K, long long worlds, unsigned char* base) {
    Workspace w;
    size_t off = 256;
    w.state = reinterpret_cast<float*>(base + off);
    off += align256((size_t)worlds * C * g.cells * sizeof(float));
    w.spec = reinterpret_cast<float2*>(base + off);
    off += align256((size_t)worlds * C * g.spec * sizeof(float2));
    w.pot = reinterpret_cast<float2*>(base + off);
    off += align256((size_t)worlds * K * g.spec * sizeof(float2));
    w.partials = reinterpret_cast<float*>(base + off);
    off += align256((size_t)worlds * g.n_slabs * NP_T * sizeof(float));
    w.carry = reinterpret_cast<WorldCarry*>(base + off);
    off += align256((size_t)worlds * sizeof(WorldCarry));
    w.bytes = off;
    return w;
}
}  // namespace th

using lnx::host::tm_kernel_exists;

// one-time per-device setup: architecture check (no fallback), twiddle constants, dynamic shared memory opt-in.  Guarded by a
// mutex: plans may be created from several host threads (the header promises concurrent calls are safe).
static std::mutex g_init_mu;
static int ensure_device_init(int* dev_out, int* sms_out) {
    static bool done[64] = {false};
    static int sms[64] = {0};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(LNX_ERR_NO_DEVICE, "no CUDA device: %s (leniax_b200 has no CPU fallback)", cudaGetErrorString(e));
    }
    if (dev < 0 || dev >= 64) return fail(LNX_ERR_INVALID, "device index %d out of range", dev);
    std::lock_guard<std::mutex> lk(g_init_mu);
    if (!done[dev]) {
        cudaDeviceProp prop;
        LNX_CUDA(cudaGetDeviceProperties(&prop, dev));
        if (prop.major != 10)
            return fail(LNX_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only (no fallback)", dev, prop.major,
                        prop.minor);
        int rc = lnx::host::tm_setup_device();
        if (rc != LNX_OK) return rc;
        rc = lnx::host::generic_setup_device();
        if (rc != LNX_OK) return rc;
        rc = lnx::host::gen2_setup_device();
        if (rc != LNX_OK) return rc;
        sms[dev] = prop.multiProcessorCount;
        done[dev] = true;
    }
    if (dev_out) *dev_out = dev;
    if (sms_out) *sms_out = sms[dev];
    return LNX_OK;
}

extern "C" {

int lnx_version(void) { return LNX_VERSION; }
const char* lnx_last_error(void) { return g_err; }

int lnx_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    int ok = 0;
    for (int i = 0; i < n; ++i) {
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, i) == cudaSuccess && p.major == 10) ++ok;
    }
    return ok;
}

int lnx_plan_create(const lnx_desc* d, lnx_plan** out) {
    if (!d || !out) return fail(LNX_ERR_INVALID, "lnx_plan_create: null argument");
    *out = nullptr;
    const bool resident = d->nb_dims == 2 && d->dims[0] == WS && d->dims[1] == WS && !(d->flags & LNX_PLAN_FORCE_TILED);
    lnx::tiled::Geom geom;
    memset(&geom, 0, sizeof(geom));
    bool stats_only = false;
    if (!resident) {
        const char* why = "";
        if (!th::make_geom(d->nb_dims, d->dims, &geom, &why)) {
            if (!th::make_geom_any(d->nb_dims, d->dims, &geom, &why))
                return fail(LNX_ERR_UNSUPPORTED, "unsupported world shape (nb_dims=%d, dims=%d x %d x %d): %s", d->nb_dims, d->dims[0],
                            d->dims[1], d->nb_dims > 2 ? d->dims[2] : 1, why);
            stats_only = true;
        }
        if (!stats_only && (th::smem_a(geom) > th::SMEM_LIMIT || th::smem_b(geom) > th::SMEM_LIMIT || th::smem_c(geom, d->nb_channels) > th::SMEM_LIMIT))
            return fail(LNX_ERR_UNSUPPORTED, "world too large for the tiled engine's shared-memory slabs (A %zu, B %zu, C %zu bytes)",
                        th::smem_a(geom), th::smem_b(geom), th::smem_c(geom, d->nb_channels));
    }
    if (d->nb_channels < 1 || d->nb_channels > MAX_C) return fail(LNX_ERR_INVALID, "nb_channels must be in [1, %d]", MAX_C);
    if (d->nb_kernels < 1 || d->nb_kernels > MAX_K) return fail(LNX_ERR_INVALID, "nb_kernels must be in [1, %d]", MAX_K);
    if (d->nb_slots < d->nb_kernels) return fail(LNX_ERR_INVALID, "nb_slots < nb_kernels");
    for (int k = 0; k < d->nb_kernels; ++k) {
        if (d->c_in[k] < 0 || d->c_in[k] >= d->nb_channels) return fail(LNX_ERR_INVALID, "c_in[%d] out of range", k);
        if (d->slot[k] < 0 || d->slot[k] >= d->nb_slots) return fail(LNX_ERR_INVALID, "slot[%d] out of range", k);
        if (d->gf_id[k] < 0 || d->gf_id[k] >= GF_COUNT) return fail(LNX_ERR_INVALID, "gf_id[%d]: unknown growth function", k);
        if (d->c_out[k] < LNX_COUT_NONE || d->c_out[k] >= d->nb_channels) return fail(LNX_ERR_INVALID, "c_out[%d] out of range", k);
    }
    if (d->state_fn < 0 || d->state_fn >= SF_COUNT) return fail(LNX_ERR_INVALID, "unknown state function %d", d->state_fn);
    if (!(d->R > 0.f) || !(d->stats_dt > 0.f)) return fail(LNX_ERR_INVALID, "R and stats_dt must be positive");

    int dev = 0, sms = 0;
    const int rc = ensure_device_init(&dev, &sms);
    if (rc != LNX_OK) return rc;
    lnx_plan* p = new (std::nothrow) lnx_plan;
    if (!p) return fail(LNX_ERR_INVALID, "out of host memory");
    p->d = *d;
    p->device = dev;
    p->sm_count = sms;
    p->tiled = !resident;
    p->stats_only = stats_only;
    p->g = geom;
    if (p->tiled && !stats_only && th::ensure_tiled_init(dev) != LNX_OK) {
        delete p;
        return LNX_ERR_CUDA;  // message set by ensure_tiled_init
    }
    *out = p;
    return LNX_OK;
}

int lnx_plan_destroy(lnx_plan* p) {
    if (!p) return LNX_OK;
    delete p;
    return LNX_OK;
}

size_t lnx_workspace_bytes(const lnx_plan* p);

size_t lnx_kernel_table_bytes(const lnx_plan* p) {
    if (p && p->stats_only) return 0;
    if (!p) return 0;
    // 2048^2 one-channel one-kernel plans hold the table twice: [n_sols][spec] in the generic layout, then in lnx_tiled2k.cuh's
    if (p->tiled) return (size_t)(th::line2k_plan(p->g, p->d.nb_channels, p->d.nb_kernels) ? 2 : 1) * p->d.nb_kernels * p->g.spec * sizeof(float2);
    return (size_t)p->d.nb_kernels * KTAB_F4 * sizeof(float4);
}

size_t lnx_workspace_bytes_for(const lnx_plan* p, int32_t n_sols, int32_t n_init) {
    if (p && p->stats_only) return 0;
    if (!p) return 0;
    if (!p->tiled) return lnx_workspace_bytes(p);
    return th::carve(p->g, p->d.nb_channels, p->d.nb_kernels, (long long)n_sols * n_init, nullptr).bytes;
}

size_t lnx_workspace_bytes(const lnx_plan* p) {
    if (p && p->stats_only) return 0;
    if (!p) return 0;
    if (p->tiled) return th::carve(p->g, p->d.nb_channels, p->d.nb_kernels, 1, nullptr).bytes;
    // 256 B header (world queue counter) + per-CTA scratch of the multi-channel kernels: [3][C] thread-private images for one CTA per SM
    // (lnx_world128_generic), [C + 1] for two CTAs per SM (lnx_world128_gen2)
    const size_t one = (size_t)3 * p->d.nb_channels, two = 2 * lnx::host::gen2_scratch_planes(p->d.nb_channels);
    return 256 + (size_t)p->sm_count * (one > two ? one : two) * PLANE_F4 * sizeof(float4);
}

int lnx_kernels_prepare(const lnx_plan* p, int32_t n_sols, const void* K_fft, void* table, void* stream) {
    if (p && p->stats_only) return fail(LNX_ERR_UNSUPPORTED, "this plan describes a world whose size is not a power of two: it serves lnx_compute_stats only (lnx_update_conv steps such worlds; the FFT scans need powers of two)");
    if (!p || !K_fft || !table || n_sols < 1) return fail(LNX_ERR_INVALID, "lnx_kernels_prepare: bad argument");
    if (p->tiled) {
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        int* slots_dev = nullptr;
        LNX_CUDA(cudaMallocAsync(&slots_dev, sizeof(int) * MAX_K, st));
        LNX_CUDA(cudaMemcpyAsync(slots_dev, p->d.slot, sizeof(int) * p->d.nb_kernels, cudaMemcpyHostToDevice, st));
        const dim3 grid((unsigned)((p->g.spec + 255) / 256 > 1024 ? 1024 : (p->g.spec + 255) / 256), p->d.nb_kernels, n_sols);
        lnx::tiled::gather_ktab_kernel<<<grid, 256, 0, st>>>(static_cast<const float2*>(K_fft), static_cast<float2*>(table), p->g,
                                                            p->d.nb_kernels, p->d.nb_slots, slots_dev, 1.0f / (float)p->g.cells);
        LNX_CUDA(cudaGetLastError());
        LNX_CUDA(cudaFreeAsync(slots_dev, st));
        if (th::line2k_plan(p->g, p->d.nb_channels, p->d.nb_kernels)) {
            lnx::t2k::gather_ktab_kernel<<<dim3(2048, 1, n_sols), 256, 0, st>>>(static_cast<const float2*>(K_fft),
                                                                               static_cast<float2*>(table) + (size_t)n_sols * p->g.spec, p->d.nb_slots,
                                                                               p->d.slot[0], 1.0f / (float)p->g.cells);
            LNX_CUDA(cudaGetLastError());
        }
        return LNX_OK;
    }
    return lnx::host::prepare_launch(p->d, n_sols, K_fft, table, static_cast<cudaStream_t>(stream));
}

int lnx_rfft2(const lnx_plan* p, int32_t n_images, const float* images, void* spectra, void* stream) {
    if (p && p->stats_only) return fail(LNX_ERR_UNSUPPORTED, "this plan describes a world whose size is not a power of two: it serves lnx_compute_stats only (lnx_update_conv steps such worlds; the FFT scans need powers of two)");
    (void)p;  // the plan is optional here
    if (!images || !spectra || n_images < 1) return fail(LNX_ERR_INVALID, "lnx_rfft2: bad argument");
    const int rc = ensure_device_init(nullptr, nullptr);
    if (rc != LNX_OK) return rc;
    return lnx::host::rfft2_launch(n_images, images, spectra, static_cast<cudaStream_t>(stream));
}

int lnx_rfftn(int32_t nb_dims, const int32_t* dims, int32_t n_images, const float* images, void* spectra, void* stream) {
    if (!dims || !images || !spectra || n_images < 1) return fail(LNX_ERR_INVALID, "lnx_rfftn: bad argument");
    if (nb_dims == 2 && dims[0] == WS && dims[1] == WS) return lnx_rfft2(nullptr, n_images, images, spectra, stream);
    int dev = 0;
    int rc = ensure_device_init(&dev, nullptr);
    if (rc != LNX_OK) return rc;
    lnx::tiled::Geom g;
    const char* why = "";
    if (!th::make_geom(nb_dims, dims, &g, &why)) return fail(LNX_ERR_UNSUPPORTED, "lnx_rfftn: %s", why);
    if (th::smem_a(g) > th::SMEM_LIMIT || th::smem_b(g) > th::SMEM_LIMIT) return fail(LNX_ERR_UNSUPPORTED, "lnx_rfftn: world too large");
    if (th::ensure_tiled_init(dev) != LNX_OK) return LNX_ERR_CUDA;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float2 *sa = nullptr, *sb = nullptr;
    const size_t bytes = (size_t)n_images * g.spec * sizeof(float2);
    LNX_CUDA(cudaMallocAsync(&sa, bytes, st));
    LNX_CUDA(cudaMallocAsync(&sb, bytes, st));
    lnx::tiled::PassAArgs a;
    memset(&a, 0, sizeof(a));
    a.state = images;
    a.spec = sa;
    a.tw = th::g_tw[dev];
    a.g = g;
    a.C = 1;
    const bool line64 = th::is_cube64(g) && n_images <= 65535;
    if (line64)
        lnx::t64::plane_fwd_kernel<<<dim3(64, 1, n_images), 32, 0, st>>>(a);
    else
        lnx::tiled::pass_a_kernel<<<dim3(g.n_slabs, 1, n_images), lnx::tiled::TPB, th::smem_a(g), st>>>(a);
    lnx::tiled::PassBArgs b;
    memset(&b, 0, sizeof(b));
    b.spec = sa;
    b.fwd_out = sb;
    b.tw = th::g_tw[dev];
    b.g = g;
    b.C = 1;
    b.K = 0;
    b.n_init = 1;
    const long long M = g.spec / g.L;
    if (line64)
        th::launch_lead64(b, (unsigned)n_images, st);
    else
        lnx::tiled::pass_b_kernel<<<dim3((unsigned)((M + g.tc - 1) / g.tc), 1, n_images), lnx::tiled::TPB, th::smem_b(g), st>>>(b);
    lnx::tiled::expand_hermitian_kernel<<<dim3(1024, 1, n_images), 256, 0, st>>>(sb, static_cast<float2*>(spectra), g);
    LNX_CUDA(cudaGetLastError());
    LNX_CUDA(cudaFreeAsync(sa, st));
    LNX_CUDA(cudaFreeAsync(sb, st));
    return LNX_OK;
}

int lnx_compute_stats(const lnx_plan* p, int32_t n_worlds, const float* cells, const float* field, const float* potential,
                      int32_t* total_shift_idx, float* mass_centroid, float* mass_angle, float* stats, float* channel_mass, void* stream) {
    using namespace lnx::tiled;
    if (!p || n_worlds < 1 || !cells || !field || !potential || !total_shift_idx || !mass_centroid || !mass_angle || !stats || !channel_mass)
        return fail(LNX_ERR_INVALID, "lnx_compute_stats: bad argument");
    if (n_worlds > 65535) return fail(LNX_ERR_INVALID, "lnx_compute_stats: at most 65535 worlds per call");
    Geom g;
    const char* why = "";
    if (!th::make_geom(p->d.nb_dims, p->d.dims, &g, &why) && !th::make_geom_any(p->d.nb_dims, p->d.dims, &g, &why))
        return fail(LNX_ERR_UNSUPPORTED, "lnx_compute_stats: %s", why);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float* partials = nullptr;
    WorldCarry* carry = nullptr;
    float* n_alive = nullptr;
    LNX_CUDA(cudaMallocAsync(&partials, (size_t)n_worlds * g.n_slabs * NP_T * sizeof(float), st));
    LNX_CUDA(cudaMallocAsync(&carry, (size_t)n_worlds * sizeof(WorldCarry), st));
    LNX_CUDA(cudaMallocAsync(&n_alive, (size_t)n_worlds * sizeof(float), st));
    const int nd = g.nd, tb = 128, nb = (n_worlds + tb - 1) / tb;
    carry_pack_kernel<<<nb, tb, 0, st>>>(carry, total_shift_idx, mass_centroid, mass_angle, n_worlds, nd, false, nullptr, nullptr, nullptr);
    StatsPartialArgs a;
    a.cells = cells;
    a.field = field;
    a.potential = potential;
    a.carry = carry;
    a.partials = partials;
    a.g = g;
    a.C = p->d.nb_channels;
    a.K = p->d.nb_kernels;
    stats_partials_kernel<<<dim3(g.n_slabs, 1, n_worlds), TPB, 0, st>>>(a);
    PassDArgs d;
    memset(&d, 0, sizeof(d));
    d.partials = partials;
    d.carry = carry;
    d.stats = stats;
    d.channel_mass = channel_mass;
    d.n_alive = n_alive;
    d.g = g;
    d.C = p->d.nb_channels;
    d.n_sols = 1;
    d.n_init = n_worlds;
    d.max_iter = 1;
    d.t = 0;
    d.R = p->d.R;
    d.stats_dt = p->d.stats_dt;
    pass_d_kernel<<<n_worlds, 128, 0, st>>>(d);
    carry_pack_kernel<<<nb, tb, 0, st>>>(carry, nullptr, nullptr, nullptr, n_worlds, nd, true, total_shift_idx, mass_centroid, mass_angle);
    LNX_CUDA(cudaGetLastError());
    LNX_CUDA(cudaFreeAsync(partials, st));
    LNX_CUDA(cudaFreeAsync(carry, st));
    LNX_CUDA(cudaFreeAsync(n_alive, st));
    return LNX_OK;
}

int lnx_measure_fp32_peak(int32_t iters, double* tflops, double* ms, void* stream) {
    if (iters < 1 || !tflops) return fail(LNX_ERR_INVALID, "lnx_measure_fp32_peak: bad argument");
    int dev = 0, sms = 0;
    const int rc = ensure_device_init(&dev, &sms);
    if (rc != LNX_OK) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float* d_out = nullptr;
    LNX_CUDA(cudaMalloc(&d_out, 64));
    cudaEvent_t e0, e1;
    LNX_CUDA(cudaEventCreate(&e0));
    LNX_CUDA(cudaEventCreate(&e1));
    const int grid = sms * 4, block = 512;
    int prc = lnx::host::fp32_peak_launch(grid, block, d_out, iters / 8 + 1, st);  // warm-up
    if (prc != LNX_OK) return prc;
    LNX_CUDA(cudaEventRecord(e0, st));
    prc = lnx::host::fp32_peak_launch(grid, block, d_out, iters, st);
    if (prc != LNX_OK) return prc;
    LNX_CUDA(cudaEventRecord(e1, st));
    LNX_CUDA(cudaEventSynchronize(e1));
    float t = 0.f;
    LNX_CUDA(cudaEventElapsedTime(&t, e0, e1));
    const double flops = 2.0 * 128.0 * (double)iters * (double)grid * (double)block;
    *tflops = flops / ((double)t * 1e-3) / 1e12;
    if (ms) *ms = t;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_out);
    return LNX_OK;
}

static int run_scan_tiled(const lnx_plan* p, int32_t n_sols, int32_t n_init, int32_t max_run_iter, const float* cells0, const void* table,
                          const float* gf_params, const float* weights, const float* dt, float* stats, float* channel_mass, float* n_alive,
                          float* final_cells, float* cells_out, float* field_out, float* potential_out, void* workspace,
                          size_t workspace_bytes, void* stream, uint32_t run_flags) {
    using namespace lnx::tiled;
    const Geom& g = p->g;
    const int C = p->d.nb_channels, K = p->d.nb_kernels;
    const long long worlds = (long long)n_sols * n_init;
    if (worlds > 65535) return fail(LNX_ERR_INVALID, "tiled engine: at most 65535 worlds per call (got %lld)", worlds);
    const th::Workspace ws = th::carve(g, C, K, worlds, static_cast<unsigned char*>(workspace));
    if (!workspace || workspace_bytes < ws.bytes)
        return fail(LNX_ERR_INVALID, "lnx_run_scan: workspace too small (%zu < %zu); use lnx_workspace_bytes_for()", workspace_bytes, ws.bytes);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const float2* tw = th::g_tw[p->device];
    float* state = final_cells ? final_cells : ws.state;  // working state (updated in place every step)
    LNX_CUDA(cudaMemcpyAsync(state, cells0, (size_t)worlds * C * g.cells * sizeof(float), cudaMemcpyDeviceToDevice, st));
    LNX_CUDA(cudaMemsetAsync(ws.carry, 0, (size_t)worlds * sizeof(WorldCarry), st));
    PassAArgs a;
    memset(&a, 0, sizeof(a));
    a.state = state;
    a.spec = ws.spec;
    a.tw = tw;
    a.g = g;
    a.C = C;
    PassBArgs b;
    memset(&b, 0, sizeof(b));
    b.spec = ws.spec;
    b.pot_spec = ws.pot;
    b.ktab = static_cast<const float2*>(table);
    b.fwd_out = nullptr;
    b.tw = tw;
    b.g = g;
    b.C = C;
    b.K = K;
    b.n_init = n_init;
    PassCArgs c;
    memset(&c, 0, sizeof(c));
    c.state = state;
    c.pot_spec = ws.pot;
    c.gf_params = gf_params;
    c.weights = weights;
    c.dt = dt;
    c.carry = ws.carry;
    c.partials = ws.partials;
    c.cells_out = cells_out;
    c.field_out = field_out;
    c.potential_out = potential_out;
    c.tw = tw;
    c.g = g;
    c.C = C;
    c.K = K;
    c.n_init = n_init;
    c.max_iter = max_run_iter;
    c.state_fn = p->d.state_fn;
    c.mean = p->d.weighted_average;
    PassDArgs d;
    memset(&d, 0, sizeof(d));
    d.partials = ws.partials;
    d.carry = ws.carry;
    d.stats = stats;
    d.channel_mass = channel_mass;
    d.n_alive = n_alive;
    d.g = g;
    d.C = C;
    d.n_sols = n_sols;
    d.n_init = n_init;
    d.max_iter = max_run_iter;
    d.R = p->d.R;
    d.stats_dt = p->d.stats_dt;
    int per_channel[MAX_C] = {0};
    for (int k = 0; k < K; ++k) {
        b.c_in[k] = p->d.c_in[k];
        c.gf_id[k] = p->d.gf_id[k];
        if (++per_channel[p->d.c_in[k]] > 1) b.two_buf = 1;
    }
    const long long M = g.spec / g.L;
    const dim3 grid_a(g.n_slabs, C, (unsigned)worlds), grid_b((unsigned)((M + g.tc - 1) / g.tc), C, (unsigned)worlds),
        grid_c(g.n_slabs, 1, (unsigned)worlds);
    // 64^3 worlds with one channel and one kernel (BASELINE config E): thread-per-line passes of lnx_tiled64.cuh
    const bool line64 = th::is_cube64(g) && C == 1 && K == 1 && !(run_flags & LNX_RUN_TILED_GENERIC);
    // 2048^2 worlds with one channel and one kernel (BASELINE config D): four-step warp-per-line passes of lnx_tiled2k.cuh
    const bool line2k = th::line2k_plan(g, C, K) && !(run_flags & LNX_RUN_TILED_GENERIC);
    lnx::t2k::Extra x2k;
    x2k.tw = th::g_tw2k[p->device];
    x2k.ktab = static_cast<const float2*>(table) + (size_t)n_sols * g.spec;
    {
        if (line2k) {
            // a step = lead (+ pass D of the previous step) + the fused (inverse rows, update, forward rows of the next step) kernel
            using namespace lnx::t2k;
            d.g.n_slabs = 1024 / ROWS_WARPS;  // rows_inv writes one row of partial sums per CTA (eight row pairs)
            const unsigned nw = (unsigned)worlds;
            // rows kernels: a packed row pair per warp (8 pairs per CTA), or one real row per warp (16 rows per CTA; twice the warps with half
            // the chain, measured no faster: 10.7 ms against 10.4 ms for 256 steps) - same layouts, same lead kernel
            const bool pairs = (run_flags & LNX_RUN_T2K_REAL_ROWS) == 0;
            const bool finite = (run_flags & LNX_RUN_ASSUME_FINITE) != 0;
            // LNX_T2K_PDL=1: programmatic dependent launches between the kernels of a graph - measured SLOWER (10.7 against 9.6 ms: the
            // early-resident CTAs of the next kernel take registers and issue slots from the one-wave kernel that is still running,
            // profiles/r2_t2k_graph_ab.txt); the default is a full dependency
            const char* pdl_env = getenv("LNX_T2K_PDL");
            const bool pdl = pdl_env && atoi(pdl_env) != 0;
            if (pairs)
                rows_fwd_kernel<<<dim3(1024 / ROWS_WARPS, 1, nw), 32 * ROWS_WARPS, ROWS_SMEM, st>>>(a, x2k);  // the first step's forward rows
            else
                rows1_fwd_kernel<<<dim3(N / RR_WARPS, 1, nw), 32 * RR_WARPS, RR_SMEM, st>>>(a, x2k);
            const int rc = th::replay_steps(
                [&](cudaStream_t cap, int u) {
                    // programmatic edges (lnx_tiled2k.cuh: pdl_wait / pdl_launch) between the kernels of one graph; the real-row
                    // kernels have no griddepcontrol.wait and are launched as full dependencies
                    launch_lead(b, x2k, d, nw, pdl && pairs && u > 0, cap);
                    if (pairs)
                        launch_rows_inv(c, x2k, a.spec, nw, finite, pdl, cap);
                    else
                        launch_rows1_inv(c, x2k, a.spec, nw, finite, cap);
                },
                max_run_iter, pairs ? (pdl ? th::GRAPH_T2K_PAIRS_PDL : th::GRAPH_T2K_PAIRS) : th::GRAPH_T2K_REAL_ROWS, st);
            if (rc != LNX_OK) return rc;
            d.t = max_run_iter - 1;  // the last step's statistics
            pass_d_kernel<<<nw, 128, 0, st>>>(d);
        } else if (line64) {
            // a step = lead + the fused (inverse planes, update, forward planes of the next step) kernel + pass D, all worlds per launch
            // (two experiments that did not pay are kept behind environment switches: t64_batch, t64_two_streams)
            const int batch = th::t64_batch(worlds);
            const bool line64_round1 = (run_flags & LNX_RUN_T64_LINE) != 0;
            // Up to 128 worlds: the whole scan as ONE persistent launch (t64h::scan_kernel: work queue + per-world completion counters; no
            // wave tails, no launch gaps: 32 worlds 3.6 ms against 4.3 ms, 96 worlds 9.0 against 9.9 ms).  Above that the launch per pass and
            // step is faster (256 worlds: 22.1 ms against 23.0 ms - eight CTAs per SM at different places of one 70 KB kernel saturate the
            // instruction cache); profiles/r2_scan_ab.jsonl.  LNX_RUN_T64_WHOLE_SCAN / LNX_RUN_T64_STEPWISE force either.
            const bool scan_ok = !line64_round1 && !cells_out && !field_out && !potential_out && batch >= worlds && !th::t64_two_streams() &&
                                 (long long)worlds * max_run_iter * lnx::t64h::SCAN_ITEMS < 2000000000LL;
            const bool whole_scan = scan_ok && !(run_flags & LNX_RUN_T64_STEPWISE) && (worlds <= 128 || (run_flags & LNX_RUN_T64_WHOLE_SCAN));
            if (whole_scan) {
                lnx::t64::plane_fwd_kernel<<<dim3(64, 1, (unsigned)worlds), 32, 0, st>>>(a);  // the first step's forward planes
                int* ctr = nullptr;
                const size_t ctr_bytes = (1 + 3 * (size_t)worlds) * sizeof(int);
                LNX_CUDA(cudaMallocAsync(&ctr, ctr_bytes, st));
                LNX_CUDA(cudaMemsetAsync(ctr, 0, ctr_bytes, st));
                lnx::t64h::ScanArgs m;
                m.queue = ctr;
                m.cnt = ctr + 1;
                m.worlds = (int)worlds;
                m.steps = max_run_iter;
                m.window = th::t64_window((int)worlds);
                const cudaError_t e = lnx::t64h::launch_scan(b, c, d, a.spec, m, (run_flags & LNX_RUN_ASSUME_FINITE) != 0, p->sm_count, st);
                if (e != cudaSuccess) {
                    cudaFreeAsync(ctr, st);
                    return fail(LNX_ERR_CUDA, "64^3 scan kernel: %s", cudaGetErrorString(e));
                }
                d.t = max_run_iter - 1;  // the last step's statistics
                pass_d_kernel<<<(unsigned)worlds, th::pass_d_threads(g, worlds), 0, st>>>(d);
                LNX_CUDA(cudaFreeAsync(ctr, st));
                LNX_CUDA(cudaGetLastError());
                return LNX_OK;
            }
            const bool two = batch >= worlds && worlds >= 2 && th::t64_two_streams();
            th::SideStream* side = two ? th::side_stream(p->device) : nullptr;
            if (two && !side) return LNX_ERR_CUDA;
            auto half = [&](cudaStream_t s, long long w0, unsigned nb, int phase, int tt) {
                PassAArgs a2 = a;
                PassBArgs b2 = b;
                PassCArgs c2 = c;
                PassDArgs d2 = d;
                a2.world0 = b2.world0 = c2.world0 = d2.world0 = (int)w0;
                c2.t = d2.t = tt;
                if (phase == 0) {
                    lnx::t64::plane_fwd_kernel<<<dim3(64, 1, nb), 32, 0, s>>>(a2);
                } else if (line64_round1) {
                    th::launch_lead64(b2, nb, s);
                    lnx::t64::plane_inv_kernel<<<dim3(64, 1, nb), 32, 0, s>>>(c2, a2.spec);
                    pass_d_kernel<<<nb, th::pass_d_threads(g, nb), 0, s>>>(d2);
                } else {  // half-line kernels (lnx_tiled64h.cuh): two threads per 64-point line, 20 warps per SM
                    lnx::t64h::launch_step(b2, c2, a2.spec, nb, (run_flags & LNX_RUN_ASSUME_FINITE) != 0, s);
                    pass_d_kernel<<<nb, th::pass_d_threads(g, nb), 0, s>>>(d2);
                }
            };
            if (two) {
                const long long nA = (worlds + 1) / 2, nB = worlds - nA;
                LNX_CUDA(cudaEventRecord(side->fork, st));
                LNX_CUDA(cudaStreamWaitEvent(side->stream, side->fork, 0));
                half(st, 0, (unsigned)nA, 0, 0);
                half(side->stream, nA, (unsigned)nB, 0, 0);
                for (int tt = 0; tt < max_run_iter; ++tt) {
                    half(st, 0, (unsigned)nA, 1, tt);
                    half(side->stream, nA, (unsigned)nB, 1, tt);
                }
                LNX_CUDA(cudaEventRecord(side->join, side->stream));
                LNX_CUDA(cudaStreamWaitEvent(st, side->join, 0));
            } else {
                for (long long w0 = 0; w0 < worlds; w0 += batch) {
                    const unsigned nb = (unsigned)(worlds - w0 < batch ? worlds - w0 : batch);
                    half(st, w0, nb, 0, 0);
                    for (int tt = 0; tt < max_run_iter; ++tt) half(st, w0, nb, 1, tt);
                }
            }
        } else {
            // generic passes: four launches with step-independent arguments (t < 0: the step index is read from the carry), replayed
            c.t = -1;
            d.t = -1;
            const int rc = th::replay_steps(
                [&](cudaStream_t cap, int) {
                    pass_a_kernel<<<grid_a, TPB, th::smem_a(g), cap>>>(a);
                    pass_b_kernel<<<grid_b, TPB, th::smem_b(g, b.two_buf != 0), cap>>>(b);
                    pass_c_kernel<<<grid_c, TPB, th::smem_c(g, C), cap>>>(c);
                    pass_d_kernel<<<(unsigned)worlds, th::pass_d_threads(g, worlds), 0, cap>>>(d);
                },
                max_run_iter, th::GRAPH_GENERIC_PASSES, st);
            if (rc != LNX_OK) return rc;
        }
    }
    LNX_CUDA(cudaGetLastError());
    return LNX_OK;
}

int lnx_update_conv(const lnx_desc* d, int32_t n_worlds, int32_t kh, int32_t kw, const float* state, const float* kernels,
                    const float* gf_params, const float* weights, float dt, float* state_out, float* field_out, float* potential_out,
                    void* stream) {
    if (!d || !state || !kernels || !gf_params || !weights || !state_out || !field_out || !potential_out)
        return fail(LNX_ERR_INVALID, "lnx_update_conv: null argument");
    if (d->nb_dims != 2) return fail(LNX_ERR_UNSUPPORTED, "lnx_update_conv: the direct-convolution potential is 2-D only (core.py:136: strides (1, 1))");
    if (n_worlds < 1 || n_worlds > 65535 || kh < 1 || kw < 1 || d->dims[0] < 1 || d->dims[1] < 1)
        return fail(LNX_ERR_INVALID, "lnx_update_conv: bad sizes");
    if (d->nb_channels < 1 || d->nb_channels > MAX_C || d->nb_kernels < 1 || d->nb_kernels > MAX_K)
        return fail(LNX_ERR_INVALID, "lnx_update_conv: C must be in [1, %d] and K in [1, %d]", MAX_C, MAX_K);
    if ((long long)n_worlds * d->nb_kernels > 65535) return fail(LNX_ERR_INVALID, "lnx_update_conv: n_worlds * K must be <= 65535");
    const int rc = ensure_device_init(nullptr, nullptr);
    if (rc != LNX_OK) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    lnx::conv::ConvArgs a;
    memset(&a, 0, sizeof(a));
    a.state = state;
    a.kernels = kernels;
    a.potential = potential_out;
    a.C = d->nb_channels;
    a.K = d->nb_kernels;
    a.H = d->dims[0];
    a.W = d->dims[1];
    a.kh = kh;
    a.kw = kw;
    lnx::conv::FieldArgs f;
    memset(&f, 0, sizeof(f));
    for (int k = 0; k < a.K; ++k) {
        if (d->c_in[k] < 0 || d->c_in[k] >= a.C || d->slot[k] < 0 || d->slot[k] >= d->nb_slots || d->gf_id[k] < 0 || d->gf_id[k] >= GF_COUNT)
            return fail(LNX_ERR_INVALID, "lnx_update_conv: kernel %d: bad c_in / slot / gf_id", k);
        a.slot[k] = d->slot[k];
        a.c_in[k] = d->c_in[k];
        f.gf_id[k] = d->gf_id[k];
    }
    lnx::conv::potential_kernel<<<dim3((a.W + 31) / 32, (a.H + 7) / 8, n_worlds * a.K), dim3(32, 8), 0, st>>>(a);
    f.state = state;
    f.potential = potential_out;
    f.gf_params = gf_params;
    f.weights = weights;
    f.state_out = state_out;
    f.field_out = field_out;
    f.cells = (long long)a.H * a.W;
    f.C = a.C;
    f.K = a.K;
    f.state_fn = d->state_fn;
    f.mean = d->weighted_average;
    f.dt = dt;
    lnx::conv::field_update_kernel<<<dim3((unsigned)((f.cells + 255) / 256), n_worlds), 256, 0, st>>>(f);
    LNX_CUDA(cudaGetLastError());
    return LNX_OK;
}

static bool use_fused(const lnx_plan* p, bool trajectory) {
    const lnx_desc& d = p->d;
    if (d.nb_channels != 1 || d.nb_kernels != 1 || trajectory) return false;
    // the TMEM kernel exists for every growth function with v1 and for gaussian_target with v2
    return tm_kernel_exists(d.gf_id[0], d.state_fn);
}

int lnx_gen2_schedule(const lnx_desc* d, int32_t* acc_slot, int32_t* acc_first, int32_t* upd_mask, int32_t* chan_slot) {
    if (!d || !acc_slot || !acc_first || !upd_mask || !chan_slot) return fail(LNX_ERR_INVALID, "lnx_gen2_schedule: null argument");
    if (d->nb_channels < 1 || d->nb_channels > MAX_C || d->nb_kernels < 1 || d->nb_kernels > MAX_K)
        return fail(LNX_ERR_INVALID, "lnx_gen2_schedule: C must be in [1, %d] and K in [1, %d]", MAX_C, MAX_K);
    RunArgs a;
    memset(&a, 0, sizeof(a));
    if (!lnx::host::gen2_plan(*d, a)) return 0;
    for (int k = 0; k < d->nb_kernels; ++k) {
        acc_slot[k] = a.acc_slot[k];
        acc_first[k] = (a.acc_first >> k) & 1u;
        upd_mask[k] = a.upd_mask[k];
    }
    for (int c = 0; c < d->nb_channels; ++c) chan_slot[c] = a.chan_slot[c];
    return 1;
}

const char* lnx_run_scan_variant(const lnx_plan* p, int32_t with_trajectory) {
    if (p && p->stats_only) return "unsupported";
    if (!p) return "";
    if (p->tiled) return "tiled";
    if (use_fused(p, with_trajectory != 0)) return "fused";
    RunArgs a;
    return !with_trajectory && lnx::host::gen2_plan(p->d, a) ? "generic2" : "generic";  // generic2 also needs LNX_RUN_WEIGHTS_MATCH_COUT
}

int lnx_run_scan(const lnx_plan* p, int32_t n_sols, int32_t n_init, int32_t max_run_iter, uint32_t run_flags, const float* cells0,
                 const void* table, const float* gf_params, const float* weights, const float* dt, float* stats, float* channel_mass,
                 float* n_alive, float* final_cells, float* cells_out, float* field_out, float* potential_out, void* workspace,
                 size_t workspace_bytes, void* stream) {
    if (p && p->stats_only) return fail(LNX_ERR_UNSUPPORTED, "this plan describes a world whose size is not a power of two: it serves lnx_compute_stats only (lnx_update_conv steps such worlds; the FFT scans need powers of two)");
    if (!p) return fail(LNX_ERR_INVALID, "lnx_run_scan: null plan");
    if (n_sols < 1 || n_init < 1) return fail(LNX_ERR_INVALID, "lnx_run_scan: n_sols and n_init must be >= 1");
    if (max_run_iter < 1) return fail(LNX_ERR_INVALID, "max_run_iter must be positive, value given: %d", max_run_iter);  // runner.py:51
    if (!cells0 || !table || !gf_params || !weights || !dt || !stats || !channel_mass || !n_alive)
        return fail(LNX_ERR_INVALID, "lnx_run_scan: null required pointer");
    if (p->tiled)
        return run_scan_tiled(p, n_sols, n_init, max_run_iter, cells0, table, gf_params, weights, dt, stats, channel_mass, n_alive, final_cells,
                              cells_out, field_out, potential_out, workspace, workspace_bytes, stream, run_flags);
    const bool trajectory = cells_out || field_out || potential_out;
    const bool fused = use_fused(p, trajectory);
    if (!workspace || workspace_bytes < (fused ? (size_t)256 : lnx_workspace_bytes(p)))
        return fail(LNX_ERR_INVALID, "lnx_run_scan: workspace too small (%zu < %zu)", workspace_bytes, lnx_workspace_bytes(p));

    cudaStream_t st = static_cast<cudaStream_t>(stream);
    RunArgs a;
    memset(&a, 0, sizeof(a));
    a.cells0 = cells0;
    a.table = static_cast<const float4*>(table);
    a.gf_params = gf_params;
    a.weights = weights;
    a.dt = dt;
    a.stats = stats;
    a.channel_mass = channel_mass;
    a.n_alive = n_alive;
    a.final_cells = final_cells;
    a.cells_out = cells_out;
    a.field_out = field_out;
    a.potential_out = potential_out;
    a.scratch = reinterpret_cast<float4*>(static_cast<unsigned char*>(workspace) + 256);
    a.queue = static_cast<int*>(workspace);
    a.n_sols = n_sols;
    a.n_init = n_init;
    a.max_iter = max_run_iter;
    a.C = p->d.nb_channels;
    a.K = p->d.nb_kernels;
    a.state_fn = p->d.state_fn;
    a.mean = p->d.weighted_average;
    a.R = p->d.R;
    a.stats_dt = p->d.stats_dt;
    a.flags = run_flags;
    for (int k = 0; k < a.K; ++k) {
        a.c_in[k] = p->d.c_in[k];
        a.gf_id[k] = p->d.gf_id[k];
    }
    LNX_CUDA(cudaMemsetAsync(a.queue, 0, sizeof(int), st));
    const long long n_worlds = (long long)n_sols * n_init;
    const int grid = (int)(n_worlds < p->sm_count ? n_worlds : p->sm_count);
    if (fused) {
        // NaN can only be born from s == 0 or a zero weight (0 * inf); the fast variant assumes neither.  The host
        // cannot see device-side parameters without a sync, so the NaN-propagating variant is the default and the
        // caller opts into the fast one with LNX_RUN_ASSUME_FINITE (set by the Python layer after checking the parameters).
        const int grid2 = (int)(n_worlds < 2 * p->sm_count ? n_worlds : 2 * p->sm_count);  // TMEM-resident state, two worlds per SM
        return lnx::host::tm_launch(p->d.gf_id[0], p->d.state_fn, !(run_flags & LNX_RUN_ASSUME_FINITE), grid2, a, st);
    }
    // several channels / kernels: two worlds per SM when the declared weight pattern (desc.c_out) lets two tensor-memory accumulators
    // suffice and the caller vouches that the weights tensor has no other non-zero entry; else one world per SM
    if (!trajectory && (run_flags & LNX_RUN_WEIGHTS_MATCH_COUT) && !(run_flags & (LNX_RUN_GENERIC_OLD | LNX_RUN_GENERIC_1CTA)) &&
        lnx::host::gen2_plan(p->d, a)) {
        const int grid2 = (int)(n_worlds < 2 * p->sm_count ? n_worlds : 2 * p->sm_count);
        return lnx::host::gen2_launch(grid2, a, st);
    }
    return lnx::host::generic_launch(lnx::host::gen_tm_supports(a.C) && !(run_flags & LNX_RUN_GENERIC_OLD), grid, a, st);
}

int lnx_summarize_stats(const float* const* planes, const float* n_alive, int32_t n_sols, int32_t T, int32_t n_init, int32_t window,
                        float* out, void* stream) {
    if (!planes || !n_alive || !out || n_sols < 1 || T < 1 || n_init < 1 || window < 1 || n_sols > 65535)
        return fail(LNX_ERR_INVALID, "lnx_summarize_stats: bad argument");
    const int rc = ensure_device_init(nullptr, nullptr);
    if (rc != LNX_OK) return rc;
    SummArgs a;
    for (int k = 0; k < ST_COUNT; ++k) {
        if (!planes[k]) return fail(LNX_ERR_INVALID, "lnx_summarize_stats: null statistics plane %d", k);
        a.planes[k] = planes[k];
    }
    a.n_alive = n_alive;
    a.out = out;
    a.n_sols = n_sols;
    a.T = T;
    a.n_init = n_init;
    a.window = window;
    lnx_summarize_kernel<<<dim3((n_init + 127) / 128, n_sols, ST_COUNT), 128, 0, static_cast<cudaStream_t>(stream)>>>(a);
    LNX_CUDA(cudaGetLastError());
    return LNX_OK;
}

int lnx_update(const lnx_plan* p, int32_t n_worlds, const float* state, const void* table, const float* gf_params, const float* weights,
               const float* dt, float* state_out, float* field_out, float* potential_out, void* stream) {
    if (p && p->stats_only) return fail(LNX_ERR_UNSUPPORTED, "this plan describes a world whose size is not a power of two: it serves lnx_compute_stats only (lnx_update_conv steps such worlds; the FFT scans need powers of two)");
    if (!p || !state || !table || !gf_params || !weights || !dt || !state_out || !field_out || !potential_out || n_worlds < 1)
        return fail(LNX_ERR_INVALID, "lnx_update: bad argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int C = p->d.nb_channels;
    const size_t ws_bytes = lnx_workspace_bytes_for(p, 1, n_worlds);
    // one step of the scan with the trajectory outputs = (pre-update cells, field, potential); the statistics it also produces
    // go to stream-ordered scratch and are dropped
    const size_t cells_bytes = (size_t)n_worlds * C * (p->tiled ? (size_t)p->g.cells : (size_t)WS * WS) * sizeof(float);
    float *stats = nullptr, *cm = nullptr, *na = nullptr, *cells_tmp = nullptr;
    void* ws = nullptr;
    LNX_CUDA(cudaMallocAsync(&stats, (size_t)LNX_NB_STATS * n_worlds * sizeof(float), st));
    LNX_CUDA(cudaMallocAsync(&cm, (size_t)n_worlds * C * sizeof(float), st));
    LNX_CUDA(cudaMallocAsync(&na, (size_t)n_worlds * sizeof(float), st));
    LNX_CUDA(cudaMallocAsync(&cells_tmp, cells_bytes, st));
    LNX_CUDA(cudaMallocAsync(&ws, ws_bytes, st));
    const int rc = lnx_run_scan(p, 1, n_worlds, 1, 0, state, table, gf_params, weights, dt, stats, cm, na, state_out, cells_tmp, field_out,
                                potential_out, ws, ws_bytes, st);
    cudaFreeAsync(stats, st);
    cudaFreeAsync(cm, st);
    cudaFreeAsync(na, st);
    cudaFreeAsync(cells_tmp, st);
    cudaFreeAsync(ws, st);
    return rc;
}


}  // extern "C"
