// gta_b200.cu -- C-ABI host layer of libgtessla_b200.so (declared in include/gta_b200.h).
//
// Replaces the reference's OpenMP frame loop (src/gta_tri.c:90-95, :106, :301) by
// pinned-buffer batched frame staging with async H2D on several streams, one fused kernel
// launch per chunk (gta_kernel.cuh) and a slice-wise D2H of the per-frame results.
// Frames are independent: multi-GPU = contiguous frame ranges, no collective.
//
// There is deliberately no CPU path in this file: every entry point needs a CUDA device.
#include <cuda_runtime.h>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include "gta_kernel.cuh"
#include "gta_lpf.cuh"
#include "gta_prepass.cuh"
#include "gta_star.cuh"

namespace {

thread_local std::string g_err;
thread_local double g_last_kernel_ms = 0.0;
thread_local int g_last_launches = 0;
unsigned long long* g_dbg = nullptr;      // device phase-timer array (debug builds only)

int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define CK(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return fail(GTA_B200_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

int round_even(int v) { return (v + 1) & ~1; }
// Levels of the recursion tree with at most this many nodes run as warp-cooperative merges
// (tunable for experiments: GTA_B200_COOP_NODES).
// (environment knobs are read once, under std::call_once: the entry points are called from several host threads)
int env_once(const char* name, int dflt, int* slot, std::once_flag* once) {
    std::call_once(*once, [&] { const char* e = getenv(name); *slot = e ? atoi(e) : dflt; });
    return *slot;
}
int coop_nodes_for(int threads) {
    static int env; static std::once_flag once;
    const int v = env_once("GTA_B200_COOP_NODES", -1, &env, &once);
    return v >= 0 ? v : threads / 32;
}
int pow2_ge(int v) { int p = 2; while (p < v) p <<= 1; return p; }

struct Shape { int threads; size_t smem; int cap, cap_p2; };

int smem_limit() {
    static int lim; static std::once_flag once;
    std::call_once(once, [] {
        int dev = 0, v = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess || v <= 0) { cudaGetLastError(); v = 227 * 1024; }
        lim = v;
    });
    return lim;
}

Shape shape_for(int max_points, int natoms = 0) {
    Shape s;
    s.cap = round_even(std::max(max_points, 4));
    s.cap_p2 = pow2_ge(s.cap);
    s.smem = gt::smem_bytes_for(s.cap, s.cap_p2, natoms);
    s.threads = s.cap <= 192 ? 32 : (s.cap <= 512 ? 64 : 128);
    static int env_slot; static std::once_flag once;        // experiments: GTA_B200_THREADS = 32 | 64 | 128
    const int env_t = env_once("GTA_B200_THREADS", -1, &env_slot, &once);
    if (env_t == 32 || env_t == 64 || env_t == 128) s.threads = env_t;
    return s;
}

template <int T, int MINB>
int launch_t(const gt::KParams& p, const Shape& s, cudaStream_t st) {
    auto kern = gt::tessellate_kernel<T, MINB>;
    static std::mutex mu;
    static size_t configured[64] = {0};
    int dev = 0;
    CK(cudaGetDevice(&dev));
    {
        std::lock_guard<std::mutex> lk(mu);
        if (dev < 64 && configured[dev] < s.smem) {
            CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s.smem));
            configured[dev] = s.smem;
        }
    }
    int occ = 0, sms = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, T, s.smem));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (occ < 1) return fail(GTA_B200_E_TOO_LARGE, "kernel does not fit on an SM");
    int grid = std::min(p.nframes, occ * sms);           // persistent CTAs, frames handed out by an atomic counter
    if (grid < 1) grid = 1;
    kern<<<grid, T, s.smem, st>>>(p);
    CK(cudaGetLastError());
    return GTA_B200_OK;
}

// The throughput path's first stage (gta_prepass.cuh): bulk-copied coordinate tiles, -corr, bucket sort -> point records.
int launch_prepass(const gt::KParams& p, const Shape& s, cudaStream_t st) {
    auto kern = gt::prepass_kernel<256>;
    const size_t smem = gt::PrepassSmem(s.cap, p.natoms, p.in_dim).total;
    int dev = 0, occ = 0, sms = 0;
    CK(cudaGetDevice(&dev));
    if (smem > (size_t)smem_limit()) return fail(GTA_B200_E_TOO_LARGE, "frame too large for the pre-pass kernel");
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, smem));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (occ < 1) return fail(GTA_B200_E_TOO_LARGE, "kernel does not fit on an SM");
    const int grid = std::max(1, std::min(p.nframes, occ * sms));
    kern<<<grid, 256, smem, st>>>(p);
    CK(cudaGetLastError());
    return GTA_B200_OK;
}

int launch(const gt::KParams& p, const Shape& s, cudaStream_t st) {
    switch (s.threads) {
        case 32: return launch_t<32, 8>(p, s, st);
        case 64: return launch_t<64, 4>(p, s, st);
        default: return launch_t<128, 3>(p, s, st);
    }
}

int occupancy_for(const Shape& s) {
    int occ = 0;
    if (s.threads == 32) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, gt::tessellate_kernel<32, 8>, 32, s.smem);
    else if (s.threads == 64) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, gt::tessellate_kernel<64, 4>, 64, s.smem);
    else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, gt::tessellate_kernel<128, 3>, 128, s.smem);
    return occ;
}

// ---- throughput path (gta_lpf.cuh): workspace cache (one per device, only ever grows) and launcher
// One of the device's two workspaces: consecutive launches (on different streams) alternate between them, so that the
// pre-pass of launch k+1 fills the SMs the tail of launch k's walker kernel leaves idle.
struct LpfSet {
    gt::LpfWs ws{}; size_t frames = 0; int cap = 0, cap_e = 0; unsigned int* counter = nullptr; cudaEvent_t done = nullptr;
    int* fb_list = nullptr; unsigned int* fb_count = nullptr;        // star path: the frames of the launch that need this workspace
};
struct LpfCache {
    LpfSet set[2]; unsigned turn = 0;
    cudaEvent_t last_done = nullptr;
    // feedback for the choice of kernel variant (launch_lpf): predicates the FP filter left undecided, as counted on the device
    unsigned long long* undecided_h = nullptr;      // pinned; the device's running count after the last launch
    unsigned long long undecided_seen = 0; int last_frames = 0; bool pending = false, degenerate = false;
    // feedback for the star path: how many frames of the last star launch fell back to the divide-and-conquer pipeline
    unsigned int* fb_h = nullptr; bool fb_pending = false; int fb_frames = 0, fb_natoms = -1, fb_cap = -1, star_skip = 0, skip_natoms = -1, skip_cap = -1; cudaEvent_t fb_done = nullptr;
    unsigned long long* star_counters = nullptr;     // device: [0] frames certified by their stars, [1] frames handed to the divide-and-conquer pipeline
    // stage timers: five events (start, after stars, after pre-pass, after walkers, after areas) per launch since gta_b200_stage_reset
    std::vector<cudaEvent_t> ev; size_t ev_used = 0;
};
LpfCache g_lpf[64];
std::mutex g_lpf_mu;
std::mutex g_lpf_launch_mu[64];          // one launch sequence at a time per device: they share the device's workspace

// Environment knobs, read once (experiments; the defaults are the measured best).
struct LpfEnv { int threshold, slots, threads, carve, split, fslots, defer, defer_phases, defer_rounds; };
const LpfEnv& lpf_env() {
    static LpfEnv e;
    static std::once_flag once;
    std::call_once(once, [] {
        auto geti = [](const char* n, int d) { const char* v = getenv(n); return v ? atoi(v) : d; };
        e.threshold = geti("GTA_B200_LPF", GTA_B200_DEFAULT_BATCH_THRESHOLD);
        e.slots = geti("GTA_B200_LPF_SLOTS", gt::LPF_DEFAULT_SLOTS);
        e.threads = geti("GTA_B200_LPF_THREADS", 0);
        e.carve = geti("GTA_B200_LPF_CARVE", 0);
        e.split = geti("GTA_B200_LPF_SPLIT", -1);
        e.fslots = geti("GTA_B200_LPF_FSLOTS", 0);
        e.defer = geti("GTA_B200_LPF_DEFER", 14);
        e.defer_phases = geti("GTA_B200_LPF_DEFER_PHASES", (1 << gt::LP_LEAF) | (1 << gt::LP_TINIT));
        e.defer_rounds = geti("GTA_B200_LPF_DEFER_ROUNDS", 3);
    });
    return e;
}

// Launches of at least this many frames use the throughput path; smaller ones the fused CTA-per-frame kernel, whose
// per-frame latency is lower. GTA_B200_LPF overrides (0 = never, 1 = always).
std::atomic<int> g_lpf_threshold{-2};
int lpf_threshold() {
    int v = g_lpf_threshold.load();
    if (v == -2) { v = lpf_env().threshold; g_lpf_threshold.store(v); }
    return v;
}
// Which engine serves the batched calls: 0 = automatic (star path unless triangle lists are wanted: their emission order is
// the divide and conquer's), 1 = divide and conquer only, 2 = star path always. GTA_B200_PATH overrides the default.
std::atomic<int> g_path{-1};
int path_mode() {
    int v = g_path.load();
    if (v < 0) { static int env; static std::once_flag once; v = env_once("GTA_B200_PATH", 0, &env, &once); v = (v < 0 || v > 2) ? 0 : v; g_path.store(v); }
    return v;
}
int star_threads(int cap, int forced) {
    if (forced == 64 || forced == 128 || forced == 256 || forced == 320 || forced == 384) return forced;
    return cap <= 160 ? 64 : (cap <= 512 ? 128 : 256);
}
bool use_star(const gt::KParams& p, const Shape& s) {
    const int m = path_mode();
    if (m == 1 || s.cap > 4000 || gt::StarSmem(s.cap, p.natoms, p.in_dim, 384).total > (size_t)smem_limit()) return false;
    return m == 2 || p.tri == nullptr;
}
bool use_lpf(int nframes, const Shape& s) { return lpf_threshold() > 0 && nframes >= lpf_threshold() && s.cap < 4000; }

// Most frames one launch of the throughput path takes: bounds the workspace (32 B per point + 16 B per edge, per frame).
long long lpf_frames_per_launch(const Shape& s) {
    return std::max<long long>(4096, (12ll << 30) / (32ll * std::max(s.cap, 1) + 16ll * gt::pool_edges(s.cap)));
}

int lpf_ensure(int dev, int which, size_t frames, int cap, int cap_e, LpfCache** out) {
    std::lock_guard<std::mutex> lk(g_lpf_mu);
    LpfCache& cc = g_lpf[dev];
    LpfSet& c = cc.set[which];
    if (c.frames < frames || c.cap < cap || c.cap_e < cap_e) {
        // grow only: keep the larger of the old and the new request in every dimension
        frames = std::max(frames, c.frames); cap = std::max(cap, c.cap); cap_e = std::max(cap_e, c.cap_e);
        if (c.done) cudaEventSynchronize(c.done);
        cudaFree(c.ws.pt); cudaFree(c.ws.rec); cudaFree(c.ws.n); cudaFree(c.ws.st); cudaFree(c.ws.err); cudaFree(c.fb_list);
        c.ws = gt::LpfWs{}; c.frames = 0; c.cap = c.cap_e = 0; c.fb_list = nullptr;
        CK(cudaMalloc((void**)&c.ws.pt, sizeof(gt::LPt) * frames * cap));
        CK(cudaMalloc((void**)&c.ws.rec, sizeof(gt::LRec) * (frames + 1) * 2 * (size_t)cap_e));      // (+1: slack behind the last frame)
        CK(cudaMalloc((void**)&c.ws.n, 4 * frames)); CK(cudaMalloc((void**)&c.ws.st, 4 * frames)); CK(cudaMalloc((void**)&c.ws.err, 4 * frames));
        CK(cudaMalloc((void**)&c.fb_list, 4 * frames));
        if (!c.counter) CK(cudaMalloc((void**)&c.counter, 4));
        if (!c.fb_count) CK(cudaMalloc((void**)&c.fb_count, 4));
        if (!c.done) CK(cudaEventCreateWithFlags(&c.done, cudaEventDisableTiming));
        c.ws.cap = cap; c.ws.cap_e = cap_e; c.frames = frames; c.cap = cap; c.cap_e = cap_e;
    }
    if (!cc.undecided_h) { CK(cudaMallocHost((void**)&cc.undecided_h, 8)); *cc.undecided_h = 0; }
    if (!cc.fb_h) { CK(cudaMallocHost((void**)&cc.fb_h, 4)); *cc.fb_h = 0; CK(cudaEventCreateWithFlags(&cc.fb_done, cudaEventDisableTiming)); }
    if (!cc.star_counters) { CK(cudaMalloc((void**)&cc.star_counters, 32)); CK(cudaMemset(cc.star_counters, 0, 32)); }
    *out = &cc;
    return GTA_B200_OK;
}

int launch(const gt::KParams& p, const Shape& s, cudaStream_t st);

// K0 for ONE frame that is too large for the batched kernels (gta_large.cu): the same corr_points() (gta_corr.cuh,
// reference src/gta_tri.c:112-272), one CTA of 1,024 threads reading the atoms from global memory -- the per-interval
// extremes it keeps are a few KB of shared memory however many atoms there are. out[0] = points appended, out[1] = status bits.
constexpr int LC_T = 1024;
__global__ void __launch_bounds__(LC_T) large_corr_kernel(const double* xyz, int natoms, int cap, double bx, double by, double espace, double* corr3, int* out) {
    extern __shared__ __align__(16) unsigned char csm[];
    uint32_t* ctl = reinterpret_cast<uint32_t*>(csm);                          // [0] status
    double* red = reinterpret_cast<double*>(csm + 16);                         // 6 * (T / 32) doubles
    unsigned long long* ckey = reinterpret_cast<unsigned long long*>(csm + 16 + 8 * 6 * (LC_T / 32));
    if (threadIdx.x == 0) ctl[0] = 0;
    __syncthreads();
    int ncorr = 0;
    const int n = gt::corr_points<LC_T>(xyz, 3, natoms, cap, bx, by, espace, ckey, red, ctl, corr3, nullptr, 0, 0, &ncorr);
    __syncthreads();
    if (threadIdx.x == 0) { out[0] = n - natoms; out[1] = (int)ctl[0]; }
}

// The star kernel (gta_star.cuh): one CTA per frame, a thread per point. Frames it does not certify go to ws.fb_list.
int launch_star(const gt::KParams& p, const Shape& s, int* fb_list, unsigned int* fb_count, unsigned long long* counters, cudaStream_t st) {
    gt::StarParams sp; sp.k = p; sp.fb_list = fb_list; sp.fb_count = fb_count; sp.counters = counters;
    static int env_t; static std::once_flag once;            // experiments: GTA_B200_STAR_THREADS = 64 | 128 | 256 | 384
    const int t = env_once("GTA_B200_STAR_THREADS", 0, &env_t, &once);
    const int threads = star_threads(s.cap, t);
    const size_t smem = gt::StarSmem(s.cap, p.natoms, p.in_dim, threads).total;
    if (smem > (size_t)smem_limit()) return fail(GTA_B200_E_TOO_LARGE, "frame too large for the star kernel");
    int dev = 0, sms = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    auto go = [&](auto kern, int threads) -> int {
        int occ = 0;
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        static int env_c; static std::once_flag once_c;      // experiments: GTA_B200_STAR_CARVE = shared-memory carve-out in percent
        const int carve = env_once("GTA_B200_STAR_CARVE", 0, &env_c, &once_c);
        if (carve > 0) CK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem));
        if (occ < 1) return fail(GTA_B200_E_TOO_LARGE, "star kernel does not fit on an SM");
        const int grid = std::max(1, std::min(p.nframes, occ * sms));
        kern<<<grid, threads, smem, st>>>(sp);
        CK(cudaGetLastError());
        return GTA_B200_OK;
    };
    const bool full = !(p.flags & GTA_B200_CORRECT);          // without -corr points the hull is not the box: whole stars
    switch (threads) {
        case 64: return full ? go(gt::star_kernel<64, 8, true>, 64) : go(gt::star_kernel<64, 8, false>, 64);
        case 128: return full ? go(gt::star_kernel<128, 4, true>, 128) : go(gt::star_kernel<128, 4, false>, 128);
        case 320: return full ? go(gt::star_kernel<320, 2, true>, 320) : go(gt::star_kernel<320, 2, false>, 320);
        case 384: return full ? go(gt::star_kernel<384, 2, true>, 384) : go(gt::star_kernel<384, 2, false>, 384);
        default: return full ? go(gt::star_kernel<256, GT_STAR_MINB, true>, 256) : go(gt::star_kernel<256, GT_STAR_MINB, false>, 256);
    }
}

int launch_lpf(gt::KParams p, const Shape& s, cudaStream_t st, bool star = false) {
    int dev = 0, sms = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const LpfEnv& env = lpf_env();
    // A frame starts as 2^split subtree walkers. More walkers per frame = fewer frames (ring pools) in flight for the
    // same number of busy slots, and small batches still fill the GPU.
    // Measured (cfg 3, n = 1,156): split 2 / 3 / 4 / 5 -> 237 / 250 / 263 / 277 k frames/s at 100k frames; cfg 1 (n = 100): 1 is
    // best. Default: subtrees of ~32-64 points.
    int split = env.split;
    if (split < 0) { split = 1; while (split < gt::LPF_MAX_SPLIT && (s.cap >> (split + 6)) >= 1) ++split; }
    split = std::max(0, std::min(split, gt::LPF_MAX_SPLIT));
    // spread the walkers over every SM: fewer slots per block shorten the round when the batch is small
    const long long walkers = (long long)p.nframes << split;
    const int blocks = (int)std::max<long long>(1, std::min<long long>(sms, (walkers + 63) / 64));
    int S = (int)std::min<long long>((walkers + blocks - 1) / blocks, 1 << 20);
    S = std::max(64, std::min((S + 31) & ~31, std::min(env.slots, gt::LPF_MAX_SLOTS)));
    // frame slots per block (12 bytes of shared memory each): a frame averages ~0.6 * 2^split live walkers over its life
    int FS = env.fslots > 0 ? env.fslots : (split == 0 ? S : (5 * S) / (2 << split) + 8);
    FS = std::max(2, std::min(FS, S));
    // the workspace is shared by all streams (and host threads) of this device: launches queue behind each other
    // (copies of the next chunk still overlap). The lock covers grow-workspace, wait-on-previous ... record-done, so
    // that two host threads cannot both queue behind the same earlier launch and then run side by side in one workspace.
    std::lock_guard<std::mutex> launch_lock(g_lpf_launch_mu[dev & 63]);
    LpfCache* c = nullptr;
    const int which = (int)(g_lpf[dev & 63].turn++ & 1u);
    int rc = lpf_ensure(dev, which, (size_t)p.nframes, s.cap, gt::pool_edges(s.cap), &c);
    if (rc) return rc;
    LpfSet& ws = c->set[which];
    CK(cudaStreamWaitEvent(st, ws.done, 0));
    // stage timers (gta_b200_stage_ms): events on the launching stream around each of the three kernels
    cudaEvent_t* tev = nullptr;
    if (c->ev_used + 5 <= 5 * 64) {
        while (c->ev.size() < c->ev_used + 5) { cudaEvent_t e; CK(cudaEventCreate(&e)); c->ev.push_back(e); }
        tev = c->ev.data() + c->ev_used; c->ev_used += 5;
        CK(cudaEventRecord(tev[0], st));
    }
    // 0. star path: every frame it certifies is finished by this launch; the others are listed for the three kernels below.
    // Input that keeps falling back (lattices: every frame has cocircular ties) skips the stars for a while.
    // (automatic mode only, and only for the trajectory shape that fell back)
    if (star && c->fb_pending && cudaEventQuery(c->fb_done) == cudaSuccess) {
        c->fb_pending = false;
        if ((long long)*c->fb_h * 2 > (long long)c->fb_frames) { c->star_skip = 16; c->skip_natoms = c->fb_natoms; c->skip_cap = c->fb_cap; }
    }
    if (star && path_mode() == 0 && c->star_skip > 0 && c->skip_natoms == p.natoms && c->skip_cap == s.cap) { --c->star_skip; star = false; }
    if (star) {
        CK(cudaMemsetAsync(ws.fb_count, 0, 4, st));
        p.ws = gt::LpfWs{};
        rc = launch_star(p, s, ws.fb_list, ws.fb_count, c->star_counters, st);
        if (rc) return rc;
        if (!c->fb_pending) {
            CK(cudaMemcpyAsync(c->fb_h, ws.fb_count, 4, cudaMemcpyDeviceToHost, st));
            CK(cudaEventRecord(c->fb_done, st));
            c->fb_pending = true; c->fb_frames = p.nframes; c->fb_natoms = p.natoms; c->fb_cap = s.cap;
        }
        p.list = ws.fb_list; p.list_count = ws.fb_count;
    }
    if (tev) CK(cudaEventRecord(tev[1], st));
    // 1. pre-pass (load, -corr, sort) into the workspace
    p.ws = ws.ws;
    rc = launch_prepass(p, s, st);
    if (rc) return rc;
    if (tev) CK(cudaEventRecord(tev[2], st));
    // 2. walkers
    gt::LpfParams lp; lp.ws = ws.ws; lp.nframes = p.nframes; lp.nframes_dev = p.list_count; lp.counter = ws.counter; lp.dbg = p.dbg ? p.dbg + 32 : nullptr;
    lp.split = split; lp.fslots = FS; lp.defer_below = env.defer; lp.defer_phases = env.defer_phases; lp.defer_rounds = env.defer_rounds;
    CK(cudaMemsetAsync(ws.counter, 0, 4, st));
    // Which build of the kernel? Without stage A' of the predicates (gt_predicates.cuh) the steps fit 128 registers and
    // 16 warps per SM: best when the FP filter decides everything, as it does on real coordinates (cfg 3). With stage A',
    // 12 warps: 2x faster on degenerate few-bit input (cfg 5). The previous launch on this device tells which kind of
    // data this is: the device counts the undecided predicates (1 lane in 8).
    if (c->pending && c->last_done && cudaEventQuery(c->last_done) == cudaSuccess) {
        const unsigned long long now = *c->undecided_h;
        c->degenerate = (now - c->undecided_seen) * 8 >= 16ull * (unsigned long long)c->last_frames;
        c->undecided_seen = now; c->pending = false;
    }
    {
        const size_t smem = gt::LpfSmem(S, FS).total;
        auto go = [&](auto kern, int threads) -> int {
            CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            // everything that is not shared memory should be L1: ask for the smallest carve-out that holds the block
            CK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, env.carve ? env.carve : (int)std::min<size_t>(100, (smem + 1024) * 100 / (228 * 1024) + 1)));
            kern<<<blocks, threads, smem, st>>>(lp, S);
            CK(cudaGetLastError());
            return GTA_B200_OK;
        };
        // (18 / 20 warps at the 96 registers they leave per thread, 132 bytes of spills in the incircle step: 336 ms instead of
        // 270 ms per 100k frames -- the register file, not the warp count, is what this kernel wants)
        if (c->degenerate || env.threads == 384) rc = go(gt::lpf_wave_kernel<384, true>, 384);
        else rc = go(gt::lpf_wave_kernel<512, false>, 512);
        if (rc) return rc;
    }
    if (tev) CK(cudaEventRecord(tev[3], st));
    // 3. areas
    gt::LpfAreaParams ap; ap.ws = ws.ws; ap.nframes = p.nframes; ap.natoms = p.natoms; ap.flags = p.flags; ap.list = p.list; ap.list_count = p.list_count;
    ap.area = p.area; ap.area2d = p.area2d; ap.status = p.status; ap.ntri = p.ntri; ap.tri = p.tri; ap.tri_cap = p.tri_cap;
    gt::lpf_area_kernel<<<(p.nframes + 3) / 4, 128, 0, st>>>(ap);
    CK(cudaGetLastError());
    if (tev) CK(cudaEventRecord(tev[4], st));
    CK(cudaMemcpyFromSymbolAsync(c->undecided_h, gt::g_xcalls, 8, 0, cudaMemcpyDeviceToHost, st));
    c->last_frames = p.nframes; c->pending = true;
    CK(cudaEventRecord(ws.done, st));
    c->last_done = ws.done;
    return GTA_B200_OK;
}

bool have_device() {
    int n = 0;
    return cudaGetDeviceCount(&n) == cudaSuccess && n > 0;
}

}  // namespace

// d_xyz: device [natoms + ncorr][3], atoms in front; the appended points are written behind them. ncorr = 4 + 2 nx + 2 ny as
// the caller computed it from the box (the kernel computes the same and must agree). *st = GTA_B200_ST_* bits (0: done).
int gta_b200_large_corr(double* d_xyz, int natoms, int ncorr, double bx, double by, double espace, int* st) {
    const size_t smem = 16 + 8 * 6 * (LC_T / 32) + 12 * (size_t)(ncorr + 4) + 16;
    if (smem > (size_t)smem_limit()) { *st = GTA_B200_ST_CAPACITY; return GTA_B200_OK; }       // (a box of > ~4,000 intervals per side)
    int* d_out = nullptr; int h[2] = {0, 0};
    CK(cudaMalloc((void**)&d_out, 8));
    cudaError_t e = cudaFuncSetAttribute(large_corr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) { large_corr_kernel<<<1, LC_T, smem>>>(d_xyz, natoms, natoms + ncorr, bx, by, espace, d_xyz + 3 * (size_t)natoms, d_out); e = cudaGetLastError(); }
    if (e == cudaSuccess) e = cudaMemcpy(h, d_out, 8, cudaMemcpyDeviceToHost);
    cudaFree(d_out);
    if (e != cudaSuccess) return fail(GTA_B200_E_CUDA, std::string("large_corr_kernel: ") + cudaGetErrorString(e));
    *st = h[1];
    if (h[1] == 0 && h[0] != ncorr) return fail(GTA_B200_E_CUDA, "large_corr_kernel: point count differs from the host's");
    return GTA_B200_OK;
}

extern "C" {

const char* gta_b200_large_error(void);
const char* gta_b200_last_error(void) { return g_err.empty() ? gta_b200_large_error() : g_err.c_str(); }
double gta_b200_last_kernel_ms(void) { return g_last_kernel_ms; }

// Stage timers of the throughput path on `device` (< 0: current): forget the launches seen so far / wait for the launches
// since the last reset and add up their three kernels' durations (CUDA events on the launching streams).
int gta_b200_stage_reset(int device) {
    if (device < 0 && cudaGetDevice(&device) != cudaSuccess) return fail(GTA_B200_E_CUDA, "cudaGetDevice");
    if (device < 0 || device >= 64) return fail(GTA_B200_E_ARG, "device index out of range");
    std::lock_guard<std::mutex> launch_lock(g_lpf_launch_mu[device]);
    g_lpf[device].ev_used = 0;
    return GTA_B200_OK;
}
static int stage_sum(int device, double t[4], int* launches) {
    if (device < 0 && cudaGetDevice(&device) != cudaSuccess) return fail(GTA_B200_E_CUDA, "cudaGetDevice");
    if (device < 0 || device >= 64) return fail(GTA_B200_E_ARG, "device index out of range");
    std::lock_guard<std::mutex> launch_lock(g_lpf_launch_mu[device]);
    LpfCache& c = g_lpf[device];
    t[0] = t[1] = t[2] = t[3] = 0.0;
    for (size_t i = 0; i + 5 <= c.ev_used; i += 5) {
        CK(cudaEventSynchronize(c.ev[i + 4]));
        for (int k = 0; k < 4; ++k) { float ms = 0.f; CK(cudaEventElapsedTime(&ms, c.ev[i + k], c.ev[i + k + 1])); t[k] += ms; }
    }
    if (launches) *launches = (int)(c.ev_used / 5);
    return GTA_B200_OK;
}
int gta_b200_stage_ms(int device, double* prepass_ms, double* walk_ms, double* area_ms, int* launches) {
    double t[4];
    const int rc = stage_sum(device, t, launches);
    if (rc) return rc;
    if (prepass_ms) *prepass_ms = t[1];
    if (walk_ms) *walk_ms = t[2];
    if (area_ms) *area_ms = t[3];
    return GTA_B200_OK;
}
// the same with the star kernel's time in front: out4 = {stars, pre-pass, walkers, areas} in ms
int gta_b200_stage_ms4(int device, double* out4, int* launches) {
    double t[4];
    const int rc = stage_sum(device, t, launches);
    if (rc) return rc;
    if (out4) for (int k = 0; k < 4; ++k) out4[k] = t[k];
    return GTA_B200_OK;
}
// running totals of the star kernel on `device` since the last reset (waits for the device): frames certified / handed over
int gta_b200_star_counters(int device, unsigned long long* certified, unsigned long long* fallback, int reset) {
    if (device < 0 && cudaGetDevice(&device) != cudaSuccess) return fail(GTA_B200_E_CUDA, "cudaGetDevice");
    if (device < 0 || device >= 64) return fail(GTA_B200_E_ARG, "device index out of range");
    std::lock_guard<std::mutex> launch_lock(g_lpf_launch_mu[device]);
    LpfCache& c = g_lpf[device];
    unsigned long long v[2] = {0, 0};
    if (c.star_counters) {
        int cur = 0; CK(cudaGetDevice(&cur));
        CK(cudaSetDevice(device));
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(v, c.star_counters, 16, cudaMemcpyDeviceToHost));
        if (reset) CK(cudaMemset(c.star_counters, 0, 16));
        CK(cudaSetDevice(cur));
    }
    if (certified) *certified = v[0];
    if (fallback) *fallback = v[1];
    return GTA_B200_OK;
}
// frames the last star launch on `device` handed to the divide-and-conquer pipeline (waits for that launch), < 0: none seen
long long gta_b200_last_fallback_frames(int device) {
    if (device < 0 && cudaGetDevice(&device) != cudaSuccess) return -1;
    if (device < 0 || device >= 64) return -1;
    std::lock_guard<std::mutex> launch_lock(g_lpf_launch_mu[device]);
    LpfCache& c = g_lpf[device];
    if (!c.fb_h || !c.fb_done) return -1;
    if (cudaEventSynchronize(c.fb_done) != cudaSuccess) return -1;
    return (long long)*c.fb_h;
}
int gta_b200_last_launches(void) { return g_last_launches; }

int gta_b200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int gta_b200_get_path(void) { return path_mode(); }

int gta_b200_set_path(int mode) {
    const int old = path_mode();
    g_path.store(mode < 0 || mode > 2 ? 0 : mode);
    return old;
}

int gta_b200_set_batch_threshold(int min_frames) {
    int old = lpf_threshold();
    g_lpf_threshold.store(min_frames < 0 ? GTA_B200_DEFAULT_BATCH_THRESHOLD : min_frames);
    return old;
}

int gta_b200_max_points(void) {
    // ~64 bytes of shared memory per point (see smem_bytes_for)
    int lim = have_device() ? smem_limit() : 227 * 1024;
    int lo = 4, hi = 8192;
    while (lo < hi) {
        int mid = (lo + hi + 1) / 2;
        Shape s = shape_for(mid);
        if (s.smem <= (size_t)lim && gt::pool_edges(s.cap) < 32760) lo = mid; else hi = mid - 1;
    }
    return lo & ~1;
}

int gta_b200_max_points_for(const double* box, int box_stride, int nframes, int natoms, double espace, unsigned flags) {
    if (!(flags & GTA_B200_CORRECT) || !box) return natoms;
    int by_off = box_stride == 9 ? 4 : 1, best = 0;
    for (int f = 0; f < nframes; ++f) {
        int nx = (int)(box[(size_t)f * box_stride] / espace), ny = (int)(box[(size_t)f * box_stride + by_off] / espace);
        if (nx < 0 || ny < 0 || nx > (1 << 20) || ny > (1 << 20)) continue;
        best = std::max(best, 4 + 2 * nx + 2 * ny);
    }
    return natoms + best;
}

int gta_b200_launch_shape(int max_points, int natoms, int* threads_per_frame, size_t* smem_bytes, int* frames_per_sm) {
    Shape s = shape_for(std::max(max_points, natoms), natoms);
    if (threads_per_frame) *threads_per_frame = s.threads;
    if (smem_bytes) *smem_bytes = s.smem;
    if (frames_per_sm) {
        *frames_per_sm = 0;
        if (have_device() && s.smem <= (size_t)smem_limit()) {
            // the attribute must be raised before the occupancy query sees > 48 KB
            if (s.threads == 32) cudaFuncSetAttribute(gt::tessellate_kernel<32, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s.smem);
            else if (s.threads == 64) cudaFuncSetAttribute(gt::tessellate_kernel<64, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s.smem);
            else cudaFuncSetAttribute(gt::tessellate_kernel<128, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s.smem);
            *frames_per_sm = occupancy_for(s);
        }
    }
    return GTA_B200_OK;
}

int gta_b200_tessellate_device(const double* d_xyz, int in_dim, const double* d_box, int box_stride,
                               int nframes, int natoms, int max_points, double espace, unsigned flags,
                               double* d_area, double* d_area2d, double* d_areabox, int* d_status,
                               int* d_tri, int tri_cap, int* d_ntri,
                               double* d_corr, int corr_cap, int* d_ncorr,
                               unsigned int* d_work, void* stream) {
    g_err.clear();
    if (!have_device()) return fail(GTA_B200_E_NO_DEVICE, "no CUDA device: libgtessla_b200 has no CPU path");
    if (nframes < 0 || natoms < 0 || (in_dim != 2 && in_dim != 3) || !d_xyz || !d_work)
        return fail(GTA_B200_E_ARG, "bad argument");
    if ((flags & GTA_B200_CORRECT) && (!d_box || !(espace > 0.0)))
        return fail(GTA_B200_E_ARG, "-corr needs the box and espace > 0");
    if (nframes == 0) return GTA_B200_OK;
    Shape s = shape_for(std::max(max_points, natoms), natoms);
    if (s.smem > (size_t)smem_limit() || gt::pool_edges(s.cap) >= 32760)
        return fail(GTA_B200_E_TOO_LARGE, "max_points exceeds the shared-memory kernel (see gta_b200_max_points)");
    cudaStream_t st = (cudaStream_t)stream;
    gt::KParams p;
    p.xyz = d_xyz; p.in_dim = in_dim; p.box = d_box; p.box_stride = box_stride; p.by_off = (box_stride == 9) ? 4 : 1;
    p.nframes = nframes; p.natoms = natoms; p.cap = s.cap; p.cap_p2 = s.cap_p2;
    p.espace = espace; p.flags = flags;
    p.area = d_area; p.area2d = d_area2d; p.areabox = d_areabox; p.status = d_status;
    p.tri = d_tri; p.tri_cap = tri_cap; p.ntri = d_ntri;
    p.corr = d_corr; p.corr_cap = corr_cap; p.ncorr = d_ncorr;
    p.work = d_work; p.coop_nodes = coop_nodes_for(s.threads); p.dbg = g_dbg; p.ws = gt::LpfWs{};
    CK(cudaMemsetAsync(d_work, 0, sizeof(unsigned int), st));
    const bool star = use_star(p, s);
    const bool lpf = star || use_lpf(nframes, s);
    if (!lpf) { g_last_launches = 1; return launch(p, s, st); }                      // the fused kernel
    // throughput path: pre-pass + walker kernel + area kernel per launch; batches beyond one launch's worth are split evenly
    const long long per_launch = lpf_frames_per_launch(s);
    const long long parts = (nframes + per_launch - 1) / per_launch;
    g_last_launches = (star ? 4 : 3) * (int)parts;
    for (long long k = 0; k < parts; ++k) {
        const long long f0 = (long long)nframes * k / parts, f1 = (long long)nframes * (k + 1) / parts;
        gt::KParams q = p;
        q.nframes = (int)(f1 - f0);
        q.xyz = d_xyz + (size_t)f0 * natoms * in_dim;
        if (d_box) q.box = d_box + (size_t)f0 * box_stride;
        if (d_area) q.area = d_area + f0;
        if (d_area2d) q.area2d = d_area2d + f0;
        if (d_areabox) q.areabox = d_areabox + f0;
        if (d_status) q.status = d_status + f0;
        if (d_tri) q.tri = d_tri + (size_t)f0 * tri_cap * 3;
        if (d_ntri) q.ntri = d_ntri + f0;
        if (d_corr) q.corr = d_corr + (size_t)f0 * corr_cap * 3;
        if (d_ncorr) q.ncorr = d_ncorr + f0;
        if (k) CK(cudaMemsetAsync(d_work, 0, sizeof(unsigned int), st));
        const int rc = launch_lpf(q, s, st, star);
        if (rc) return rc;
    }
    return GTA_B200_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------
// Host-buffer pipeline
// ------------------------------------------------------------------------------------------
namespace {

struct Job {
    const double* xyz = nullptr;             // flat [nframes][natoms][in_dim] or
    const double* const* frames = nullptr;   // one pointer per frame (reference layout)
    int in_dim = 3;
    const double* box = nullptr; int box_stride = 0;
    int nframes = 0, natoms = 0; double espace = 0.8; unsigned flags = 0;
    int max_points = 0;
    double *area = nullptr, *area2d = nullptr, *areabox = nullptr; int* status = nullptr;
    int* tri = nullptr; int tri_cap = 0; int* ntri = nullptr;
    double* corr = nullptr; int corr_cap = 0; int* ncorr = nullptr;
    int host_threads = 0;                    // staging threads (the meaning -nthreads keeps); <= 0: all cores
};

// One in-flight chunk: pinned + device buffers, a stream and two events.
struct Slot {
    cudaStream_t st = nullptr; cudaEvent_t k0 = nullptr, k1 = nullptr;
    double *h_in = nullptr, *d_in = nullptr;           // coords
    double *h_box = nullptr, *d_box = nullptr;
    double *h_out = nullptr, *d_out = nullptr;         // area | area2d | areabox  (3 * chunk)
    int *h_st = nullptr, *d_st = nullptr;              // status | ntri | ncorr    (3 * chunk)
    int* d_tri = nullptr; double* d_corr = nullptr;    // optional, copied straight to the caller
    unsigned int* d_work = nullptr;
    size_t in_bytes = 0; int chunk = 0, box_stride = 0; size_t tri_elems = 0, corr_elems = 0;
    int f0 = 0, nf = 0; bool busy = false;
};

void slot_free(Slot& s) {
    if (s.st) cudaStreamDestroy(s.st);
    if (s.k0) cudaEventDestroy(s.k0);
    if (s.k1) cudaEventDestroy(s.k1);
    cudaFreeHost(s.h_in); cudaFree(s.d_in); cudaFreeHost(s.h_box); cudaFree(s.d_box);
    cudaFreeHost(s.h_out); cudaFree(s.d_out); cudaFreeHost(s.h_st); cudaFree(s.d_st);
    cudaFree(s.d_tri); cudaFree(s.d_corr); cudaFree(s.d_work);
    s = Slot();
}

int slot_ensure(Slot& s, int chunk, size_t frame_bytes, int box_stride, size_t tri_elems, size_t corr_elems, bool need_pinned_in) {
    size_t in_bytes = frame_bytes * (size_t)chunk;
    if (s.st && s.in_bytes >= in_bytes && (s.h_in || !need_pinned_in) && s.chunk >= chunk && s.box_stride >= box_stride &&
        s.tri_elems >= tri_elems && s.corr_elems >= corr_elems) return GTA_B200_OK;
    slot_free(s);
    CK(cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking));
    CK(cudaEventCreate(&s.k0)); CK(cudaEventCreate(&s.k1));
    if (need_pinned_in) CK(cudaHostAlloc((void**)&s.h_in, in_bytes, cudaHostAllocDefault));     // not needed when the caller's buffer is page-locked
    CK(cudaMalloc((void**)&s.d_in, in_bytes));
    size_t bb = sizeof(double) * (size_t)std::max(box_stride, 1) * chunk;
    CK(cudaHostAlloc((void**)&s.h_box, bb, cudaHostAllocDefault)); CK(cudaMalloc((void**)&s.d_box, bb));
    CK(cudaHostAlloc((void**)&s.h_out, sizeof(double) * 3 * (size_t)chunk, cudaHostAllocDefault));
    CK(cudaMalloc((void**)&s.d_out, sizeof(double) * 3 * (size_t)chunk));
    CK(cudaHostAlloc((void**)&s.h_st, sizeof(int) * 3 * (size_t)chunk, cudaHostAllocDefault));
    CK(cudaMalloc((void**)&s.d_st, sizeof(int) * 3 * (size_t)chunk));
    if (tri_elems) CK(cudaMalloc((void**)&s.d_tri, sizeof(int) * tri_elems));
    if (corr_elems) CK(cudaMalloc((void**)&s.d_corr, sizeof(double) * corr_elems));
    CK(cudaMalloc((void**)&s.d_work, sizeof(unsigned int)));
    s.in_bytes = in_bytes; s.chunk = chunk; s.box_stride = box_stride; s.tri_elems = tri_elems; s.corr_elems = corr_elems;
    return GTA_B200_OK;
}

// Per device: a small set of reusable slots (pinned allocations cost ~1 s per GB, and the sharded entry points
// run each device from a fresh host thread, so the cache cannot be thread-local). A job holds its device's cache
// for its duration: concurrent calls on one device queue behind each other.
struct SlotCache { std::vector<Slot> slots; std::mutex mu; };
SlotCache g_caches[64];

int drain(Slot& s, const Job& j, double* kernel_ms) {
    if (!s.busy) return GTA_B200_OK;
    static int trace_slot; static std::once_flag trace_once;
    const bool trace = env_once("GTA_B200_TRACE", 0, &trace_slot, &trace_once) != 0;
    const auto t_begin = std::chrono::steady_clock::now();
    CK(cudaStreamSynchronize(s.st));
    if (trace) fprintf(stderr, "[gta_b200] chunk %d+%d: waited %.1f ms for copies + kernels\n", s.f0, s.nf, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count());
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, s.k0, s.k1) == cudaSuccess) *kernel_ms += ms;
    const int c = s.chunk;
    if (j.area) memcpy(j.area + s.f0, s.h_out, sizeof(double) * s.nf);
    if (j.area2d && (j.flags & GTA_B200_2D)) memcpy(j.area2d + s.f0, s.h_out + c, sizeof(double) * s.nf);
    if (j.areabox) memcpy(j.areabox + s.f0, s.h_out + 2 * (size_t)c, sizeof(double) * s.nf);
    if (j.status) memcpy(j.status + s.f0, s.h_st, sizeof(int) * s.nf);
    if (j.ntri) memcpy(j.ntri + s.f0, s.h_st + c, sizeof(int) * s.nf);
    if (j.ncorr) memcpy(j.ncorr + s.f0, s.h_st + 2 * (size_t)c, sizeof(int) * s.nf);
    s.busy = false;
    return GTA_B200_OK;
}

int submit(Slot& s, const Job& j, const Shape& shp, int f0, int nf, bool src_pinned) {
    static int trace_slot; static std::once_flag trace_once;
    const bool trace = env_once("GTA_B200_TRACE", 0, &trace_slot, &trace_once) != 0;
    const auto t_begin = std::chrono::steady_clock::now();
    const size_t frame_elems = (size_t)j.natoms * j.in_dim, frame_bytes = frame_elems * sizeof(double);
    s.f0 = f0; s.nf = nf;
    // stage coordinates
    const double* h_src;
    if (j.frames) {
        // gather of one heap block per frame (the reference's layout, extern/gkut/src/gkut_io.c:37) into the pinned
        // chunk; at > 200k frames/s per GPU this is several GB/s of memcpy, so it is spread over the host cores
        const int ht = j.host_threads > 0 ? j.host_threads : std::max(1u, std::thread::hardware_concurrency());
#pragma omp parallel for schedule(static) num_threads(ht) if (nf >= 256)
        for (int f = 0; f < nf; ++f) memcpy(s.h_in + (size_t)f * frame_elems, j.frames[f0 + f], frame_bytes);
        h_src = s.h_in;
    } else if (src_pinned) {
        h_src = j.xyz + (size_t)f0 * frame_elems;              // caller's buffer is already page-locked
    } else {
        const int ht = j.host_threads > 0 ? j.host_threads : std::max(1u, std::thread::hardware_concurrency());
        const long long blocks = std::max(1, std::min(ht, nf / 256));
#pragma omp parallel for schedule(static) num_threads(ht) if (blocks > 1)
        for (long long b = 0; b < blocks; ++b) {
            const size_t lo = (size_t)nf * b / blocks, hi = (size_t)nf * (b + 1) / blocks;
            memcpy(s.h_in + lo * frame_elems, j.xyz + ((size_t)f0 + lo) * frame_elems, frame_bytes * (hi - lo));
        }
        h_src = s.h_in;
    }
    if (trace) fprintf(stderr, "[gta_b200] chunk %d+%d: host staging %.1f ms\n", f0, nf, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count());
    CK(cudaMemcpyAsync(s.d_in, h_src, frame_bytes * nf, cudaMemcpyHostToDevice, s.st));
    if (j.box) {
        memcpy(s.h_box, j.box + (size_t)f0 * j.box_stride, sizeof(double) * j.box_stride * nf);
        CK(cudaMemcpyAsync(s.d_box, s.h_box, sizeof(double) * j.box_stride * nf, cudaMemcpyHostToDevice, s.st));
    }
    const int c = s.chunk;
    gt::KParams p;
    p.xyz = s.d_in; p.in_dim = j.in_dim; p.box = j.box ? s.d_box : nullptr; p.box_stride = j.box_stride;
    p.by_off = (j.box_stride == 9) ? 4 : 1;
    p.nframes = nf; p.natoms = j.natoms; p.cap = shp.cap; p.cap_p2 = shp.cap_p2; p.espace = j.espace; p.flags = j.flags;
    p.area = s.d_out; p.area2d = s.d_out + c; p.areabox = s.d_out + 2 * (size_t)c;
    p.status = s.d_st; p.ntri = s.d_st + c; p.ncorr = s.d_st + 2 * (size_t)c;
    p.tri = j.tri ? s.d_tri : nullptr; p.tri_cap = j.tri_cap;
    p.corr = j.corr ? s.d_corr : nullptr; p.corr_cap = j.corr_cap;
    p.work = s.d_work; p.coop_nodes = coop_nodes_for(shp.threads); p.dbg = g_dbg; p.ws = gt::LpfWs{};
    CK(cudaMemsetAsync(s.d_work, 0, sizeof(unsigned int), s.st));
    CK(cudaMemsetAsync(s.d_st, 0, sizeof(int) * 3 * (size_t)c, s.st));
    CK(cudaEventRecord(s.k0, s.st));
    const bool star = use_star(p, shp);
    int rc = (star || use_lpf(nf, shp)) ? launch_lpf(p, shp, s.st, star) : launch(p, shp, s.st);
    if (rc) return rc;
    CK(cudaEventRecord(s.k1, s.st));
    CK(cudaMemcpyAsync(s.h_out, s.d_out, sizeof(double) * 3 * (size_t)c, cudaMemcpyDeviceToHost, s.st));
    CK(cudaMemcpyAsync(s.h_st, s.d_st, sizeof(int) * 3 * (size_t)c, cudaMemcpyDeviceToHost, s.st));
    if (j.tri) CK(cudaMemcpyAsync(j.tri + (size_t)f0 * j.tri_cap * 3, s.d_tri, sizeof(int) * 3 * (size_t)j.tri_cap * nf, cudaMemcpyDeviceToHost, s.st));
    if (j.corr) CK(cudaMemcpyAsync(j.corr + (size_t)f0 * j.corr_cap * 3, s.d_corr, sizeof(double) * 3 * (size_t)j.corr_cap * nf, cudaMemcpyDeviceToHost, s.st));
    s.busy = true;
    return GTA_B200_OK;
}

int run_job(Job j, int device, int nstreams) {
    g_err.clear();
    if (!have_device()) return fail(GTA_B200_E_NO_DEVICE, "no CUDA device: libgtessla_b200 has no CPU path");
    if (j.nframes < 0 || j.natoms < 0 || (j.in_dim != 2 && j.in_dim != 3) || (!j.xyz && !j.frames))
        return fail(GTA_B200_E_ARG, "bad argument");
    if ((j.flags & GTA_B200_CORRECT) && (!j.box || !(j.espace > 0.0)))
        return fail(GTA_B200_E_ARG, "-corr needs the box and espace > 0");
    g_last_kernel_ms = 0.0; g_last_launches = 0;
    if (j.nframes == 0) return GTA_B200_OK;
    if (device >= 0) CK(cudaSetDevice(device));
    CK(cudaGetDevice(&device));
    j.max_points = gta_b200_max_points_for(j.box, j.box_stride, j.nframes, j.natoms, j.espace, j.flags);
    Shape shp = shape_for(j.max_points, j.natoms);
    if (shp.smem > (size_t)smem_limit() || gt::pool_edges(shp.cap) >= 32760)
        return fail(GTA_B200_E_TOO_LARGE, "frame too large for the shared-memory kernel (see gta_b200_max_points)");

    bool src_pinned = false;
    if (j.xyz) {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, j.xyz) == cudaSuccess) src_pinned = (at.type == cudaMemoryTypeHost);
        else cudaGetLastError();
    }
    nstreams = std::max(1, std::min(nstreams <= 0 ? 3 : nstreams, 8));
    const size_t frame_bytes = (size_t)j.natoms * j.in_dim * sizeof(double);
    // chunk: enough frames for several waves of persistent CTAs, bounded to ~64 MB of coordinates
    long long chunk = (j.nframes + nstreams - 1) / nstreams;
    chunk = std::min<long long>(chunk, std::max<long long>(256, (64ll << 20) / (long long)std::max<size_t>(frame_bytes, 1)));
    chunk = std::max<long long>(chunk, 1);
    gt::KParams probe; probe.natoms = j.natoms; probe.in_dim = j.in_dim; probe.tri = j.tri;
    const bool star = use_star(probe, shp);
    if (star) {
        // star path: a launch has no long critical path (a frame is a CTA for ~0.1 ms), so the chunks are sized for the overlap
        // of the copies with the kernels: several chunks of a few thousand frames in flight on three streams
        static int sc_slot; static std::once_flag sc_once;
        const long long target = std::max(256, env_once("GTA_B200_STAR_CHUNK", 8192, &sc_slot, &sc_once));
        const long long nchunks = std::max<long long>(1, (j.nframes + target - 1) / target);
        chunk = (j.nframes + nchunks - 1) / nchunks;
        nstreams = (int)std::min<long long>(3, nchunks);
    } else if (use_lpf(j.nframes, shp)) {
        // throughput path: a launch lasts at least one frame's critical path (~2,800 rounds), so launches want >= 10^4
        // frames; up to three chunks in flight so that the copies of one overlap the kernels of another
        const long long target = std::max<long long>(4096, std::min<long long>(25000, (1ll << 30) / (long long)std::max<size_t>(frame_bytes, 1)));
        long long nchunks = (j.nframes + target - 1) / target;
        // a batch of one or two launches' worth that has to be staged by the host first (per-frame blocks, pageable memory):
        // cut it in pieces of >= chunk_min frames, so that the staging of the second piece runs under the kernels of the first
        // (measured on one GPU, 12,500 / 25,000 frames in per-frame blocks: 227 -> 238 k and 242 -> 264 k frames/s; from a pinned
        // buffer, where only the copy can overlap, one launch is as fast or faster: 268 k against 260 k)
        static int cm_slot; static std::once_flag cm_once;
        const long long chunk_min = std::max(1024, env_once("GTA_B200_CHUNK_MIN", 6000, &cm_slot, &cm_once));
        if (j.frames || !src_pinned) nchunks = std::max(nchunks, std::min<long long>(3, j.nframes / chunk_min));
        chunk = (j.nframes + nchunks - 1) / nchunks;
        nstreams = (int)std::min<long long>(3, nchunks);
    }
    if (j.tri) {                                              // (parity tests) triangle lists: bound the device buffer of a chunk to 512 MB
        const long long fit = std::max<long long>(1, (512ll << 20) / (12ll * std::max(j.tri_cap, 1)));
        if (chunk > fit) { chunk = fit; nstreams = std::min(nstreams, 2); }
    }
    if (device < 0 || device >= 64) return fail(GTA_B200_E_ARG, "device index out of range");
    SlotCache& g_cache = g_caches[device];
    std::lock_guard<std::mutex> cache_lock(g_cache.mu);
    if ((int)g_cache.slots.size() < nstreams) g_cache.slots.resize(nstreams);
    for (int i = 0; i < nstreams; ++i) {
        int rc = slot_ensure(g_cache.slots[i], (int)chunk, frame_bytes, std::max(j.box_stride, 1),
                             j.tri ? (size_t)chunk * j.tri_cap * 3 : 0, j.corr ? (size_t)chunk * j.corr_cap * 3 : 0,
                             !src_pinned || j.frames != nullptr);
        if (rc) return rc;
    }
    double kms = 0.0; int launches = 0, rc = GTA_B200_OK, si = 0;
    for (int f0 = 0; f0 < j.nframes && rc == GTA_B200_OK; f0 += (int)chunk) {
        Slot& s = g_cache.slots[si];
        si = (si + 1) % nstreams;
        rc = drain(s, j, &kms);
        if (rc) break;
        int nf = std::min<long long>(chunk, j.nframes - f0);
        rc = submit(s, j, shp, f0, nf, src_pinned);
        launches += star ? 4 : (use_lpf(nf, shp) ? 3 : 1);                // (stars +) pre-pass + walker kernel + area kernel, or the fused kernel
    }
    for (int i = 0; i < nstreams; ++i) { int r2 = drain(g_cache.slots[i], j, &kms); if (rc == GTA_B200_OK) rc = r2; }
    g_last_kernel_ms = kms; g_last_launches = launches;
    return rc;
}

}  // namespace

extern "C" {

int gta_b200_tessellate_batch(const double* xyz, int in_dim, const double* box, int box_stride,
                              int nframes, int natoms, double espace, unsigned flags,
                              int device, int nstreams,
                              double* area, double* area2d, double* areabox, int* status,
                              int* tri, int tri_cap, int* ntri,
                              double* corr, int corr_cap, int* ncorr) {
    Job j;
    j.xyz = xyz; j.in_dim = in_dim; j.box = box; j.box_stride = box_stride; j.nframes = nframes; j.natoms = natoms;
    j.espace = espace; j.flags = flags; j.area = area; j.area2d = area2d; j.areabox = areabox; j.status = status;
    j.tri = tri; j.tri_cap = tri_cap; j.ntri = ntri; j.corr = corr; j.corr_cap = corr_cap; j.ncorr = ncorr;
    return run_job(j, device, nstreams);
}

static int run_sharded(const Job& base, int ngpus, int nstreams) {
    int avail = gta_b200_device_count();
    if (avail < 1) return fail(GTA_B200_E_NO_DEVICE, "no CUDA device: libgtessla_b200 has no CPU path");
    ngpus = std::max(1, std::min(ngpus <= 0 ? avail : ngpus, avail));
    const int per = (base.nframes + ngpus - 1) / ngpus;     // contiguous ranges [g*per, (g+1)*per)
    std::vector<int> rcs(ngpus, GTA_B200_OK);
    std::vector<std::string> errs(ngpus);
    std::vector<double> kms(ngpus, 0.0);
    std::vector<int> ln(ngpus, 0);
    std::vector<std::thread> th;
    for (int g = 0; g < ngpus; ++g) {
        th.emplace_back([&, g]() {
            int f0 = g * per, f1 = std::min(base.nframes, f0 + per);
            if (f0 >= f1) return;
            Job j = base;
            {   // host staging threads are shared between the per-device threads
                const int hw = (int)std::max(1u, std::thread::hardware_concurrency());
                j.host_threads = std::max(1, (base.host_threads > 0 ? base.host_threads : hw) / ngpus);
            }
            const size_t fe = (size_t)j.natoms * j.in_dim;
            if (j.xyz) j.xyz += (size_t)f0 * fe;
            if (j.frames) j.frames += f0;
            if (j.box) j.box += (size_t)f0 * j.box_stride;
            j.nframes = f1 - f0;
            if (j.area) j.area += f0;
            if (j.area2d) j.area2d += f0;
            if (j.areabox) j.areabox += f0;
            if (j.status) j.status += f0;
            rcs[g] = run_job(j, g, nstreams);
            errs[g] = g_err; kms[g] = g_last_kernel_ms; ln[g] = g_last_launches;
        });
    }
    for (auto& t : th) t.join();
    g_last_kernel_ms = *std::max_element(kms.begin(), kms.end());
    g_last_launches = 0; for (int v : ln) g_last_launches += v;
    for (int g = 0; g < ngpus; ++g) if (rcs[g]) return fail(rcs[g], errs[g]);
    return GTA_B200_OK;
}

int gta_b200_tessellate_multi(const double* xyz, int in_dim, const double* box, int box_stride,
                              int nframes, int natoms, double espace, unsigned flags,
                              int ngpus, int nstreams,
                              double* area, double* area2d, double* areabox, int* status) {
    Job j;
    j.xyz = xyz; j.in_dim = in_dim; j.box = box; j.box_stride = box_stride; j.nframes = nframes; j.natoms = natoms;
    j.espace = espace; j.flags = flags; j.area = area; j.area2d = area2d; j.areabox = areabox; j.status = status;
    return run_sharded(j, ngpus, nstreams);
}

int gta_b200_tessellate_frames(const double* const* x, const double* box9, int nframes, int natoms,
                               double espace, unsigned flags, int ngpus, int nthreads,
                               double* area, double* area2d, double* areabox, int* status) {
    Job j;
    j.frames = x; j.in_dim = 3; j.box = box9; j.box_stride = 9; j.nframes = nframes; j.natoms = natoms;
    j.espace = espace; j.flags = flags; j.area = area; j.area2d = area2d; j.areabox = areabox; j.status = status;
    j.host_threads = nthreads;
    return run_sharded(j, ngpus <= 0 ? 1 : ngpus, 4);
}

static int signs_common(const double* pts, int n, int* out, int device, int width) {
    if (!have_device()) return fail(GTA_B200_E_NO_DEVICE, "no CUDA device: libgtessla_b200 has no CPU path");
    if (n <= 0) return GTA_B200_OK;
    if (device >= 0) CK(cudaSetDevice(device));
    double* d = nullptr; int* o = nullptr;
    CK(cudaMalloc((void**)&d, sizeof(double) * width * (size_t)n));
    CK(cudaMalloc((void**)&o, sizeof(int) * (size_t)n));
    CK(cudaMemcpy(d, pts, sizeof(double) * width * (size_t)n, cudaMemcpyHostToDevice));
    if (width == 6) gt::orient2d_signs_kernel<<<(n + 127) / 128, 128>>>(d, n, o);
    else gt::incircle_signs_kernel<<<(n + 127) / 128, 128>>>(d, n, o);
    CK(cudaGetLastError());
    CK(cudaMemcpy(out, o, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost));
    cudaFree(d); cudaFree(o);
    return GTA_B200_OK;
}
// FP64 peak of this device, measured: independent DFMA chains in registers (the denominator of the FP64 roofline;
// MEASURED_PEAKS.json has no FP64 figure). Returns TFLOP/s (2 flops per DFMA), <= 0 on error.
namespace gt {
__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 4
    for (int i = 0; i < iters; ++i) {
        x0 = __fma_rn(x0, a, b); x1 = __fma_rn(x1, a, b); x2 = __fma_rn(x2, a, b); x3 = __fma_rn(x3, a, b);
        x4 = __fma_rn(x4, a, b); x5 = __fma_rn(x5, a, b); x6 = __fma_rn(x6, a, b); x7 = __fma_rn(x7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}
}  // namespace gt
double gta_b200_fp64_peak_tflops(int device) {
    if (!have_device()) return -1.0;
    if (device >= 0 && cudaSetDevice(device) != cudaSuccess) return -1.0;
    int sms = 0, dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1.0;
    const int blocks = sms * 8, threads = 256, iters = 1 << 15;
    double* out = nullptr; cudaEvent_t e0, e1;
    if (cudaMalloc((void**)&out, sizeof(double) * blocks * threads) != cudaSuccess) return -1.0;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {                          // first repetition warms the clocks up
        cudaEventRecord(e0);
        gt::dfma_peak_kernel<<<blocks, threads>>>(out, iters, 0.999999, 1e-9);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { best = -1.0; break; }
        float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1);
        const double tf = 2.0 * 8.0 * (double)iters * blocks * threads / (ms * 1e-3) / 1e12;
        if (rep && tf > best) best = tf;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
    return best;
}

// Debug hook (not in the public header): allocate / read the phase timers of GT_PHASE_TIMING builds.
int gta_b200_dbg_timers(unsigned long long* out64, int reset) {       // [0..31] fused kernel, [32..63] lane-per-frame kernel
    if (!g_dbg) { if (cudaMalloc((void**)&g_dbg, 64 * 8) != cudaSuccess) return -1; cudaMemset(g_dbg, 0, 64 * 8); }
    if (out64) cudaMemcpy(out64, g_dbg, 64 * 8, cudaMemcpyDeviceToHost);
    if (reset) cudaMemset(g_dbg, 0, 64 * 8);
    return 0;
}
int gta_b200_orient2d_signs(const double* pts, int n, int* sign_out, int device) { return signs_common(pts, n, sign_out, device, 6); }
int gta_b200_incircle_signs(const double* pts, int n, int* sign_out, int device) { return signs_common(pts, n, sign_out, device, 8); }

}  // extern "C"
