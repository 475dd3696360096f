// gta_large.cu -- one large frame (BASELINE cfg 4: 10^6 points): the same divide-and-conquer as the
// batched kernel, run level by level over GLOBAL memory with 32-bit indices.
//
// Replaces dtriangulate (reference src/delaunay_tri.c:413-507) for point sets that do not fit the
// shared-memory kernel. The reference handles such a frame serially (one core, 5.5 s for 10^6 points).
// Here every level of the reference's recursion tree (:535-539, same split index) is one launch:
// the lower levels hold up to n/2 independent merges (one thread each, dt_core.cuh::merge -- exactly the
// code the batched kernel and the host emulation run), the upper levels hold few long merges and give
// each its own warp so that they do not serialise behind each other. Because it is the reference's own
// decision sequence, the result is the reference's triangulation on EVERY input, including exactly
// cocircular point sets where an insert-and-flip triangulator would pick different diagonals; that is
// why a flip-based kernel was not used for this path.
//
// Sorting is two stable cub radix passes (by y, then by x) on order-preserving keys, followed by the
// reference comparator's near-tie rule (delaunay_tri.c:115-123) on runs of |dx| < 1e-12.
// Freed edges are recycled only inside the merge that freed them (no shared free stack in this mode):
// the pool is sized for every edge ever created.
#include <cuda_runtime.h>
#include <cub/cub.cuh>
#include <algorithm>
#include <cstdio>
#include <limits>
#include <cmath>
#include <mutex>
#include <string>
#include <vector>
#include "dt_core.cuh"
#include "dt_coop.cuh"
#include "gta_large_star.cuh"
#include "../../include/gta_b200.h"

namespace gt {
typedef FrameT<uint32_t> FrameL;

__device__ __forceinline__ unsigned long long okey64(double v) {   // order-preserving bits
    unsigned long long b = (unsigned long long)__double_as_longlong(v + 0.0);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

__global__ void large_keys(const double* pts, int n, unsigned long long* kx, unsigned long long* ky, uint32_t* idx, uint32_t* err) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double x = pts[2 * (size_t)i], y = pts[2 * (size_t)i + 1];
    if (!isfinite(x) || !isfinite(y)) atomicOr(err, 0x100u);
    kx[i] = okey64(x); ky[i] = okey64(y); idx[i] = (uint32_t)i;
}
__global__ void large_gather_key(const unsigned long long* kx, const uint32_t* idx, int n, unsigned long long* out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = kx[idx[i]];
}
__global__ void large_gather_pts(const double* pts, const uint32_t* perm, int n, Pt* pt, uint32_t* head) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t o = perm[i];
    pt[i].x = pts[2 * (size_t)o]; pt[i].y = pts[2 * (size_t)o + 1];
    head[i] = FrameL::NIL;
}
// Near-tie runs (see gta_kernel.cuh) and the duplicate check of the reference's contract.
__global__ void large_fix_ties(Pt* pt, uint32_t* perm, int n, uint32_t* err) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    auto tie = [&](int a, int b) { double d = pt[a].x - pt[b].x; return d < 1e-12 && d > -1e-12; };
    if ((i == 0 || !tie(i, i - 1)) && i + 1 < n && tie(i + 1, i)) {
        int e = i + 1; bool inexact = false;
        while (e < n && tie(e, e - 1)) { inexact |= (pt[e].x != pt[e - 1].x); ++e; }
        if (inexact) {
            if (!tie(e - 1, i)) atomicOr(err, 0x200u);
            for (int a = i + 1; a < e; ++a) {
                Pt pa = pt[a]; uint32_t oa = perm[a]; int b = a - 1;
                while (b >= i && pt[b].y > pa.y) { pt[b + 1] = pt[b]; perm[b + 1] = perm[b]; --b; }
                pt[b + 1] = pa; perm[b + 1] = oa;
            }
        }
    }
}
__global__ void large_check_dups(const Pt* pt, int n, uint32_t* err) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 1 || i >= n) return;
    double dx = pt[i].x - pt[i - 1].x, dy = pt[i].y - pt[i - 1].y;
    if (dx < 1e-12 && dx > -1e-12 && dy < 1e-12 && dy > -1e-12) atomicOr(err, 0x400u);
}

// One level of the recursion tree. lanes_per_node = 1: node k is thread k (lower levels, many short merges);
// lanes_per_node = 32: node k is warp k (upper levels: lanes of one warp that walk different merges would
// serialise) -- with coop, the warp evaluates every predicate a merge step could ask for in one round
// (dt_coop.cuh::coop_merge, the same code as the top of the tree in the batched kernel); without, lane 0 walks the
// merge alone.
__global__ void __launch_bounds__(128) large_level(FrameL f, int depth, int nodes, int lanes_per_node, int coop) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long k = t / lanes_per_node;
    if (k >= nodes) return;
    Lane ln; ln.free_head = FrameL::NIL;
    if (lanes_per_node == 32 && coop) {
        int ia, ib;
        if (!tree_node(f.n, depth, (int)k, &ia, &ib) || ib - ia < 1) return;       // (uniform over the warp)
        if (ib - ia < 3) { if ((t & 31) == 0) leaf(f, ln, ia, ib); }
        else coop_merge(f, ln, (ia + ib) / 2);
        return;
    }
    if ((t % lanes_per_node) != 0) return;
    process_node(f, ln, depth, (int)k);
}

__global__ void large_count(FrameL f, uint32_t* cnt) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= f.n) return;
    cnt[i] = (uint32_t)vertex_triangles(f, i, [](int, int) {});
}
__global__ void large_emit(FrameL f, const uint32_t* offs, const uint32_t* perm, int* tri, int tri_cap) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= f.n) return;
    uint32_t o = offs[i];
    const int oi = (int)perm[i];
    vertex_triangles(f, i, [&](int n1, int n2) {
        if ((int)o < tri_cap) { tri[3 * (size_t)o] = oi; tri[3 * (size_t)o + 1] = (int)perm[n1]; tri[3 * (size_t)o + 2] = (int)perm[n2]; }
        ++o;
    });
}
// Area epilogue of a large frame (reference src/gta_tri.c:368-410): per sorted vertex, the sum over the triangles it owns
// in emission order; the per-vertex sums are then added up by a fixed-shape reduction (cub).
__global__ void large_area(FrameL f, const uint32_t* perm, const double* xyz, int want2d, double* a3, double* a2) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= f.n) return;
    const double* pi = xyz + 3 * (size_t)perm[i];
    double s3 = 0.0, s2 = 0.0;
    vertex_triangles(f, i, [&](int n1, int n2) {
        const double* p1 = xyz + 3 * (size_t)perm[n1]; const double* p2 = xyz + 3 * (size_t)perm[n2];
        s3 += tri_area3(pi[0], pi[1], pi[2], p1[0], p1[1], p1[2], p2[0], p2[1], p2[2]);
        if (want2d) s2 += tri_area3(pi[0], pi[1], 0.0, p1[0], p1[1], 0.0, p2[0], p2[1], 0.0);
    });
    a3[i] = s3; a2[i] = s2;
}
__global__ void large_xy(const double* xyz, int n, double* pts) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { pts[2 * (size_t)i] = xyz[3 * (size_t)i]; pts[2 * (size_t)i + 1] = xyz[3 * (size_t)i + 1]; }
}
}  // namespace gt

namespace {
thread_local std::string g_lerr;
thread_local double g_large_ms = 0.0;
thread_local int g_large_star = 0;
}
extern "C" int gta_b200_large_last_engine(void) { return g_large_star; }
extern "C" const char* gta_b200_large_error(void) { return g_lerr.c_str(); }
extern "C" double gta_b200_large_last_ms(void) { return g_large_ms; }

extern "C" int gta_b200_get_path(void);

// One large frame through the star engine (gta_large_star.cuh). d_pts: device [n][2]; d_xyz: device [n][3] or null.
// Returns 0: certified, results written (triangles in cell order); 1: not certified (a tie, duplicates, count rule) -- the
// caller runs the level-by-level path; < 0: error (g_lerr set).
// Buffers of the large-frame star path come from the device's stream-ordered pool with a release threshold that keeps
// them between calls: the ~20 cudaMalloc / cudaFree of a call were 10 ms of its 44.
static cudaError_t ls_alloc(void** p, size_t bytes) {
    static std::once_flag once[64];
    int dev = 0; cudaGetDevice(&dev);
    std::call_once(once[dev & 63], [dev] {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) { unsigned long long keep = ~0ull; cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep); }
    });
    // Sizes are rounded up (64 KB below 1 MB, 2 MB above): the frames of a trajectory differ by a few points, and blocks of
    // slightly different sizes are not reused -- the pool then maps new memory on most calls (sporadic stalls of 20-600 ms
    // per frame were measured on a trajectory of 10^5-atom frames).
    const size_t gran = bytes >= (1u << 20) ? (size_t)2 << 20 : (size_t)64 << 10;
    bytes = ((bytes ? bytes : 1) + gran - 1) / gran * gran;
    return cudaMallocAsync(p, bytes, 0);
}
static void ls_free(void* p) { if (p) cudaFreeAsync(p, 0); }

static int large_star(const double* d_pts, const double* d_xyz, int n, unsigned flags, int* tri, int tri_cap, int* ntri, double* area3, double* area2) {
    using namespace gt;
#define SCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { g_lerr = std::string(#call) + ": " + cudaGetErrorString(e_); rc = GTA_B200_E_CUDA; goto out; } } while (0)
    int rc = 1;
    unsigned long long* d_key = nullptr; uint32_t *d_ctl = nullptr, *d_cell = nullptr, *d_cell2 = nullptr, *d_idx = nullptr, *d_idx2 = nullptr, *d_cstart = nullptr;
    uint32_t *d_ll = nullptr, *d_cl = nullptr, *d_cnt = nullptr, *d_off = nullptr, *d_pairs = nullptr, *d_ovf = nullptr, *d_def = nullptr;
    SPt* d_spt = nullptr; SPtF* d_sptf = nullptr; double *d_sz = nullptr, *d_a3 = nullptr, *d_a2 = nullptr, *d_sum = nullptr, *d_rowx = nullptr; int* d_tri = nullptr;
    void* d_tmp = nullptr; size_t tb = 0, tb2 = 0, tb3 = 0;
    const int B = 256, G = (n + B - 1) / B;
    const bool want_area = d_xyz != nullptr && area3 != nullptr, want_tri = tri != nullptr && tri_cap > 0;
    unsigned long long hkey[4] = {~0ull, 0ull, ~0ull, 0ull};
    uint32_t hctl[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int dev = 0, sms = 0;
    SCK(cudaGetDevice(&dev));
    SCK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    {
        SCK(ls_alloc((void**)&d_key, 32)); SCK(ls_alloc((void**)&d_ctl, 32));
        SCK(cudaMemcpy(d_key, hkey, 32, cudaMemcpyHostToDevice)); SCK(cudaMemset(d_ctl, 0, 32));
        ls_bbox<<<sms * 4, 256>>>(d_pts, n, d_key, d_ctl);
        SCK(cudaMemcpy(hkey, d_key, 32, cudaMemcpyDeviceToHost));
        SCK(cudaMemcpy(hctl, d_ctl, 16, cudaMemcpyDeviceToHost));
        if (hctl[0]) { rc = 1; goto out; }                                   // non-finite input: the other path reports it
        const double xmin = ls_unkey(hkey[0]), xmax = ls_unkey(hkey[1]), ymin = ls_unkey(hkey[2]), ymax = ls_unkey(hkey[3]);
        const StarDims dm = star_dims(n, xmin, xmax, ymin, ymax, std::max(n, 1), 1.0);
        const int ncell = dm.gx * dm.gy;
        SCK(ls_alloc((void**)&d_cell, 4 * (size_t)n)); SCK(ls_alloc((void**)&d_cell2, 4 * (size_t)n));
        SCK(ls_alloc((void**)&d_idx, 4 * (size_t)n)); SCK(ls_alloc((void**)&d_idx2, 4 * (size_t)n));
        SCK(ls_alloc((void**)&d_cstart, 4 * (size_t)(ncell + 2))); SCK(cudaMemset(d_cstart, 0, 4 * (size_t)(ncell + 2)));
        SCK(ls_alloc((void**)&d_spt, sizeof(SPt) * (size_t)n)); SCK(ls_alloc((void**)&d_sptf, sizeof(SPtF) * (size_t)n));
        if (want_area) { SCK(ls_alloc((void**)&d_sz, 8 * (size_t)n)); SCK(ls_alloc((void**)&d_a3, 8 * (size_t)n)); SCK(ls_alloc((void**)&d_a2, 8 * (size_t)n)); SCK(ls_alloc((void**)&d_sum, 16)); }
        SCK(ls_alloc((void**)&d_cnt, 4 * (size_t)n)); SCK(ls_alloc((void**)&d_off, 4 * (size_t)n));
        const uint32_t ovf_cap = (uint32_t)std::max(n / 4, 4096);
        if (want_tri) { SCK(ls_alloc((void**)&d_pairs, 8 * (size_t)LS_MAXOWN * n)); SCK(ls_alloc((void**)&d_ovf, 12 * (size_t)ovf_cap)); }
        // cells, stable sort by cell (order inside a cell = input order), cell starts
        ls_cells<<<G, B>>>(d_pts, n, xmin, ymin, dm.sx, dm.sy, dm.gx, dm.gy, d_cell, d_idx, d_cstart + 1);
        {
            // strongly clustered input (a quarter of the points in cells of more than 24): not for the grid
            uint32_t dense = 0;
            ls_dense<<<(ncell + 255) / 256, 256>>>(d_cstart + 1, ncell, 24u, d_ctl + 6);
            SCK(cudaMemcpy(&dense, d_ctl + 6, 4, cudaMemcpyDeviceToHost));
            if ((long long)dense * 4 > (long long)n) { rc = 1; goto out; }
        }
        int bits = 1; while ((1ll << bits) < ncell && bits < 32) ++bits;
        cub::DeviceRadixSort::SortPairs(nullptr, tb, d_cell, d_cell2, d_idx, d_idx2, n, 0, bits);
        cub::DeviceScan::InclusiveSum(nullptr, tb2, d_cstart + 1, d_cstart + 1, ncell);
        cub::DeviceScan::ExclusiveSum(nullptr, tb3, d_cnt, d_off, n);
        tb = std::max(tb, std::max(tb2, tb3));
        if (want_area) { size_t t4 = 0; cub::DeviceReduce::Sum(nullptr, t4, d_a3, d_sum, n); tb = std::max(tb, t4); }
        SCK(ls_alloc(&d_tmp, tb));
        cub::DeviceRadixSort::SortPairs(d_tmp, tb, d_cell, d_cell2, d_idx, d_idx2, n, 0, bits);
        cub::DeviceScan::InclusiveSum(d_tmp, tb, d_cstart + 1, d_cstart + 1, ncell);          // cstart[c + 1] = end of cell c
        ls_gather<<<G, B>>>(d_pts, d_xyz, d_idx2, n, d_spt, d_sptf, d_sz);
        SCK(ls_alloc((void**)&d_rowx, 16 * (size_t)dm.gy));
        ls_rowx<<<(unsigned)((32 * (size_t)dm.gy + 255) / 256), 256>>>(d_spt, d_cstart, dm.gx, dm.gy, d_rowx);
        // stars
        StarGrid<uint32_t, true> g;
        g.pt = d_spt; g.ptf = d_sptf; g.cstart = d_cstart; g.pt_s = g.ptf_s = g.cstart_s = 0;
        g.kf = (float)(2.5 * 5.9604644775390625e-08 * std::max(std::max(fabs(xmin), fabs(xmax)), std::max(fabs(ymin), fabs(ymax)))) * 1.0001f;
        g.n = n; g.gx = dm.gx; g.gy = dm.gy; g.x0 = xmin; g.y0 = ymin; g.sx = dm.sx; g.sy = dm.sy;
        g.xmin = xmin; g.xmax = xmax; g.ymin = ymin; g.ymax = ymax;
        g.cx0 = g.cy0 = 0; g.ggx = dm.gx; g.ggy = dm.gy; g.open = 0; g.rowx = d_rowx;
        int occ = 0;
        SCK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ls_star_kernel, 128, 0));
        const int blocks = std::max(1, std::min((n + 127) / 128, std::max(1, occ) * sms));
        const size_t nt = (size_t)blocks * 128;
        SCK(ls_alloc((void**)&d_ll, 4 * (size_t)LS_LCAP * nt)); SCK(ls_alloc((void**)&d_cl, 4 * (size_t)LS_CCAP * nt));
        SCK(cudaMemset(d_ctl, 0, 32));
        static const bool trace = getenv("GTA_B200_TRACE") != nullptr;
        static const int defer_per = getenv("GTA_B200_DEFER_PER") ? std::max(1, std::min(32, atoi(getenv("GTA_B200_DEFER_PER")))) : 1;   // deferred points per warp
        static const bool tiles = !getenv("GTA_B200_LARGE_TILES") || atoi(getenv("GTA_B200_LARGE_TILES")) != 0;     // experiments: 0 = every point from global memory
        unsigned long long* d_clk = nullptr;
        if (trace) { ls_alloc((void**)&d_clk, 8 * (size_t)n); cudaMemset(d_clk, 0, 8 * (size_t)n); }      // (a chunk may be one deferred point)
        cudaEvent_t t0 = nullptr, t1 = nullptr;
        if (trace) { cudaEventCreate(&t0); cudaEventCreate(&t1); cudaEventRecord(t0); }
        if (tiles) {
            // tiles of the grid staged in shared memory; what they defer goes through ls_star_kernel afterwards
            const int tiles_x = (dm.gx + LT_TW - 1) / LT_TW, tiles_y = (dm.gy + LT_TH - 1) / LT_TH, ntiles = tiles_x * tiles_y;
            const size_t smem = LtSmem().total;
            int tocc = 0;
            SCK(ls_alloc((void**)&d_def, 4 * (size_t)n));
            SCK(cudaFuncSetAttribute(ls_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            SCK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&tocc, ls_tile_kernel, LT_T, smem));
            ls_tile_kernel<<<std::max(1, std::min(ntiles, std::max(1, tocc) * sms)), LT_T, smem>>>(g, tiles_x, ntiles, d_sz, (flags & GTA_B200_2D) ? 1 : 0, d_cnt, d_pairs, d_a3, d_a2, d_ctl, d_def);
            SCK(cudaGetLastError());
            if (trace) { cudaEvent_t tm; cudaEventCreate(&tm); cudaEventRecord(tm); cudaEventSynchronize(tm); float ms = 0.f; cudaEventElapsedTime(&ms, t0, tm); fprintf(stderr, "[gta_b200] tile kernel %.2f ms (%d tiles)\n", ms, ntiles); cudaEventDestroy(tm); }
            ls_star_kernel<<<blocks, 128>>>(g, d_ll, d_cl, d_sz, (flags & GTA_B200_2D) ? 1 : 0, d_cnt, d_pairs, d_ovf, ovf_cap, d_a3, d_a2, d_ctl, d_clk, d_def, d_ctl + 4, defer_per);
        } else
        ls_star_kernel<<<blocks, 128>>>(g, d_ll, d_cl, d_sz, (flags & GTA_B200_2D) ? 1 : 0, d_cnt, d_pairs, d_ovf, ovf_cap, d_a3, d_a2, d_ctl, d_clk, nullptr, nullptr, 32);
        SCK(cudaGetLastError());
        if (trace && d_clk) {
            const int nch = tiles ? std::min(n, 4096) : (n + 31) / 32;
            std::vector<unsigned long long> hc(nch);
            cudaMemcpy(hc.data(), d_clk, 8 * (size_t)nch, cudaMemcpyDeviceToHost);
            std::vector<int> order(nch); for (int i = 0; i < nch; ++i) order[i] = i;
            std::sort(order.begin(), order.end(), [&](int a, int b) { return hc[a] > hc[b]; });
            unsigned long long sum = 0; for (auto v : hc) sum += v;
            fprintf(stderr, "[gta_b200] chunk cycles: sum %.3g, median %llu, top:", (double)sum, hc[order[nch / 2]]);
            for (int i = 0; i < 8 && i < nch; ++i) fprintf(stderr, " ch %d (row %d): %.3g;", order[i], (order[i] * 32) / std::max(dm.gx, 1), (double)hc[order[i]]);
            fprintf(stderr, "\n");
        }
        if (trace) {
            cudaEventRecord(t1); cudaEventSynchronize(t1); float ms = 0.f; cudaEventElapsedTime(&ms, t0, t1);
            uint32_t c2[8]; cudaMemcpy(c2, d_ctl, 32, cudaMemcpyDeviceToHost);
            fprintf(stderr, "[gta_b200] large star: n %d grid %d x %d, %d blocks, star kernels %.2f ms, status bits 0x%x, open stars %u, deferred by the tiles %u\n", n, dm.gx, dm.gy, blocks, ms, c2[0], c2[1], c2[4]);
            cudaEventDestroy(t0); cudaEventDestroy(t1); ls_free(d_clk);
        }
        cub::DeviceScan::ExclusiveSum(d_tmp, tb, d_cnt, d_off, n);
        uint32_t last_off = 0, last_cnt = 0;
        SCK(cudaMemcpy(&last_off, d_off + (n - 1), 4, cudaMemcpyDeviceToHost));
        SCK(cudaMemcpy(&last_cnt, d_cnt + (n - 1), 4, cudaMemcpyDeviceToHost));
        SCK(cudaMemcpy(hctl, d_ctl, 16, cudaMemcpyDeviceToHost));
        const long long in_slots = (long long)last_off + last_cnt, novf = want_tri ? (long long)hctl[2] : 0;
        const long long total = in_slots + novf;
        // certified: no tie / duplicate / guard / list overflow anywhere, and the triangle count rule (delaunay_tri.c:608)
        if (hctl[0] != 0 || total != 2ll * (n - 1) - (long long)hctl[1]) { rc = 1; goto out; }
        if (ntri) *ntri = (int)total;
        if (want_area) {
            double hs[2] = {0.0, 0.0};
            cub::DeviceReduce::Sum(d_tmp, tb, d_a3, d_sum, n);
            cub::DeviceReduce::Sum(d_tmp, tb, d_a2, d_sum + 1, n);
            SCK(cudaMemcpy(hs, d_sum, 16, cudaMemcpyDeviceToHost));
            *area3 = hs[0];
            if (area2) *area2 = hs[1];
        }
        if (want_tri) {
            const int keep = (int)std::min<long long>(total, tri_cap);
            SCK(ls_alloc((void**)&d_tri, sizeof(int) * 3 * (size_t)std::max(keep, 1)));
            ls_emit<<<G, B>>>(d_cnt, d_off, d_pairs, d_idx2, n, d_tri, keep);
            if (novf > 0) ls_emit_overflow<<<(unsigned)((novf + 255) / 256), 256>>>(d_ovf, (int)novf, d_idx2, d_tri, (int)in_slots, keep);
            SCK(cudaGetLastError());
            SCK(cudaMemcpy(tri, d_tri, sizeof(int) * 3 * (size_t)keep, cudaMemcpyDeviceToHost));
            if (novf > 0 && in_slots < keep) {
                // the overflow list was filled through an atomic counter: put its triangles in a fixed order
                struct T3 { int a, b, c; };
                T3* t = reinterpret_cast<T3*>(tri) + in_slots;
                std::sort(t, t + (keep - in_slots), [](const T3& x, const T3& y) { return x.a != y.a ? x.a < y.a : (x.b != y.b ? x.b < y.b : x.c < y.c); });
            }
        }
        SCK(cudaDeviceSynchronize());
        rc = 0;
    }
out:
    ls_free(d_key); ls_free(d_ctl); ls_free(d_cell); ls_free(d_cell2); ls_free(d_idx); ls_free(d_idx2); ls_free(d_cstart);
    ls_free(d_ll); ls_free(d_cl); ls_free(d_cnt); ls_free(d_off); ls_free(d_pairs); ls_free(d_ovf); ls_free(d_spt); ls_free(d_sptf);
    ls_free(d_sz); ls_free(d_a3); ls_free(d_a2); ls_free(d_sum); ls_free(d_tri); ls_free(d_tmp); ls_free(d_def); ls_free(d_rowx);
    return rc;
#undef SCK
}

#define LCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { g_lerr = std::string(#call) + ": " + cudaGetErrorString(e_); rc = GTA_B200_E_CUDA; goto done; } } while (0)

// points: [n][2] raw 2-D points, or (xyz3 != nullptr) ignored and taken from xyz3 [n][3], whose z then enters the areas
static int large_run(const double* points, const double* xyz3, int npoints, int device, int* tri, int tri_cap, int* ntri, int* status,
                     unsigned flags, double* area3, double* area2) {
    using namespace gt;
    int rc = GTA_B200_OK;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) { cudaGetLastError(); g_lerr = "no CUDA device: libgtessla_b200 has no CPU path"; return GTA_B200_E_NO_DEVICE; }
    if ((!points && !xyz3) || npoints < 0 || (tri_cap > 0 && !tri)) { g_lerr = "bad argument"; return GTA_B200_E_ARG; }
    if (status) *status = 0;
    if (ntri) *ntri = 0;
    if (npoints < 2) { if (status) *status = GTA_B200_ST_TOO_FEW_POINTS; return GTA_B200_OK; }
    if (npoints > (1 << 27)) { g_lerr = "too many points for 32-bit half-edge indices"; return GTA_B200_E_TOO_LARGE; }
    if (device >= 0 && cudaSetDevice(device) != cudaSuccess) { g_lerr = "cudaSetDevice failed"; return GTA_B200_E_CUDA; }

    const int n = npoints;
    const uint32_t cap_e = (uint32_t)std::min<long long>(8ll * n + 1024, 0x7fffffff / 2);
    double* d_pts = nullptr; Pt* d_pt = nullptr; unsigned long long *d_k0 = nullptr, *d_k1 = nullptr, *d_ky = nullptr;
    uint32_t *d_i0 = nullptr, *d_i1 = nullptr, *d_head = nullptr, *d_dst = nullptr, *d_nxt = nullptr, *d_prv = nullptr, *d_ctl = nullptr, *d_cnt = nullptr, *d_off = nullptr;
    int* d_tri = nullptr; void* d_tmp = nullptr; size_t tmp_bytes = 0, tb2 = 0, tb3 = 0;
    double *d_xyz = nullptr, *d_a3 = nullptr, *d_a2 = nullptr, *d_sum = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    const int B = 256, G = (n + B - 1) / B;
    uint32_t hctl[4] = {0, 0, 0, 0};
    LCK(cudaEventCreate(&ev0)); LCK(cudaEventCreate(&ev1));
    LCK(ls_alloc((void**)&d_pts, sizeof(double) * 2 * (size_t)n));                 // (pool: no device-wide synchronisation per frame)
    if (xyz3) {
        LCK(ls_alloc((void**)&d_xyz, sizeof(double) * 3 * (size_t)n));
        LCK(cudaMemcpy(d_xyz, xyz3, sizeof(double) * 3 * (size_t)n, cudaMemcpyDefault));      // (host, or device: gta_b200_tessellate_large)
        large_xy<<<G, B>>>(d_xyz, n, d_pts);
    } else LCK(cudaMemcpy(d_pts, points, sizeof(double) * 2 * (size_t)n, cudaMemcpyHostToDevice));
    LCK(cudaEventRecord(ev0));
    {
        // The star engine first, when its output contract allows it: areas always; triangle lists only under GTA_B200_PATH=2
        // (they come in cell order, the kept dtriangulate promises the reference's emission order). Not certified -> below.
        const int mode = gta_b200_get_path();
        if (mode == 2 || (mode == 0 && !(tri && tri_cap > 0) && xyz3)) {
            const int sr = large_star(d_pts, d_xyz, n, flags, tri, tri_cap, ntri, area3, area2);
            if (sr < 0) { rc = sr; goto done; }
            if (sr == 0) {
                LCK(cudaEventRecord(ev1)); LCK(cudaEventSynchronize(ev1));
                float ms = 0.f; cudaEventElapsedTime(&ms, ev0, ev1); g_large_ms = ms;
                g_large_star = 1;
                goto done;
            }
        }
        g_large_star = 0;
    }
    // the level-by-level path's buffers (only now: a frame the stars certified never needs them)
    LCK(cudaMalloc((void**)&d_pt, sizeof(Pt) * (size_t)n));
    LCK(cudaMalloc((void**)&d_k0, 8 * (size_t)n)); LCK(cudaMalloc((void**)&d_k1, 8 * (size_t)n)); LCK(cudaMalloc((void**)&d_ky, 8 * (size_t)n));
    LCK(cudaMalloc((void**)&d_i0, 4 * (size_t)n)); LCK(cudaMalloc((void**)&d_i1, 4 * (size_t)n));
    LCK(cudaMalloc((void**)&d_head, 4 * (size_t)n));
    LCK(cudaMalloc((void**)&d_dst, 8 * (size_t)cap_e)); LCK(cudaMalloc((void**)&d_nxt, 8 * (size_t)cap_e)); LCK(cudaMalloc((void**)&d_prv, 8 * (size_t)cap_e));
    LCK(cudaMalloc((void**)&d_ctl, 16)); LCK(cudaMemset(d_ctl, 0, 16));
    LCK(cudaMalloc((void**)&d_cnt, 4 * (size_t)n)); LCK(cudaMalloc((void**)&d_off, 4 * (size_t)n));
    if (xyz3) { LCK(cudaMalloc((void**)&d_a3, 8 * (size_t)n)); LCK(cudaMalloc((void**)&d_a2, 8 * (size_t)n)); LCK(cudaMalloc((void**)&d_sum, 16)); }
    {
        // ---- sort: stable by y, then stable by x  => lexicographic (x, y)
        large_keys<<<G, B>>>(d_pts, n, d_k0, d_ky, d_i0, d_ctl + 2);
        cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_ky, d_k1, d_i0, d_i1, n);
        cub::DeviceScan::ExclusiveSum(nullptr, tb2, d_cnt, d_off, n);
        cub::DeviceReduce::Sum(nullptr, tb3, d_a3, d_sum, n);
        tmp_bytes = std::max(tmp_bytes, std::max(tb2, tb3));
        LCK(cudaMalloc(&d_tmp, tmp_bytes));
        cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, d_ky, d_k1, d_i0, d_i1, n);          // by y: order in d_i1
        large_gather_key<<<G, B>>>(d_k0, d_i1, n, d_ky);                                       // x keys in y order
        cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, d_ky, d_k1, d_i1, d_i0, n);          // by x (stable): perm in d_i0
        large_gather_pts<<<G, B>>>(d_pts, d_i0, n, d_pt, d_head);
        large_fix_ties<<<G, B>>>(d_pt, d_i0, n, d_ctl + 2);
        large_check_dups<<<G, B>>>(d_pt, n, d_ctl + 2);
        LCK(cudaGetLastError());
        LCK(cudaMemcpy(hctl, d_ctl, 16, cudaMemcpyDeviceToHost));
    }
    if (hctl[2]) {
        int st = 0;
        if (hctl[2] & 0x100u) st |= GTA_B200_ST_NONFINITE;
        if (hctl[2] & 0x200u) st |= GTA_B200_ST_NEAR_TIE_X;
        if (hctl[2] & 0x400u) st |= GTA_B200_ST_DUPLICATE;
        if (status) *status = st;
        goto done;
    }
    {
        FrameL f;
        f.pt = d_pt; f.head = d_head; f.dst = d_dst; f.nxt = d_nxt; f.prv = d_prv;
        f.bump = d_ctl; f.stack = nullptr; f.err = d_ctl + 1; f.n = n; f.cap_e = cap_e;
        const int D = tree_depth(n);
        // levels with at most this many merges use the warp-cooperative merge / one warp per merge (initialised once, thread-safe)
        static const int coop_nodes = [] { const char* e = getenv("GTA_B200_LARGE_COOP_NODES"); return e ? atoi(e) : 16384; }();
        static const int warp_nodes = [] { const char* e = getenv("GTA_B200_LARGE_WARP_NODES"); return e ? atoi(e) : 16384; }();
        for (int d = D; d >= 0; --d) {
            const int nodes = 1 << d;
            const int lpn = nodes <= warp_nodes ? 32 : 1;          // one warp per merge while the GPU has warps to spare
            const long long threads = (long long)nodes * lpn;
            large_level<<<(unsigned)((threads + 127) / 128), 128>>>(f, d, nodes, lpn, nodes <= coop_nodes ? 1 : 0);
        }
        LCK(cudaGetLastError());
        large_count<<<G, B>>>(f, d_cnt);
        cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_cnt, d_off, n);
        uint32_t last_off = 0, last_cnt = 0;
        LCK(cudaMemcpy(&last_off, d_off + (n - 1), 4, cudaMemcpyDeviceToHost));
        LCK(cudaMemcpy(&last_cnt, d_cnt + (n - 1), 4, cudaMemcpyDeviceToHost));
        LCK(cudaMemcpy(hctl, d_ctl, 16, cudaMemcpyDeviceToHost));
        int st = 0;
        if (hctl[1] & GT_ERR_POOL) st |= GTA_B200_ST_POOL;
        if (hctl[1] & GT_ERR_RANGE) st |= GTA_B200_ST_PRECISION_RANGE;
        if (hctl[1] & GT_ERR_RING) st |= GTA_B200_ST_INTERNAL;
        if (status) *status = st;
        if (st) goto done;
        const int total = (int)(last_off + last_cnt);
        if (ntri) *ntri = total;
        if (xyz3) {
            double hs[2] = {0.0, 0.0};
            large_area<<<G, B>>>(f, d_i0, d_xyz, (flags & GTA_B200_2D) ? 1 : 0, d_a3, d_a2);
            cub::DeviceReduce::Sum(d_tmp, tmp_bytes, d_a3, d_sum, n);
            cub::DeviceReduce::Sum(d_tmp, tmp_bytes, d_a2, d_sum + 1, n);
            LCK(cudaGetLastError());
            LCK(cudaMemcpy(hs, d_sum, 16, cudaMemcpyDeviceToHost));
            if (area3) *area3 = hs[0];
            if (area2) *area2 = hs[1];
        }
        if (tri && tri_cap > 0) {
            const int keep = std::min(total, tri_cap);
            LCK(cudaMalloc((void**)&d_tri, sizeof(int) * 3 * (size_t)std::max(keep, 1)));
            large_emit<<<G, B>>>(f, d_off, d_i0, d_tri, keep);
            LCK(cudaGetLastError());
            LCK(cudaEventRecord(ev1));
            LCK(cudaMemcpy(tri, d_tri, sizeof(int) * 3 * (size_t)keep, cudaMemcpyDeviceToHost));
        } else LCK(cudaEventRecord(ev1));
        LCK(cudaEventSynchronize(ev1));
        float ms = 0.f; cudaEventElapsedTime(&ms, ev0, ev1); g_large_ms = ms;
    }
done:
    ls_free(d_pts); cudaFree(d_pt); cudaFree(d_k0); cudaFree(d_k1); cudaFree(d_ky); cudaFree(d_i0); cudaFree(d_i1);
    cudaFree(d_head); cudaFree(d_dst); cudaFree(d_nxt); cudaFree(d_prv); cudaFree(d_ctl); cudaFree(d_cnt); cudaFree(d_off);
    cudaFree(d_tri); cudaFree(d_tmp); ls_free(d_xyz); cudaFree(d_a3); cudaFree(d_a2); cudaFree(d_sum);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    return rc;
}

extern "C" int gta_b200_triangulate_large(const double* points, int npoints, int device, int* tri, int tri_cap, int* ntri, int* status) {
    return large_run(points, nullptr, npoints, device, tri, tri_cap, ntri, status, 0u, nullptr, nullptr);
}
// One large frame of atoms [n][3] without correction points: triangulation of (x, y) + summed 3-D (and, with GTA_B200_2D,
// projected) triangle areas. Serves delaunay_surface_area / delaunay_tessellate for frames beyond gta_b200_max_points().
int gta_b200_large_corr(double* d_xyz, int natoms, int ncorr, double bx, double by, double espace, int* st);   // gta_b200.cu
// One large frame the way delaunay_tessellate treats every frame (reference src/gta_tri.c:106-296): with GTA_B200_CORRECT
// the box-edge points are generated on the device (the batched kernels' corr_points, one CTA), appended, and the frame of
// natoms + ncorr points goes through the large-frame path; without it this is gta_b200_surface_area_large.
extern "C" int gta_b200_tessellate_large(const double* xyz, int natoms, const double* box9, double espace, unsigned flags, int device,
                                         double* area3, double* area2, double* areabox, int* ntri, int* ncorr_out, int* status) {
    if (!xyz || !box9 || natoms < 0) { g_lerr = "bad argument"; return GTA_B200_E_ARG; }
    const double bx = box9[0], by = box9[4];
    if (areabox) *areabox = bx * by;
    if (ncorr_out) *ncorr_out = 0;
    if (!(flags & GTA_B200_CORRECT)) return large_run(nullptr, xyz, natoms, device, nullptr, 0, ntri, status, flags, area3, area2);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) { cudaGetLastError(); g_lerr = "no CUDA device: libgtessla_b200 has no CPU path"; return GTA_B200_E_NO_DEVICE; }
    if (device >= 0 && cudaSetDevice(device) != cudaSuccess) { g_lerr = "cudaSetDevice failed"; return GTA_B200_E_CUDA; }
    if (status) *status = 0;
    if (ntri) *ntri = 0;
    const double qn = std::numeric_limits<double>::quiet_NaN();
    if (area3) *area3 = qn;
    if (area2) *area2 = qn;
    const int nx = (int)(bx / espace), ny = (int)(by / espace);               // gta_tri.c:112-113
    if (!std::isfinite(bx) || !std::isfinite(by) || nx < 0 || ny < 0 || natoms < 1) { if (status) *status = natoms < 1 ? GTA_B200_ST_TOO_FEW_POINTS : GTA_B200_ST_NONFINITE; return GTA_B200_OK; }
    if (nx > (1 << 20) || ny > (1 << 20) || (long long)natoms + 4 + 2ll * nx + 2ll * ny > (1 << 27)) { if (status) *status = GTA_B200_ST_CAPACITY; return GTA_B200_OK; }
    const int ncorr = 4 + 2 * nx + 2 * ny, n = natoms + ncorr;
    double* d_all = nullptr;
    cudaError_t e = ls_alloc((void**)&d_all, sizeof(double) * 3 * (size_t)n);
    if (e == cudaSuccess) e = cudaMemcpy(d_all, xyz, sizeof(double) * 3 * (size_t)natoms, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { ls_free(d_all); g_lerr = std::string("gta_b200_tessellate_large: ") + cudaGetErrorString(e); return GTA_B200_E_CUDA; }
    int st = 0;
    int rc = gta_b200_large_corr(d_all, natoms, ncorr, bx, by, espace, &st);
    if (rc != GTA_B200_OK) { g_lerr = gta_b200_last_error(); ls_free(d_all); return rc; }
    if (st) { if (status) *status = st; ls_free(d_all); return GTA_B200_OK; }          // (atoms outside the box, ...: the batched calls' statuses)
    if (ncorr_out) *ncorr_out = ncorr;
    rc = large_run(nullptr, d_all, n, -1, nullptr, 0, ntri, status, flags & ~(unsigned)GTA_B200_CORRECT, area3, area2);
    ls_free(d_all);
    return rc;
}
extern "C" int gta_b200_surface_area_large(const double* xyz, int natoms, unsigned flags, int device, double* area3, double* area2, int* ntri, int* status) {
    if (flags & GTA_B200_CORRECT) { g_lerr = "the correction points (-corr) are generated by the batched kernel only: frame too large"; return GTA_B200_E_TOO_LARGE; }
    return large_run(nullptr, xyz, natoms, device, nullptr, 0, ntri, status, flags, area3, area2);
}
