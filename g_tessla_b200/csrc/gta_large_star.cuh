// gta_large_star.cuh -- one large frame through the star engine (star_core.cuh) with the frame in GLOBAL memory.
//
// The same per-point stars as the batched kernel (gta_star.cuh), for a point set that does not fit shared memory
// (BASELINE cfg 4: 10^6 points): uniform grid of ~1 point per cell built with a stable radix sort on the cell index
// (order inside a cell = input order), a thread per point, the threads' candidate lists in a global scratch that stays
// in L2. Whole stars (no -corr points: the hull is not the box), so the triangle count rule 2(n-1)-h is checked with
// h = open stars. Hull vertices and the skinny triangles along the hull are what a grid handles badly -- their empty
// regions reach far -- and are served by the scan-converted region of star_core.cuh (StarRegion): a half-plane or a
// circumdisc costs its area, not its reach.
// Replaces dtriangulate (reference src/delaunay_tri.c:413-507) + the area loop (src/gta_tri.c:368-410) for such a frame
// when its Delaunay triangulation is unique; the triangles come out in cell order, NOT in the reference's emission
// order, so the kept dtriangulate keeps using the level-by-level path (gta_large.cu) unless GTA_B200_PATH=2. A frame
// with a tie, duplicates, or a failed count rule is reported as "not certified" and the caller falls back to that path.
#pragma once
#include "star_core.cuh"
#include "dt_core.cuh"

namespace gt {

constexpr int LS_LCAP = 32, LS_CCAP = 56, LS_MAXOWN = 10;

__device__ __forceinline__ unsigned long long ls_okey(double v) {
    unsigned long long b = (unsigned long long)__double_as_longlong(v + 0.0);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__host__ __device__ inline double ls_unkey(unsigned long long k) {
    const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)b);
#else
    double v; memcpy(&v, &b, 8); return v;
#endif
}
// bounding box as order-preserving keys: key4 = {min x, max x, min y, max y}; err |= 0x100 on a non-finite coordinate
__global__ void ls_bbox(const double* pts, int n, unsigned long long* key4, uint32_t* err) {
    unsigned long long k[4] = {~0ull, 0ull, ~0ull, 0ull};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double x = pts[2 * (size_t)i], y = pts[2 * (size_t)i + 1];
        if (!isfinite(x) || !isfinite(y)) { atomicOr(err, 0x100u); continue; }
        const unsigned long long kx = ls_okey(x), ky = ls_okey(y);
        k[0] = kx < k[0] ? kx : k[0]; k[1] = kx > k[1] ? kx : k[1]; k[2] = ky < k[2] ? ky : k[2]; k[3] = ky > k[3] ? ky : k[3];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const unsigned long long t = __shfl_xor_sync(0xffffffffu, k[c], o);
            k[c] = (c & 1) ? (t > k[c] ? t : k[c]) : (t < k[c] ? t : k[c]);
        }
    }
    if ((threadIdx.x & 31) == 0) { atomicMin(&key4[0], k[0]); atomicMax(&key4[1], k[1]); atomicMin(&key4[2], k[2]); atomicMax(&key4[3], k[3]); }
}
__global__ void ls_cells(const double* pts, int n, double x0, double y0, double sx, double sy, int gx, int gy, uint32_t* cell, uint32_t* idx, uint32_t* hist) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = star_cell(pts[2 * (size_t)i + 1], y0, sy, gy) * gx + star_cell(pts[2 * (size_t)i], x0, sx, gx);
    cell[i] = (uint32_t)c; idx[i] = (uint32_t)i;
    atomicAdd(&hist[c], 1u);
}
__global__ void ls_gather(const double* pts, const double* xyz, const uint32_t* idx, int n, SPt* spt, SPtF* sptf, double* sz) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t o = idx[i];
    SPt v; v.x = pts[2 * (size_t)o]; v.y = pts[2 * (size_t)o + 1];
    spt[i] = v;
    SPtF f; f.x = (float)v.x; f.y = (float)v.y; sptf[i] = f;
    if (sz) sz[i] = xyz ? xyz[3 * (size_t)o + 2] : 0.0;
}

// How many points sit in cells of more than `limit` points? The grid has one point per cell on average; where a cell holds
// dozens, a point's 5 x 5 window overflows the lane's candidate list and its scans cost the square of the density. Strongly
// clustered input is the level-by-level path's (whose cost does not depend on the density): the caller decides on *dense.
__global__ void ls_dense(const uint32_t* count, int ncell, uint32_t limit, uint32_t* dense) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t v = c < ncell ? count[c] : 0u;
    v = v > limit ? v : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(dense, v);
}
// smallest and largest x of every grid row's points (star_core.cuh: StarGrid::rowx); a row is one run of the cell-sorted array
__global__ void ls_rowx(const SPt* spt, const uint32_t* cstart, int gx, int gy, double* rowx) {
    const int r = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (r >= gy) return;
    const uint32_t a = cstart[(size_t)r * gx], b = cstart[(size_t)(r + 1) * gx];
    double lo = __longlong_as_double(0x7ff0000000000000ll), hi = -lo;
    for (uint32_t i = a + lane; i < b; i += 32) { const double x = spt[i].x; lo = fmin(lo, x); hi = fmax(hi, x); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o)); hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o)); }
    if (lane == 0) { rowx[2 * (size_t)r] = lo; rowx[2 * (size_t)r + 1] = hi; }
}

struct LsList {
    uint32_t* l; uint32_t* c; size_t nt;
    __device__ void put(int k, int pos) { l[(size_t)k * nt] = (uint32_t)pos; }
    __device__ int get(int k) const { return (int)l[(size_t)k * nt]; }
    __device__ void cput(int k, int pos) { c[(size_t)k * nt] = (uint32_t)pos; }
    __device__ int cget(int k) const { return (int)c[(size_t)k * nt]; }
};

// ctl: [0] STAR_* bits of any point | 0x1000 when the overflow list is full, [1] open stars (hull vertices),
// [2] entries of the overflow list: triangles beyond the LS_MAXOWN a point keeps in its own slots (hull vertices own long fans)
// work / nwork (nullable): the positions to do (the points the tile kernel below deferred) instead of all of 0 .. n-1,
// `per` of them per warp. Deferred points are the ones whose searches reach far (hull edges, big empty discs), and what
// makes such a search long -- the row-by-row advance of its region -- runs inside a lane's own branch: 32 of them in one
// warp take the SUM of their scans, not the longest (1,499 deferred points of 10^6 took 42 ms at 32 per warp).
__global__ void __launch_bounds__(128) ls_star_kernel(StarGrid<uint32_t, true> g, uint32_t* llist, uint32_t* clist, const double* sz, int want2d,
                                                      uint32_t* own_cnt, uint32_t* own_pairs, uint32_t* ovf, uint32_t ovf_cap, double* a3, double* a2, uint32_t* ctl,
                                                      unsigned long long* chunk_clk, const uint32_t* work, const uint32_t* nwork, int per) {
    const size_t nt = (size_t)gridDim.x * blockDim.x, gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const int nwarps = (int)(nt >> 5), gwarp = (int)(gtid >> 5);
    LsList lst; lst.l = llist + gtid; lst.c = clist + gtid; lst.nt = nt;
    const int n = work ? (int)*nwork : g.n;
    // chunks of 32 consecutive positions; a warp's first chunk is its own number, the rest come from a counter, so that the
    // chunks along the hull -- whose searches reach far -- do not pile up on the warps they would statically fall to
    for (int ch = gwarp; (long long)ch * per < n;) {
        const bool active = lane < per && ch * per + lane < n;
        const int ip = work ? (active ? (int)work[ch * per + lane] : 0) : ch * per + lane;
        const long long clk0 = chunk_clk ? clock64() : 0;
        const SPt P = g.pt[active ? ip : 0];
        const double zp = sz ? sz[active ? ip : 0] : 0.0;
        double s3 = 0.0, s2 = 0.0; int cnt = 0;
        bool hull = false;
        const int st = star_point<LS_LCAP, LS_CCAP, true>(g, ip, active, 2, lst, &hull, [&](int n1, int n2) {
            if (a3) {
                const SPt A = g.pt[n1], B = g.pt[n2];
                const double za = sz ? sz[n1] : 0.0, zb = sz ? sz[n2] : 0.0;
                s3 += tri_area3(P.x, P.y, zp, A.x, A.y, za, B.x, B.y, zb);
                if (want2d) s2 += tri_area3(P.x, P.y, 0.0, A.x, A.y, 0.0, B.x, B.y, 0.0);
            }
            if (own_pairs) {
                if (cnt < LS_MAXOWN) { own_pairs[((size_t)ip * LS_MAXOWN + cnt) * 2] = (uint32_t)n1; own_pairs[((size_t)ip * LS_MAXOWN + cnt) * 2 + 1] = (uint32_t)n2; }
                else {
                    const uint32_t o = atomicAdd(&ctl[2], 1u);
                    if (o < ovf_cap) { ovf[3 * (size_t)o] = (uint32_t)ip; ovf[3 * (size_t)o + 1] = (uint32_t)n1; ovf[3 * (size_t)o + 2] = (uint32_t)n2; }
                    else atomicOr(&ctl[0], 0x1000u);
                }
            }
            ++cnt;
        });
        if (active) {
            own_cnt[ip] = (uint32_t)(own_pairs && cnt > LS_MAXOWN ? LS_MAXOWN : cnt);      // (slots; the rest is in the overflow list)
            if (a3) { a3[ip] = s3; a2[ip] = s2; }
            if (st) atomicOr(&ctl[0], (uint32_t)st);
            if (hull) atomicAdd(&ctl[1], 1u);
        }
        if (chunk_clk && lane == 0) chunk_clk[ch] = (unsigned long long)(clock64() - clk0);
        int nextch = 0;
        if (lane == 0) nextch = nwarps + (int)atomicAdd(&ctl[3], 1u);
        ch = __shfl_sync(0xffffffffu, nextch, 0);
    }
}
// ---------------------------------------------------------------------------------------------------------------------
// The same stars from TILES of the grid staged in shared memory. In ls_star_kernel every candidate costs two dependent
// global loads (list entry -> point) and the lanes' lists live in L2: 25 ns per point against 1.3 ns in the batched kernel,
// whose frame is on chip. Here a CTA takes a tile of LT_TW x LT_TH cells (~1,000 points), copies the tile and a halo of
// LT_H cells around it into shared memory -- every row of the staged rectangle is ONE contiguous run of the cell-sorted
// arrays -- and runs the batched kernel's code on it (StarGrid GL = 2: the tile bins exactly like the frame's grid, and
// a search that would have to look beyond a side of the tile that is not a side of the grid gives up: STAR_DEFER).
// Deferred points (long hull edges, big empty discs, overfull tiles, fans of more than LS_MAXOWN owned triangles) are
// appended to a list that ls_star_kernel works off afterwards; the per-point outputs are the same arrays.
constexpr int LT_TW = 32, LT_TH = 32, LT_H = 6, LT_W = LT_TW + 2 * LT_H, LT_T = 256, LT_CAP = 2304, LT_LCAP = 24, LT_CCAP = 48;
struct LtSmem {
    size_t pt, ptf, cstart, rowg0, rowbase, corebase, core, clist, lists, ctl, total;
    __host__ __device__ LtSmem() {
        size_t o = 0;
        pt = o; o += sizeof(SPt) * (size_t)LT_CAP;
        ptf = o; o += sizeof(SPtF) * (size_t)LT_CAP;
        cstart = o; o += 4 * (size_t)(LT_W * LT_W + 4);
        rowg0 = o; o += 4 * (size_t)(LT_W + 4);
        rowbase = o; o += 4 * (size_t)(LT_W + 4);
        corebase = o; o += 4 * (size_t)(LT_W + 4);
        core = o; o += 2 * (size_t)LT_CAP;
        clist = o; o += 2 * (size_t)LT_CCAP * LT_T;
        lists = o; o += 2 * (size_t)LT_LCAP * LT_T;
        ctl = o; o += 64;
        total = (o + 15) & ~(size_t)15;
    }
};
// ctl: as ls_star_kernel's, and [4] = entries of the deferred list
__global__ void __launch_bounds__(LT_T, 2) ls_tile_kernel(StarGrid<uint32_t, true> gg, int tiles_x, int ntiles, const double* sz, int want2d,
                                                          uint32_t* own_cnt, uint32_t* own_pairs, double* a3, double* a2, uint32_t* ctl, uint32_t* deferred) {
    extern __shared__ __align__(16) unsigned char lsm[];
    const LtSmem M;
    SPt* pt = reinterpret_cast<SPt*>(lsm + M.pt);
    SPtF* ptf = reinterpret_cast<SPtF*>(lsm + M.ptf);
    uint32_t* cs = reinterpret_cast<uint32_t*>(lsm + M.cstart);
    uint32_t* rowg0 = reinterpret_cast<uint32_t*>(lsm + M.rowg0);
    uint32_t* rowbase = reinterpret_cast<uint32_t*>(lsm + M.rowbase);
    uint32_t* corebase = reinterpret_cast<uint32_t*>(lsm + M.corebase);
    unsigned short* core = reinterpret_cast<unsigned short*>(lsm + M.core);
    unsigned short* clist = reinterpret_cast<unsigned short*>(lsm + M.clist);
    unsigned short* lists = reinterpret_cast<unsigned short*>(lsm + M.lists);
    uint32_t* sctl = reinterpret_cast<uint32_t*>(lsm + M.ctl);          // [0] staged points, [1] next chunk, [2] core points
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = LT_T / 32;
    const uint32_t* gcs = gg.cstart;
    const int gx = gg.gx, gy = gg.gy;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int tx = t % tiles_x, ty = t / tiles_x;
        const int cc0 = tx * LT_TW, cc1 = min(gx, cc0 + LT_TW) - 1, cr0 = ty * LT_TH, cr1 = min(gy, cr0 + LT_TH) - 1;       // the tile's own cells
        const int C0 = max(0, cc0 - LT_H), C1 = min(gx - 1, cc1 + LT_H), R0 = max(0, cr0 - LT_H), R1 = min(gy - 1, cr1 + LT_H);   // staged cells
        const int W = C1 - C0 + 1, Hh = R1 - R0 + 1, ncr = cr1 - cr0 + 1;
        __syncthreads();                                                   // (the previous tile is done with the arrays)
        if (tid < Hh) {
            const size_t base = (size_t)(R0 + tid) * gx;
            const uint32_t a = gcs[base + C0];
            rowg0[tid] = a; rowbase[tid + 1] = gcs[base + C1 + 1] - a;
        }
        if (tid >= 64 && tid < 64 + ncr) {
            const size_t base = (size_t)(cr0 + tid - 64) * gx;
            corebase[tid - 64 + 1] = gcs[base + cc1 + 1] - gcs[base + cc0];
        }
        __syncthreads();
        if (tid == 0) {
            uint32_t run = 0; rowbase[0] = 0;
            for (int r = 0; r < Hh; ++r) { run += rowbase[r + 1]; rowbase[r + 1] = run; }
            sctl[0] = run; sctl[1] = NW;
        }
        if (tid == 32) {
            uint32_t run = 0; corebase[0] = 0;
            for (int r = 0; r < ncr; ++r) { run += corebase[r + 1]; corebase[r + 1] = run; }
            sctl[2] = run;
        }
        __syncthreads();
        const int nloc = (int)sctl[0], ncore = (int)sctl[2];
        if (nloc > LT_CAP || nloc < 3) {
            // an overfull (clustered input) or nearly empty tile: all its points go to the kernel that reads global memory
            for (int r = warp; r < ncr; r += NW) {
                const size_t base = (size_t)(cr0 + r) * gx;
                const uint32_t a = gcs[base + cc0], b = gcs[base + cc1 + 1];
                for (uint32_t q = a + lane; q < b; q += 32) deferred[atomicAdd(&ctl[4], 1u)] = q;
            }
            continue;
        }
        for (int k = tid; k < W * Hh; k += LT_T) {
            const int r = k / W, c = k - r * W;
            cs[k] = gcs[(size_t)(R0 + r) * gx + C0 + c] - rowg0[r] + rowbase[r];
        }
        if (tid == 0) cs[W * Hh] = (uint32_t)nloc;
        for (int r = warp; r < Hh; r += NW) {
            const uint32_t a = rowg0[r], lb = rowbase[r], cnt = rowbase[r + 1] - lb;
            for (uint32_t j = lane; j < cnt; j += 32) { pt[lb + j] = gg.pt[a + j]; ptf[lb + j] = gg.ptf[a + j]; }
        }
        __syncthreads();
        for (int r = warp; r < ncr; r += NW) {
            const int lr = cr0 - R0 + r;
            const uint32_t a = cs[lr * W + (cc0 - C0)], cnt = corebase[r + 1] - corebase[r], cb = corebase[r];
            for (uint32_t j = lane; j < cnt; j += 32) core[cb + j] = (unsigned short)(a + j);
        }
        __syncthreads();
        StarGrid<uint32_t, 2> g;
        g.pt = pt; g.ptf = ptf; g.cstart = cs;
        g.pt_s = (uint32_t)__cvta_generic_to_shared(pt); g.ptf_s = (uint32_t)__cvta_generic_to_shared(ptf); g.cstart_s = (uint32_t)__cvta_generic_to_shared(cs);
        g.kf = gg.kf; g.n = nloc; g.gx = W; g.gy = Hh; g.x0 = gg.x0; g.y0 = gg.y0; g.sx = gg.sx; g.sy = gg.sy;
        g.xmin = gg.xmin; g.xmax = gg.xmax; g.ymin = gg.ymin; g.ymax = gg.ymax;
        g.cx0 = C0; g.cy0 = R0; g.ggx = gx; g.ggy = gy;
        g.open = (C0 > 0 ? 1 : 0) | (C1 < gx - 1 ? 2 : 0) | (R0 > 0 ? 4 : 0) | (R1 < gy - 1 ? 8 : 0);
        // staged position -> position in the frame's cell-sorted arrays (the point's row comes from its y, as it was binned)
        auto gpos = [&](int lp) -> uint32_t { const int r = star_row(g, pt[lp].y); return rowg0[r] + (uint32_t)lp - rowbase[r]; };
        struct { uint32_t base, cbase;
                 __device__ void put(int k, int pos) { asm volatile("st.shared.u16 [%0], %1;" :: "r"(base + 2u * LT_T * (uint32_t)k), "h"((unsigned short)pos) : "memory"); }
                 __device__ int get(int k) const { unsigned short v; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(base + 2u * LT_T * (uint32_t)k) : "memory"); return v; }
                 __device__ void cput(int k, int pos) { asm volatile("st.shared.u16 [%0], %1;" :: "r"(cbase + 2u * LT_T * (uint32_t)k), "h"((unsigned short)pos) : "memory"); }
                 __device__ int cget(int k) const { unsigned short v; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(cbase + 2u * LT_T * (uint32_t)k) : "memory"); return v; } } lst;
        lst.base = (uint32_t)__cvta_generic_to_shared(lists + tid); lst.cbase = (uint32_t)__cvta_generic_to_shared(clist + tid);
        for (int ch = warp; ch * 32 < ncore;) {
            const bool active = ch * 32 + lane < ncore;
            const int lp = active ? (int)core[ch * 32 + lane] : 0;
            const uint32_t gip = gpos(lp);
            const SPt P = pt[lp];
            const double zp = sz ? sz[gip] : 0.0;
            double s3 = 0.0, s2 = 0.0; int cnt = 0;
            bool hull = false, fan = false;
            const int st = star_point<LT_LCAP, LT_CCAP, true>(g, lp, active, 2, lst, &hull, [&](int n1, int n2) {
                const uint32_t g1 = gpos(n1), g2 = gpos(n2);
                if (a3) {
                    const SPt A = pt[n1], B = pt[n2];
                    const double za = sz ? sz[g1] : 0.0, zb = sz ? sz[g2] : 0.0;
                    s3 += tri_area3(P.x, P.y, zp, A.x, A.y, za, B.x, B.y, zb);
                    if (want2d) s2 += tri_area3(P.x, P.y, 0.0, A.x, A.y, 0.0, B.x, B.y, 0.0);
                }
                if (own_pairs) {
                    if (cnt < LS_MAXOWN) { own_pairs[((size_t)gip * LS_MAXOWN + cnt) * 2] = g1; own_pairs[((size_t)gip * LS_MAXOWN + cnt) * 2 + 1] = g2; }
                    else fan = true;                                       // (a long fan: the other kernel has the overflow list)
                }
                ++cnt;
            });
            if (active) {
                if ((st & STAR_DEFER) || fan) deferred[atomicAdd(&ctl[4], 1u)] = gip;
                else {
                    own_cnt[gip] = (uint32_t)cnt;
                    if (a3) { a3[gip] = s3; a2[gip] = s2; }
                    if (st) atomicOr(&ctl[0], (uint32_t)st);
                    if (hull) atomicAdd(&ctl[1], 1u);
                }
            }
            int nextch = 0;
            if (lane == 0) nextch = (int)atomicAdd(&sctl[1], 1u);
            ch = __shfl_sync(0xffffffffu, nextch, 0);
        }
    }
}

__global__ void ls_emit(const uint32_t* own_cnt, const uint32_t* offs, const uint32_t* own_pairs, const uint32_t* idx, int n, int* tri, int tri_cap) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t c = own_cnt[i], o = offs[i];
    for (uint32_t k = 0; k < c && k < (uint32_t)LS_MAXOWN; ++k) {
        if ((int)(o + k) >= tri_cap) break;
        tri[3 * (size_t)(o + k)] = (int)idx[i];
        tri[3 * (size_t)(o + k) + 1] = (int)idx[own_pairs[((size_t)i * LS_MAXOWN + k) * 2]];
        tri[3 * (size_t)(o + k) + 2] = (int)idx[own_pairs[((size_t)i * LS_MAXOWN + k) * 2 + 1]];
    }
}

__global__ void ls_emit_overflow(const uint32_t* ovf, int novf, const uint32_t* idx, int* tri, int base, int tri_cap) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= novf || base + k >= tri_cap) return;
    tri[3 * (size_t)(base + k)] = (int)idx[ovf[3 * (size_t)k]];
    tri[3 * (size_t)(base + k) + 1] = (int)idx[ovf[3 * (size_t)k + 1]];
    tri[3 * (size_t)(base + k) + 2] = (int)idx[ovf[3 * (size_t)k + 2]];
}

}  // namespace gt
