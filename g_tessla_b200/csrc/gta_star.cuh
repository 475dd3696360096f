// gta_star.cuh -- the fast path of the batched calls: ONE fused launch, one CTA per frame, one thread per point.
//
// Replaces, per frame, the whole body of the reference's OpenMP loop (src/gta_tri.c:106-296) when the frame's
// Delaunay triangulation is unique (star_core.cuh has the argument and the certificate):
//   load        the frame's coordinate tile by one bulk asynchronous copy (cp.async.bulk + mbarrier, the TMA unit);
//               the next frame's tile is fetched while this frame's stars are built
//   K0  -corr   gta_corr.cuh (gta_tri.c:112-272), appended points bit-identical to the reference's
//   bin         uniform grid over the points' bounding box, ~1 point per cell: shared-memory histogram, scan, scatter;
//               points are stored in cell order (input order inside a cell), so a window of cells is a few contiguous
//               runs and neighbouring threads read neighbouring points
//   K1  stars   star_point() for every point: nearest neighbour, then the ring of Delaunay neighbours by exact
//               orient2d / incircle over the window's candidates; no rings in memory, no dependent loads beyond
//               the candidate's 16 bytes of shared memory
//   K2  area    every owned triangle's area_tri (include/gta_tri.h:91-99) in the walk itself, fixed-order block sum
// Nothing but the coordinates is read from HBM and 8-28 bytes per frame are written.
// A frame the stars do not certify (cocircular ties, duplicates, a status the divide-and-conquer path has to report,
// triangle count != 2(n-1)-h) is appended to the batch's fallback list and goes through the divide-and-conquer
// pipeline (pre-pass -> walkers -> areas), which keeps the reference's tie-breaking.
#pragma once
#include "gta_kernel.cuh"
#include "gta_corr.cuh"
#include "gta_prepass.cuh"
#include "star_core.cuh"

namespace gt {

#ifndef GT_STAR_W0
#define GT_STAR_W0 2          // half-width of a point's first window, in cells
#endif
#ifndef GT_STAR_LCAP
#define GT_STAR_LCAP 24
#endif
#ifndef GT_STAR_CCAP
#define GT_STAR_CCAP 40
#endif
#ifndef GT_STAR_MINB
#define GT_STAR_MINB 2        // CTAs of 256 threads per SM the kernel is compiled for (3: 80 registers per thread)
#endif
#ifndef GT_STAR_DENS
#define GT_STAR_DENS 1.0      // points per cell
#endif

constexpr int STAR_LCAP = GT_STAR_LCAP;                  // left candidates a lane keeps per round of a walk step (star_core.cuh)
constexpr int STAR_CCAP = GT_STAR_CCAP;                  // candidates a lane keeps (the points of its window); more: it walks the grid

struct StarParams {
    KParams k;
    int* fb_list; unsigned int* fb_count;      // frames that need the divide-and-conquer path
    unsigned long long* counters;              // [0] frames certified, [1] fallback frames, [2] exact-stage calls (sampled)
};

struct StarSmem {
    // persistent over a frame: cell-sorted points (double, float, z, input index), cell starts, chunk sums, control words.
    // setup (tile, -corr points, scatter scratch) and star phase (the threads' candidate / left lists) share one region.
    size_t spt, sptf, sz, cstart, sid, csum, ctl, tile, corr, tmp, cell, clist, lists, total;
    int ncell_max;
    __host__ __device__ StarSmem(int cap, int natoms, int in_dim, int threads) {
        size_t o = 0;
        spt = o; o += sizeof(SPt) * (size_t)cap;                                // cell-sorted (x, y); -corr scratch before that
        sz = o; o += 8 * (size_t)cap;                                           // cell-sorted z
        ncell_max = cap;
        cstart = o; o += 4 * (size_t)(ncell_max + 2);
        sid = o; o += 2 * (size_t)cap;                                          // cell-sorted input index
        o = (o + 15) & ~(size_t)15;
        csum = o; o += 16 * (size_t)((cap + 31) / 32 + 1);                      // area sums of every chunk of 32 points
        ctl = o; o += 16 * 4 + 8 * 32 * 8 + 8;                                  // status / counters, reduction scratch, one mbarrier
        o = (o + 127) & ~(size_t)127;
        const size_t shared0 = o;
        tile = o; o += (((size_t)natoms * in_dim * 8) + 15) & ~(size_t)15;       // bulk-copy destination
        corr = o; o += 24 * (size_t)(cap > natoms ? cap - natoms : 0); o = (o + 15) & ~(size_t)15;
        tmp = o; o += 2 * (size_t)cap;                                          // scatter order
        cell = o; o += 2 * (size_t)cap;                                         // cell of each input point
        const size_t setup_end = o;
        o = shared0;
        clist = o; o += 2 * (size_t)STAR_CCAP * threads;                        // every thread's candidates (its window's points)
        lists = o; o += 2 * (size_t)STAR_LCAP * threads;                        // every thread's left candidates of the current step
        o = (o + 15) & ~(size_t)15;
        sptf = o; o += sizeof(SPtF) * (size_t)cap;                              // cell-sorted (x, y) in float (screening), written when the star phase starts
        o = o > setup_end ? o : setup_end;
        total = (o + 15) & ~(size_t)15;
    }
};

// FULL = false (frames with -corr points: the hull is the box, every star is closed or ends on it): a point walks only the
// arc of its star that holds the triangles it owns (star_core.cuh); FULL = true: whole stars.
template <int T, int MINB, bool FULL>
__global__ void __launch_bounds__(T, MINB) star_kernel(const StarParams sp) {
    extern __shared__ __align__(128) unsigned char ssm[];
    const KParams& p = sp.k;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = T / 32;
    const int cap = p.cap, natoms = p.natoms, in_dim = p.in_dim;
    const StarSmem L(cap, natoms, in_dim, T);
    double* tile = reinterpret_cast<double*>(ssm + L.tile);
    double* corr3 = reinterpret_cast<double*>(ssm + L.corr);
    SPt* spt = reinterpret_cast<SPt*>(ssm + L.spt);
    SPtF* sptf = reinterpret_cast<SPtF*>(ssm + L.sptf);
    u16* clist = reinterpret_cast<u16*>(ssm + L.clist);                     // [STAR_CCAP][T]
    double* sz = reinterpret_cast<double*>(ssm + L.sz);
    uint32_t* cbase = reinterpret_cast<uint32_t*>(ssm + L.cstart);         // cstart[0 .. ncell]
    u16* sid = reinterpret_cast<u16*>(ssm + L.sid);
    u16* tmp = reinterpret_cast<u16*>(ssm + L.tmp);
    u16* cell = reinterpret_cast<u16*>(ssm + L.cell);
    double* csum = reinterpret_cast<double*>(ssm + L.csum);
    u16* lists = reinterpret_cast<u16*>(ssm + L.lists);                     // [STAR_LCAP][T]
    uint32_t* ctl = reinterpret_cast<uint32_t*>(ssm + L.ctl);              // [0] status, [1] triangles written, [3] next chunk of points
    double* red = reinterpret_cast<double*>(ctl + 16);                      // [8][32]
    uint64_t* mbar = reinterpret_cast<uint64_t*>(red + 8 * 32);
    unsigned long long* ckey = reinterpret_cast<unsigned long long*>(ssm + L.spt);

    const bool corr = (p.flags & GTA_B200_CORRECT) != 0;
    const bool want2d = (p.flags & GTA_B200_2D) != 0;
    const uint32_t frame_bytes = (uint32_t)((size_t)natoms * in_dim * 8);
    const bool bulk = (frame_bytes % 16 == 0) && ((reinterpret_cast<uintptr_t>(p.xyz) & 15) == 0) && frame_bytes > 0;

    if (tid == 0) pp_mbar_init(mbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    const int first = blockIdx.x, stride = gridDim.x;
    if (bulk && tid == 0 && first < p.nframes) pp_bulk_load(tile, p.xyz + (size_t)first * natoms * in_dim, frame_bytes, mbar);

    int it = 0;
    for (int fr = first; fr < p.nframes; fr += stride, ++it) {
        if (tid == 0) { ctl[0] = 0; ctl[1] = 0; ctl[3] = NW; ctl[4] = 0; ctl[5] = 0; }
        if (bulk) pp_mbar_wait(mbar, (uint32_t)(it & 1));
        else {
            const double* src = p.xyz + (size_t)fr * natoms * in_dim;
            for (int k = tid; k < natoms * in_dim; k += T) tile[k] = src[k];
        }
        __syncthreads();
        bool bad = false;
        for (int k = tid; k < natoms * in_dim; k += T) bad |= !isfinite(tile[k]);
        if (bad) atomicOr(&ctl[0], GTA_B200_ST_NONFINITE);
        double bx = 0.0, by = 0.0;
        if (p.box) { bx = p.box[(size_t)fr * p.box_stride]; by = p.box[(size_t)fr * p.box_stride + p.by_off]; }
        int n = natoms, ncorr_out = 0;
        __syncthreads();
        if (corr) n = corr_points<T>(tile, in_dim, natoms, cap, bx, by, p.espace, ckey, red, ctl, corr3, p.corr, p.corr_cap, fr, &ncorr_out);
        else if (natoms > cap && tid == 0) atomicOr(&ctl[0], GTA_B200_ST_CAPACITY);
        if (n < 2 && tid == 0) atomicOr(&ctl[0], GTA_B200_ST_TOO_FEW_POINTS);
        __syncthreads();
        auto px = [&](int i) -> double { return i < natoms ? tile[(size_t)i * in_dim] : corr3[3 * (i - natoms)]; };
        auto py = [&](int i) -> double { return i < natoms ? tile[(size_t)i * in_dim + 1] : corr3[3 * (i - natoms) + 1]; };
        auto pz = [&](int i) -> double { return i < natoms ? (in_dim == 3 ? tile[(size_t)i * 3 + 2] : 0.0) : corr3[3 * (i - natoms) + 2]; };
        const bool ok = ctl[0] == 0;                                       // (block-uniform: read after the barrier)

        int ntri = 0, nhull = 0, st = 0;
        if (ok) {
            // ---------------------------------------------------------------- bounding box of the points
            double xlo = 1.7976931348623157e308, xhi = -xlo, ylo = xlo, yhi = -xlo;
            for (int i = tid; i < n; i += T) { const double x = px(i), y = py(i); xlo = fmin(xlo, x); xhi = fmax(xhi, x); ylo = fmin(ylo, y); yhi = fmax(yhi, y); }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                xlo = fmin(xlo, __shfl_xor_sync(0xffffffffu, xlo, o)); xhi = fmax(xhi, __shfl_xor_sync(0xffffffffu, xhi, o));
                ylo = fmin(ylo, __shfl_xor_sync(0xffffffffu, ylo, o)); yhi = fmax(yhi, __shfl_xor_sync(0xffffffffu, yhi, o));
            }
            if (lane == 0) { red[4 * warp] = xlo; red[4 * warp + 1] = xhi; red[4 * warp + 2] = ylo; red[4 * warp + 3] = yhi; }
            __syncthreads();
#pragma unroll
            for (int w = 0; w < NW; ++w) { xlo = fmin(xlo, red[4 * w]); xhi = fmax(xhi, red[4 * w + 1]); ylo = fmin(ylo, red[4 * w + 2]); yhi = fmax(yhi, red[4 * w + 3]); }
            const StarDims dm = star_dims(n, xlo, xhi, ylo, yhi, L.ncell_max, GT_STAR_DENS);      // (every thread: same operands, same result)
            const int ncell = dm.gx * dm.gy;
            uint32_t* cs = cbase + 1;                                       // cs[c]: count -> start -> end of cell c (= start of c + 1)
            for (int c = tid; c <= ncell; c += T) cbase[c] = 0;
            __syncthreads();
            // ---------------------------------------------------------------- histogram, scan, scatter
            for (int i = tid; i < n; i += T) {
                const int c = star_cell(py(i), ylo, dm.sy, dm.gy) * dm.gx + star_cell(px(i), xlo, dm.sx, dm.gx);
                cell[i] = (u16)c;
                atomicAdd(&cs[c], 1u);
            }
            __syncthreads();
            {
                const int per = (ncell + T - 1) / T, c0 = tid * per, c1 = min(ncell, c0 + per);
                uint32_t sum = 0;
                for (int c = c0; c < c1; ++c) sum += cs[c];
                uint32_t tot = 0;
                uint32_t run = block_exclusive_scan<T>(sum, reinterpret_cast<unsigned*>(red), &tot);
                for (int c = c0; c < c1; ++c) { const uint32_t v = cs[c]; cs[c] = run; run += v; }
            }
            __syncthreads();
            for (int i = tid; i < n; i += T) tmp[atomicAdd(&cs[cell[i]], 1u)] = (u16)i;
            __syncthreads();
            // input order inside a cell (the scatter's order depends on the atomics): rank among the cell's members
            for (int a = tid; a < n; a += T) {
                const int i = tmp[a], c = cell[i];
                const int s = (int)cbase[c], e = (int)cbase[c + 1];
                int rank = 0;
                for (int q = s; q < e; ++q) rank += tmp[q] < i;
                SPt v; v.x = px(i); v.y = py(i);
                spt[s + rank] = v; sz[s + rank] = pz(i); sid[s + rank] = (u16)i;
                if (!FULL) {
                    // points on the bounding box: when its four corners are points, they are the hull vertices (triangle count rule)
                    const bool ex = v.x == xlo || v.x == xhi, ey = v.y == ylo || v.y == yhi;
                    if (ex || ey) atomicAdd(&ctl[5], 1u);
                    if (ex && ey) atomicAdd(&ctl[4], 1u);
                }
            }
            __syncthreads();
            for (int i = tid; i < n; i += T) { const SPt v = spt[i]; SPtF vf; vf.x = (float)v.x; vf.y = (float)v.y; sptf[i] = vf; }     // (over the dead tile)
            __syncthreads();
            // the next frame's coordinates are pulled into L2 while this frame's stars are built (the tile's shared memory holds
            // the threads' lists during the star phase, so the copy itself starts when the frame is done)
            if (bulk && tid == 0 && fr + stride < p.nframes)
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(p.xyz + (size_t)(fr + stride) * natoms * in_dim), "r"(frame_bytes) : "memory");
            // ---------------------------------------------------------------- stars + areas
            StarGrid<uint32_t> g;
            g.pt = spt; g.ptf = sptf; g.cstart = cbase;
            g.pt_s = pp_smem_addr(spt); g.ptf_s = pp_smem_addr(sptf); g.cstart_s = pp_smem_addr(cbase);
            g.kf = (float)(2.5 * 5.9604644775390625e-08 * fmax(fmax(fabs(xlo), fabs(xhi)), fmax(fabs(ylo), fabs(yhi)))) * 1.0001f; g.n = n; g.gx = dm.gx; g.gy = dm.gy;
            g.x0 = xlo; g.y0 = ylo; g.sx = dm.sx; g.sy = dm.sy; g.xmin = xlo; g.xmax = xhi; g.ymin = ylo; g.ymax = yhi;
            int* tri_out = p.tri ? p.tri + (size_t)fr * p.tri_cap * 3 : nullptr;
            // A warp's 32 lanes build the stars of 32 consecutive positions in lockstep (all of them call star_point, lanes
            // beyond n without a point). Chunks of 32 are handed out by a shared counter (a warp's first one is static), so
            // the warps finish together whatever their chunks cost; every chunk's area sums are kept apart and added in
            // chunk order below, so the frame's sum does not depend on which warp took which chunk.
            for (int ch = warp; ch * 32 < n;) {
                const int ip = ch * 32 + lane;
                const bool active = ip < n;
                const SPt P = spt[active ? ip : 0]; const double zp = sz[active ? ip : 0];
                double c3 = 0.0, c2 = 0.0;
                bool hull = false;
                // (the thread's two lists, addressed in the shared window: entry k of T-interleaved 16-bit positions)
                struct { uint32_t base, cbase;
                         __device__ void put(int k, int pos) { asm volatile("st.shared.u16 [%0], %1;" :: "r"(base + 2u * T * (uint32_t)k), "h"((unsigned short)pos) : "memory"); }
                         __device__ int get(int k) const { unsigned short v; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(base + 2u * T * (uint32_t)k) : "memory"); return v; }
                         __device__ void cput(int k, int pos) { asm volatile("st.shared.u16 [%0], %1;" :: "r"(cbase + 2u * T * (uint32_t)k), "h"((unsigned short)pos) : "memory"); }
                         __device__ int cget(int k) const { unsigned short v; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(cbase + 2u * T * (uint32_t)k) : "memory"); return v; } } lst;
                lst.base = pp_smem_addr(lists + tid); lst.cbase = pp_smem_addr(clist + tid);
                st |= star_point<STAR_LCAP, STAR_CCAP, FULL>(g, ip, active, GT_STAR_W0, lst, &hull, [&](int n1, int n2) {
                    const SPt A = spt[n1], B = spt[n2];
                    c3 += tri_area3(P.x, P.y, zp, A.x, A.y, sz[n1], B.x, B.y, sz[n2]);
                    if (want2d) c2 += tri_area3(P.x, P.y, 0.0, A.x, A.y, 0.0, B.x, B.y, 0.0);
                    if (tri_out) {
                        const int o = (int)atomicAdd(&ctl[1], 1u);
                        if (o < p.tri_cap) { tri_out[3 * o] = sid[ip]; tri_out[3 * o + 1] = sid[n1]; tri_out[3 * o + 2] = sid[n2]; }
                    }
                    ++ntri;
                });
                nhull += hull;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) { c3 += __shfl_down_sync(0xffffffffu, c3, o); c2 += __shfl_down_sync(0xffffffffu, c2, o); }
                int nextch = 0;
                if (lane == 0) { csum[2 * ch] = c3; csum[2 * ch + 1] = c2; nextch = (int)atomicAdd(&ctl[3], 1u); }
                ch = __shfl_sync(0xffffffffu, nextch, 0);
            }
        }
        // ---------------------------------------------------------------- the frame's sums (fixed order) and its verdict
        unsigned packed = ((unsigned)ntri << 12) | (unsigned)nhull;          // (n <= 4,094 points here: 2n triangles fit 20 bits)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { packed += __shfl_down_sync(0xffffffffu, packed, o); st |= __shfl_down_sync(0xffffffffu, st, o); }
        __syncthreads();                                                   // (red was the scan's scratch)
        if (lane == 0) red[warp] = __hiloint2double((int)packed, st);
        __syncthreads();
        if (tid == 0) {
            double s3 = 0.0, s2 = 0.0; unsigned pk = 0; int sst = 0;
            for (int w = 0; w < NW; ++w) { pk += (unsigned)__double2hiint(red[w]); sst |= __double2loint(red[w]); }
            if (ok) for (int ch = 0; ch * 32 < n; ++ch) { s3 += csum[2 * ch]; s2 += csum[2 * ch + 1]; }
            const int tris = (int)(pk >> 12);
            // triangle count rule (delaunay_tri.c:608): hull vertices = open stars, or (arcs only) the points on the bounding
            // box when its corners are points -- otherwise the hull size is not known here and the rule is not applied
            const int hulls = FULL ? (int)(pk & 0xfffu) : (int)ctl[5];
            const bool count_ok = (!FULL && ctl[4] < 4u) || tris == 2 * (n - 1) - hulls;
            const bool certified = ok && sst == 0 && count_ok && (!p.tri || tris <= p.tri_cap);
            if (certified) {
                if (p.area) p.area[fr] = s3;
                if (p.area2d && want2d) p.area2d[fr] = s2;
                if (p.status) p.status[fr] = 0;
                if (p.ntri) p.ntri[fr] = tris;
                if (p.areabox) p.areabox[fr] = bx * by;
                if (p.ncorr) p.ncorr[fr] = ncorr_out;
            } else {
                sp.fb_list[atomicAdd(sp.fb_count, 1u)] = fr;
            }
            if (sp.counters) atomicAdd(&sp.counters[certified ? 0 : 1], 1ull);
        }
        __syncthreads();
        if (bulk && tid == 0 && fr + stride < p.nframes) {                  // (the lists are dead: the tile's memory takes the next frame)
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            pp_bulk_load(tile, p.xyz + (size_t)(fr + stride) * natoms * in_dim, frame_bytes, mbar);
        }
    }
}

}  // namespace gt
