// gta_kernel.cuh -- the batched per-frame kernel: one CTA per frame, whole frame in shared memory.
//
// Replaces, per frame, the body of the reference's OpenMP loop (src/gta_tri.c:106-296):
//   K0  -corr box-edge point generation            gta_tri.c:112-272
//   K1  lexicographic sort + Delaunay D&C           delaunay_tri.c:436, :511-601
//   K2  triangle enumeration + area sum             delaunay_tri.c:606-635, gta_tri.c:368-410
// fused in one launch: coordinates are read from HBM once (coalesced), everything else lives in
// shared memory (<= ~56 B per point; z stays in global memory), and 8-24 bytes per frame go back.
#pragma once
#include <cuda_runtime.h>
#include <float.h>
#include "dt_core.cuh"
#include "dt_coop.cuh"
#include "lpf_core.cuh"
#include "../../include/gta_b200.h"

namespace gt {

// -DGT_PHASE_TIMING: thread 0 of every CTA adds clock64() deltas per phase into p.dbg
// (slots: 0 load, 1 corr, 2 sort, 3 post-sort, 4..19 D&C level d = slot-4, 20 enumerate+area, 21 frames).
#ifdef GT_PHASE_TIMING
#define GT_TICK(slot) do { if (tid == 0 && p.dbg) { long long t_ = clock64(); atomicAdd(&p.dbg[slot], (unsigned long long)(t_ - tick_)); tick_ = t_; } } while (0)
#else
#define GT_TICK(slot) do { } while (0)
#endif

// Global-memory workspace of the throughput path (gta_lpf.cuh): one slot per frame of the batch.
struct LpfWs {
    LPt* pt; LRec* rec;                                      // strides per frame: cap (x, y, z, ring head, input index), 2*cap_e
    int* n; int* st; uint32_t* err;                          // per frame: point count, status bits of the pre-pass, D&C error bits
    int cap, cap_e;
};

struct KParams {
    const double* xyz; int in_dim;            // [nframes][natoms][in_dim]
    const double* box; int box_stride, by_off;
    int nframes, natoms, cap;                  // cap = max points per frame (even)
    int cap_p2;                                // power of two >= cap (sort network width bound)
    double espace; unsigned flags;
    double* area; double* area2d; double* areabox; int* status;
    int* tri; int tri_cap; int* ntri;
    double* corr; int corr_cap; int* ncorr;
    unsigned int* work;                        // global frame counter
    unsigned long long* dbg;                   // phase timers (GT_PHASE_TIMING builds), else unused
    int coop_nodes;                            // levels with at most this many nodes use warp-cooperative merges
    LpfWs ws;                                  // ws.pt != nullptr: pre-pass only (load, -corr, sort), frames go to the workspace
    // Divide-and-conquer pipeline as the fallback of the star path (gta_star.cuh): work on the frames list[0 .. *list_count)
    // of the batch instead of all of them; workspace slot i holds frame list[i], results go to the frame's own index.
    const int* list = nullptr; const unsigned int* list_count = nullptr;
};

__host__ __device__ inline int pool_edges(int cap) { return 3 * cap + 64; }
// dynamic shared memory for a capacity of `cap` points of which `natoms` come from the caller
// (the other cap - natoms slots can only be -corr points, whose z is generated in the kernel)
// (prepass: the build of the kernel that stops after the sort -- it needs the pool's space only for the points in
// input order that alias it)
__host__ __device__ inline size_t pool_bytes_for(int cap, bool prepass) {
    return prepass ? sizeof(Pt) * (size_t)cap : 12 * (size_t)pool_edges(cap);
}
__host__ __device__ inline size_t smem_bytes_for(int cap, int cap_p2, int natoms, bool prepass = false) {
    size_t b = 0;
    b += sizeof(Pt) * (size_t)cap;             // sorted xy
    b += pool_bytes_for(cap, prepass);         // half-edge pool (dst, nxt, prv: u16 x 2 per edge each)
    b += 2 * (size_t)cap_p2;                   // perm
    b += 2 * (size_t)cap;                      // head
    b = (b + 7) & ~(size_t)7;
    b += 8 * (size_t)(cap > natoms ? cap - natoms : 0);   // z of the -corr points
    b = (b + 15) & ~(size_t)15;
    b += 16 * 4 + 64 * 8 + 32 * 4;             // ctl, reduction scratch, scan scratch
    return b;
}

__device__ __forceinline__ unsigned long long okey(double v) {   // order-preserving bits
    unsigned long long b = (unsigned long long)__double_as_longlong(v + 0.0);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

template <int T>
__device__ __forceinline__ unsigned block_exclusive_scan(unsigned v, unsigned* scratch, unsigned* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { unsigned t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (T > 32) {
        if (lane == 31) scratch[warp] = inc;
        __syncthreads();
        unsigned base = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < T / 32; ++w) { unsigned s = scratch[w]; if (w < warp) base += s; tot += s; }
        __syncthreads();
        *total = tot;
        return base + inc - v;
    } else {
        *total = __shfl_sync(0xffffffffu, inc, 31);
        return inc - v;
    }
}

// PREPASS = true: the lane-per-frame path's first launch -- load, -corr, sort, write the frame to its workspace slot
// (p.ws) and stop. Its own build: without the D&C it needs a third of the registers and 2/3 of the shared memory, so
// twice the threads work on the sort.
template <int T, int MINB, bool PREPASS = false>
__global__ void __launch_bounds__(T, MINB) tessellate_kernel(const KParams p) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x;
    const int cap = p.cap, cap_e = pool_edges(cap);
    Pt* pt = reinterpret_cast<Pt*>(smem);
    unsigned char* pool = reinterpret_cast<unsigned char*>(pt + cap);
    u16* dst = reinterpret_cast<u16*>(pool);
    u16* nxt = dst + 2 * cap_e;
    u16* prv = nxt + 2 * cap_e;
    Pt* tmp_xy = reinterpret_cast<Pt*>(pool);                  // setup-time alias of the pool (input order)
    u16* perm = reinterpret_cast<u16*>(pool + pool_bytes_for(cap, PREPASS));
    u16* head = perm + p.cap_p2;
    size_t off = (size_t)(reinterpret_cast<unsigned char*>(head + cap) - smem);
    off = (off + 7) & ~(size_t)7;
    double* zcorr = reinterpret_cast<double*>(smem + off);     // z of the appended -corr points
    off += 8 * (size_t)(cap > p.natoms ? cap - p.natoms : 0);
    off = (off + 15) & ~(size_t)15;
    uint32_t* ctl = reinterpret_cast<uint32_t*>(smem + off);    // [0] bump [1] stack [2] err [3] frame [4] status
    double* red = reinterpret_cast<double*>(ctl + 16);
    unsigned* scan = reinterpret_cast<unsigned*>(red + 64);
    // -corr scratch aliases the (not yet filled) sorted-xy region
    unsigned long long* ckey = reinterpret_cast<unsigned long long*>(pt);

    const bool corr = (p.flags & GTA_B200_CORRECT) != 0, want2d = (p.flags & GTA_B200_2D) != 0;
    const int natoms = p.natoms;

    for (;;) {
        __syncthreads();
        if (tid == 0) { ctl[3] = atomicAdd(p.work, 1u); ctl[4] = 0; ctl[0] = 0; ctl[1] = GT_NIL; ctl[2] = 0; }
        __syncthreads();
        const int fr = (int)ctl[3];
        if (fr >= p.nframes) break;

#ifdef GT_PHASE_TIMING
        long long tick_ = clock64();
#endif
        // ------------------------------------------------------------ load (coalesced)
        const double* src = p.xyz + (size_t)fr * natoms * p.in_dim;
        bool bad = false;
        // z stays in global memory (L2-resident after this pass): it is needed again only by the
        // -corr z interpolation and the area epilogue.
        if (p.in_dim == 3) {
            for (int k = tid; k < 3 * natoms; k += T) {
                double v = src[k];
                bad |= !isfinite(v);
                int a = k / 3, c = k - 3 * a;
                if (c == 0) tmp_xy[a].x = v; else if (c == 1) tmp_xy[a].y = v;
            }
        } else {
            for (int k = tid; k < 2 * natoms; k += T) {
                double v = src[k];
                bad |= !isfinite(v);
                if (k & 1) tmp_xy[k >> 1].y = v; else tmp_xy[k >> 1].x = v;
            }
        }
        auto atom_z = [&](int a) -> double { return p.in_dim == 3 ? src[3 * a + 2] : 0.0; };
        double bx = 0.0, by = 0.0;
        if (p.box) { bx = p.box[(size_t)fr * p.box_stride]; by = p.box[(size_t)fr * p.box_stride + p.by_off]; }
        if (bad) atomicOr(&ctl[4], GTA_B200_ST_NONFINITE);
        int n = natoms;

        __syncthreads(); GT_TICK(0);
        // ------------------------------------------------------------ K0: -corr points
        if (corr) {
            const int nx = (int)(bx / p.espace), ny = (int)(by / p.espace);      // gta_tri.c:112-113
            const bool boxok = isfinite(bx) && isfinite(by) && nx >= 0 && ny >= 0 && natoms >= 1;
            // a tiny espace or a huge box: more edge points than any frame can hold -- and 4 + 2 nx + 2 ny must not wrap
            // (the same bound as gta_b200_max_points_for on the host)
            const bool toomany = boxok && (nx > (1 << 20) || ny > (1 << 20));
            const int ncorr = (boxok && !toomany) ? 4 + 2 * nx + 2 * ny : 0;
            const int K = ncorr + 4;                                             // 4 corners + 2(nx+1) + 2(ny+1) slots
            int* cidx = reinterpret_cast<int*>(ckey + K);
            if (!boxok) { if (tid == 0) atomicOr(&ctl[4], natoms < 1 ? GTA_B200_ST_TOO_FEW_POINTS : GTA_B200_ST_NONFINITE); }
            else if (toomany || natoms + ncorr > cap) { if (tid == 0) atomicOr(&ctl[4], GTA_B200_ST_CAPACITY); }
            else {
                const int oymin = 4, oymax = 4 + (nx + 1), oxmin = 4 + 2 * (nx + 1), oxmax = oxmin + (ny + 1);
                const unsigned long long kmin0 = okey((double)FLT_MAX), kmax0 = okey((double)FLT_MIN);   // gta_tri.c:119-120,137-144
                for (int s = tid; s < K; s += T) {
                    bool ismin = (s == 0 || s == 2) || (s >= oymin && s < oymax) || (s >= oxmin && s < oxmax);
                    ckey[s] = ismin ? kmin0 : kmax0;
                    cidx[s] = 0x7fffffff;
                }
                __syncthreads();
                // pass 1: extreme values per slot (gta_tri.c:153-202)
                for (int j = tid; j < natoms; j += T) {
                    double X = tmp_xy[j].x, Y = tmp_xy[j].y;
                    double d0 = X * X + Y * Y, dY = by - Y, d1 = X * X + dY * dY;
                    double fx = (X / bx) * nx, fy = (Y / by) * ny;
                    if (!(fx > -1.0 && fx < (double)(nx + 1) && fy > -1.0 && fy < (double)(ny + 1))) {
                        atomicOr(&ctl[4], GTA_B200_ST_OUT_OF_BOX);
                        continue;
                    }
                    int xi = (int)fx, yi = (int)fy;
                    atomicMin(&ckey[0], okey(d0)); atomicMax(&ckey[1], okey(d0));
                    atomicMin(&ckey[2], okey(d1)); atomicMax(&ckey[3], okey(d1));
                    atomicMin(&ckey[oymin + xi], okey(Y)); atomicMax(&ckey[oymax + xi], okey(Y));
                    atomicMin(&ckey[oxmin + yi], okey(X)); atomicMax(&ckey[oxmax + yi], okey(X));
                }
                __syncthreads();
                // pass 2: lowest atom index attaining the extreme (strict < / > keeps the first one)
                if (!(ctl[4] & GTA_B200_ST_OUT_OF_BOX)) {
                    for (int j = tid; j < natoms; j += T) {
                        double X = tmp_xy[j].x, Y = tmp_xy[j].y;
                        double d0 = X * X + Y * Y, dY = by - Y, d1 = X * X + dY * dY;
                        int xi = (int)((X / bx) * nx), yi = (int)((Y / by) * ny);
                        const double fmx = (double)FLT_MAX, fmn = (double)FLT_MIN;
                        if (d0 < fmx && okey(d0) == ckey[0]) atomicMin(&cidx[0], j);
                        if (d0 > fmn && okey(d0) == ckey[1]) atomicMin(&cidx[1], j);
                        if (d1 < fmx && okey(d1) == ckey[2]) atomicMin(&cidx[2], j);
                        if (d1 > fmn && okey(d1) == ckey[3]) atomicMin(&cidx[3], j);
                        if (Y < fmx && okey(Y) == ckey[oymin + xi]) atomicMin(&cidx[oymin + xi], j);
                        if (Y > fmn && okey(Y) == ckey[oymax + xi]) atomicMin(&cidx[oymax + xi], j);
                        if (X < fmx && okey(X) == ckey[oxmin + yi]) atomicMin(&cidx[oxmin + yi], j);
                        if (X > fmn && okey(X) == ckey[oxmax + yi]) atomicMin(&cidx[oxmax + yi], j);
                    }
                }
                __syncthreads();
                for (int s = tid; s < K; s += T) if (cidx[s] == 0x7fffffff) cidx[s] = 0;      // index arrays start at 0 (gta_tri.c:146-149)
                __syncthreads();
                // corner z (gta_tri.c:209-212) and the appended points in the reference's order (:220-272)
                const double zc = (atom_z(cidx[0]) + atom_z(cidx[1]) + atom_z(cidx[2]) + atom_z(cidx[3])) / 4.0;
                for (int s = tid; s < ncorr; s += T) {
                    double X, Y, Z;
                    if (s < 4) {
                        X = (s == 1 || s == 2) ? bx : 0.0; Y = (s >= 2) ? by : 0.0; Z = zc;
                    } else if (s < 4 + 2 * nx) {
                        int j = (s - 4) >> 1;
                        int imin = cidx[oymin + j], imax = cidx[oymax + j];
                        double dist1 = tmp_xy[imin].y, dist2 = by - tmp_xy[imax].y, dist = dist1 + dist2;
                        double zmin = atom_z(imin), zmax = atom_z(imax);
                        Z = zmin - (dist1 / dist) * zmin + zmax - (dist2 / dist) * zmax;
                        X = j * p.espace + p.espace / 2; Y = ((s - 4) & 1) ? by : 0.0;
                    } else {
                        int j = (s - 4 - 2 * nx) >> 1;
                        int imin = cidx[oxmin + j], imax = cidx[oxmax + j];
                        double dist1 = tmp_xy[imin].x, dist2 = bx - tmp_xy[imax].x, dist = dist1 + dist2;
                        double zmin = atom_z(imin), zmax = atom_z(imax);
                        Z = zmin - (dist1 / dist) * zmin + zmax - (dist2 / dist) * zmax;
                        Y = j * p.espace + p.espace / 2; X = ((s - 4 - 2 * nx) & 1) ? bx : 0.0;
                    }
                    tmp_xy[natoms + s].x = X; tmp_xy[natoms + s].y = Y; zcorr[s] = Z;
                    if (p.corr && s < p.corr_cap) {
                        double* o = p.corr + ((size_t)fr * p.corr_cap + s) * 3;
                        o[0] = X; o[1] = Y; o[2] = Z;
                    }
                }
                n = natoms + ncorr;
            }
            if (tid == 0 && p.ncorr) p.ncorr[fr] = ncorr;
        } else if (natoms > cap) {
            if (tid == 0) atomicOr(&ctl[4], GTA_B200_ST_CAPACITY);
        }
        if (n < 2 && tid == 0) atomicOr(&ctl[4], GTA_B200_ST_TOO_FEW_POINTS);    // delaunay_tri.c:416-421
        __syncthreads();

        GT_TICK(1);
        double a3 = 0.0, a2 = 0.0;
        unsigned ntri_total = 0;
        if (ctl[4] == 0) {
            // -------------------------------------------------------- K1a: sort positions by (x, y)
            int np2 = 2; while (np2 < n) np2 <<= 1;
            for (int i = tid; i < np2; i += T) perm[i] = (i < n) ? (u16)i : (u16)GT_NIL;
            __syncthreads();
            for (int k = 2; k <= np2; k <<= 1) {
                for (int j = k >> 1; j > 0; j >>= 1) {
                    for (int t = tid; t < (np2 >> 1); t += T) {
                        int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));      // lower index of the pair
                        int l = i | j;
                        unsigned a = perm[i], b = perm[l];
                        // strict (x, y) order, GT_NIL padding = +inf; exact duplicates tie and stay put
                        auto greater = [&](unsigned u, unsigned v) -> bool {
                            if (u == GT_NIL) return v != GT_NIL;
                            if (v == GT_NIL) return false;
                            Pt pu = tmp_xy[u], pv = tmp_xy[v];
                            return (pu.x > pv.x) || (pu.x == pv.x && pu.y > pv.y);
                        };
                        bool up = ((i & k) == 0);
                        if (up ? greater(a, b) : greater(b, a)) { perm[i] = (u16)b; perm[l] = (u16)a; }
                    }
                    __syncthreads();
                }
            }
            GT_TICK(2);
            // gather into sorted order (the -corr scratch in `pt` is dead now)
            for (int i = tid; i < n; i += T) pt[i] = tmp_xy[perm[i]];
            __syncthreads();
            // The reference comparator (delaunay_tri.c:115-123) treats |dx| < 1e-12 as an x-tie and then
            // orders by y. For exact-x ties that is the order just produced. A run of sorted points whose
            // consecutive x gaps are all < 1e-12 but not all 0 (e.g. a float32 atom at x = 2.0 next to the
            // -corr points at 2*0.8+0.4 = 2.0000000000000004) must be re-ordered by y alone; the thread
            // owning the run's first element does it (runs are a handful of points). If the run is wider
            // than 1e-12 the comparator is not transitive and qsort's result is unspecified: flagged.
            for (int i = tid; i < n; i += T) {
                auto tie = [&](int a, int b) { double d = pt[a].x - pt[b].x; return d < 1e-12 && d > -1e-12; };
                if ((i == 0 || !tie(i, i - 1)) && i + 1 < n && tie(i + 1, i)) {
                    int e = i + 1; bool inexact = false;
                    while (e < n && tie(e, e - 1)) { inexact |= (pt[e].x != pt[e - 1].x); ++e; }
                    if (inexact) {
                        if (!tie(e - 1, i)) atomicOr(&ctl[4], GTA_B200_ST_NEAR_TIE_X);
                        for (int a = i + 1; a < e; ++a) {          // insertion sort of [i, e) by y
                            Pt pa = pt[a]; u16 oa = perm[a];
                            int b = a - 1;
                            while (b >= i && pt[b].y > pa.y) { pt[b + 1] = pt[b]; perm[b + 1] = perm[b]; --b; }
                            pt[b + 1] = pa; perm[b + 1] = oa;
                        }
                    }
                }
            }
            __syncthreads();
            // duplicates within 1e-12 in x and y corrupt the reference's vertex array (delaunay_tri.c:440-452)
            for (int i = tid + 1; i < n; i += T) {
                double dx = pt[i].x - pt[i - 1].x, dy = pt[i].y - pt[i - 1].y;
                if (dx < 1e-12 && dx > -1e-12 && dy < 1e-12 && dy > -1e-12) atomicOr(&ctl[4], GTA_B200_ST_DUPLICATE);
            }
            for (int i = tid; i < n; i += T) head[i] = (u16)GT_NIL;
            __syncthreads();
        }
        GT_TICK(3);
        if (PREPASS || p.ws.pt) {
            // pre-pass of the lane-per-frame path: sorted points, z, permutation and the status go to the frame's
            // workspace slot; the triangulation and the area sum are separate launches (gta_lpf.cuh)
            const size_t slot = (size_t)fr;
            if (ctl[4] == 0) {
                LPt* wpt = p.ws.pt + slot * p.ws.cap;
                for (int i = tid; i < n; i += T) {
                    const int o = perm[i];
                    LPt r; r.x = pt[i].x; r.y = pt[i].y; r.z = o < natoms ? atom_z(o) : zcorr[o - natoms];
                    r.head = (u16)GT_NIL; r.perm = (u16)o; r.pad = 0;
                    wpt[i] = r;
                }
            }
            if (tid == 0) {
                p.ws.n[slot] = n; p.ws.st[slot] = (int)ctl[4]; p.ws.err[slot] = 0;
                if (p.areabox) p.areabox[fr] = bx * by;
            }
            continue;
        }
        if (ctl[4] == 0) {
            // -------------------------------------------------------- K1b: bottom-up D&C
            Frame f;
            f.pt = pt; f.head = head; f.dst = dst; f.nxt = nxt; f.prv = prv;
            f.bump = &ctl[0]; f.stack = &ctl[1]; f.err = &ctl[2]; f.n = n; f.cap_e = cap_e;
            Lane ln; ln.free_head = GT_NIL;
            const int D = tree_depth(n);
            for (int d = D; d >= 0; --d) {
                const int nodes = 1 << d;
                if (nodes > p.coop_nodes) {
                    // deep levels: many short merges, one lane each. Lanes of one warp that walk
                    // different merges diverge and serialise, so nodes are dealt to warps round-robin
                    // (node k -> warp k % W, lane k / W): every warp carries as few of them as possible.
                    constexpr int W = T / 32;
                    const int slot = (tid & 31) * W + (tid >> 5);
                    for (int k = slot; k < nodes; k += T) process_node(f, ln, d, k);
                } else {
                    // upper levels: few long merges, one warp each (dt_coop.cuh)
                    for (int k = tid >> 5; k < nodes; k += T / 32) {
                        int ia, ib;
                        if (!tree_node(n, d, k, &ia, &ib) || ib - ia < 1) continue;
                        if (ib - ia < 3) { if ((tid & 31) == 0) leaf(f, ln, ia, ib); }
                        else coop_merge(f, ln, (ia + ib) / 2, p.dbg);
                    }
                }
                __syncthreads();
                flush_lane(f, ln);
                __syncthreads();
                GT_TICK(4 + d);
                if (ctl[2]) break;
            }
            // -------------------------------------------------------- K2: enumerate + area
            if (ctl[2] == 0) {
                unsigned base = 0;
                for (int i0 = 0; i0 < n; i0 += T) {
                    const int i = i0 + tid;
                    unsigned offs = 0, tot = 0;
                    if (p.tri || p.ntri) {
                        unsigned cnt = (i < n) ? (unsigned)vertex_triangles(f, i, [](int, int) {}) : 0u;
                        offs = block_exclusive_scan<T>(cnt, scan, &tot) + base;
                    }
                    if (i < n) {
                        auto zof = [&](int v) -> double { int o = perm[v]; return o < natoms ? atom_z(o) : zcorr[o - natoms]; };
                        const Pt pi = pt[i]; const double zi = zof(i);
                        const int oi = perm[i];
                        int* out = p.tri ? p.tri + ((size_t)fr * p.tri_cap) * 3 : nullptr;
                        vertex_triangles(f, i, [&](int n1, int n2) {
                            Pt p1 = pt[n1], p2 = pt[n2];
                            a3 += tri_area3(pi.x, pi.y, zi, p1.x, p1.y, zof(n1), p2.x, p2.y, zof(n2));
                            if (want2d) a2 += tri_area3(pi.x, pi.y, 0.0, p1.x, p1.y, 0.0, p2.x, p2.y, 0.0);
                            if (out && (int)offs < p.tri_cap) {
                                out[3 * offs] = oi; out[3 * offs + 1] = perm[n1]; out[3 * offs + 2] = perm[n2];
                            }
                            ++offs;
                        });
                    }
                    base += tot;
                }
                ntri_total = base;
            }
        }
        GT_TICK(20);
#ifdef GT_PHASE_TIMING
        if (tid == 0 && p.dbg) atomicAdd(&p.dbg[21], 1ull);
#endif
        // ------------------------------------------------------------ block reduction (fixed order)
        {
            const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                a3 += __shfl_down_sync(0xffffffffu, a3, o);
                a2 += __shfl_down_sync(0xffffffffu, a2, o);
            }
            if (lane == 0) { red[warp] = a3; red[32 + warp] = a2; }
            __syncthreads();
            if (tid == 0) {
                double s3 = 0.0, s2 = 0.0;
                for (int w = 0; w < T / 32; ++w) { s3 += red[w]; s2 += red[32 + w]; }
                unsigned e = ctl[2];
                int st = (int)ctl[4];
                if (e & GT_ERR_POOL) st |= GTA_B200_ST_POOL;
                if (e & GT_ERR_RANGE) st |= GTA_B200_ST_PRECISION_RANGE;
                if (e & GT_ERR_RING) st |= GTA_B200_ST_INTERNAL;
                if (st & ~GTA_B200_ST_TOO_FEW_POINTS) { s3 = __longlong_as_double(0x7ff8000000000000ll); s2 = s3; }
                if (p.area) p.area[fr] = s3;
                if (p.area2d && want2d) p.area2d[fr] = s2;
                if (p.areabox) p.areabox[fr] = bx * by;                       // gta_tri.c:116, :316
                if (p.status) p.status[fr] = st;
                if (p.ntri) p.ntri[fr] = st ? 0 : (int)ntri_total;
            }
        }
    }
}

// Array evaluation of the predicates (self-test hook; also exercises the exact stage directly).
__global__ void orient2d_signs_kernel(const double* pts, int n, int* out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* c = pts + 6 * (size_t)i;
    out[i] = orient2d_sign(c[0], c[1], c[2], c[3], c[4], c[5]);
}
__global__ void incircle_signs_kernel(const double* pts, int n, int* out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* c = pts + 8 * (size_t)i;
    out[i] = incircle_sign(c[0], c[1], c[2], c[3], c[4], c[5], c[6], c[7]);
}

}  // namespace gt
