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

struct LsList {
    uint32_t* l; uint32_t* c; size_t nt;
    __device__ void put(int k, int pos) { l[(size_t)k * nt] = (uint32_t)pos; }
    __device__ int get(int k) const { return (int)l[(size_t)k * nt]; }
    __device__ void cput(int k, int pos) { c[(size_t)k * nt] = (uint32_t)pos; }
    __device__ int cget(int k) const { return (int)c[(size_t)k * nt]; }
};

// ctl: [0] STAR_* bits of any point | 0x1000 when the overflow list is full, [1] open stars (hull vertices),
// [2] entries of the overflow list: triangles beyond the LS_MAXOWN a point keeps in its own slots (hull vertices own long fans)
__global__ void __launch_bounds__(128) ls_star_kernel(StarGrid<uint32_t, true> g, uint32_t* llist, uint32_t* clist, const double* sz, int want2d,
                                                      uint32_t* own_cnt, uint32_t* own_pairs, uint32_t* ovf, uint32_t ovf_cap, double* a3, double* a2, uint32_t* ctl,
                                                      unsigned long long* chunk_clk) {
    const size_t nt = (size_t)gridDim.x * blockDim.x, gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const int nwarps = (int)(nt >> 5), gwarp = (int)(gtid >> 5);
    LsList lst; lst.l = llist + gtid; lst.c = clist + gtid; lst.nt = nt;
    const int n = g.n;
    // chunks of 32 consecutive positions; a warp's first chunk is its own number, the rest come from a counter, so that the
    // chunks along the hull -- whose searches reach far -- do not pile up on the warps they would statically fall to
    for (int ch = gwarp; (long long)ch * 32 < n;) {
        const int ip = ch * 32 + lane;
        const bool active = ip < n;
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
