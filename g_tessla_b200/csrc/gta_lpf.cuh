// gta_lpf.cuh -- lane-per-frame execution of the triangulation (see lpf_core.cuh) and its area epilogue.
//
// Pipeline for a large batch (>= tens of thousands of frames; smaller batches stay on the CTA-per-frame
// kernel, whose per-frame latency is 20x lower):
//   1. tessellate_kernel in pre-pass mode: load, -corr, sort -> workspace          (gta_kernel.cuh)
//   2. lpf_kernel: every LANE owns a frame and runs the D&C state machine on it     (this file)
//   3. lpf_area_kernel: one warp per frame enumerates the triangles and sums areas  (this file)
// Replaces the same reference loops as the fused kernel (src/gta_tri.c:106-296).
#pragma once
#include "gta_kernel.cuh"

namespace gt {

struct LpfParams {
    LpfWs ws;
    int nframes;
    unsigned int* counter;                 // next frame to hand out
};

__device__ __forceinline__ LFrame lpf_frame(const LpfWs& ws, int fr) {
    LFrame f;
    const size_t s = (size_t)fr;
    // LFrame.pt is declared with 32-byte points for the host emulation; on the device the workspace keeps
    // 16-byte (x, y) points and z apart, and lf_pt is overloaded below through the PT16 flag
    f.pt = reinterpret_cast<LPt*>(ws.pt + s * ws.cap);
    f.rec = ws.rec + s * 2 * (size_t)ws.cap_e;
    f.head = ws.head + s * ws.cap;
    f.n = ws.n[s];
    f.cap_e = ws.cap_e;
    return f;
}

template <int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB) lpf_kernel(const LpfParams p) {
    LState s;
    LFrame f;
    int fr = -1;
    bool have = false, out_of_work = false;
    s.phase = LP_DONE; s.err = 0;
    for (;;) {
        if (!have && !out_of_work) {
            // fetch the next frame that passed the pre-pass
            for (;;) {
                fr = (int)atomicAdd(p.counter, 1u);
                if (fr >= p.nframes) { out_of_work = true; break; }
                if (p.ws.st[fr] != 0) continue;
                f = lpf_frame(p.ws, fr);
                if (f.n < 2) continue;
                lp_begin(f, s);
                have = s.phase != LP_DONE;
                if (!have) continue;
                break;
            }
        }
        if (__all_sync(0xffffffffu, out_of_work && !have)) break;
        if (have) {
            if (!lpf_step(f, s)) {
                p.ws.err[fr] = s.err;
                have = false;
            }
        }
    }
}

// One warp per frame: triangles owned by each sorted vertex, area of each (reference area_tri,
// include/gta_tri.h:91-99), fixed-order warp reduction.
struct LpfAreaParams {
    LpfWs ws; int nframes; int natoms; unsigned flags;
    double* area; double* area2d; int* status; int* ntri; int* tri; int tri_cap;
};
__global__ void __launch_bounds__(128) lpf_area_kernel(const LpfAreaParams p) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= p.nframes) return;
    const int fr = warp;
    int st = p.ws.st[fr];
    const uint32_t e = p.ws.err[fr];
    if (e & GT_ERR_POOL) st |= GTA_B200_ST_POOL;
    if (e & GT_ERR_RANGE) st |= GTA_B200_ST_PRECISION_RANGE;
    if (e & GT_ERR_RING) st |= GTA_B200_ST_INTERNAL;
    double a3 = 0.0, a2 = 0.0; unsigned total = 0;
    if (st == 0) {
        const LFrame f = lpf_frame(p.ws, fr);
        const double* z = p.ws.z + (size_t)fr * p.ws.cap;
        const u16* perm = p.ws.perm + (size_t)fr * p.ws.cap;
        const bool want2d = (p.flags & GTA_B200_2D) != 0;
        unsigned base = 0;
        for (int i0 = 0; i0 < f.n; i0 += 32) {
            const int i = i0 + lane;
            unsigned offs = 0, cnt = 0;
            if (p.tri || p.ntri) {
                cnt = (i < f.n) ? (unsigned)lf_vertex_triangles(f, i, [](int, int) {}) : 0u;
                unsigned inc = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { unsigned t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
                offs = base + inc - cnt;
                base += __shfl_sync(0xffffffffu, inc, 31);
            }
            if (i < f.n) {
                const Pt pi = lf_pt(f, i); const double zi = z[i];
                int* out = p.tri ? p.tri + ((size_t)fr * p.tri_cap) * 3 : nullptr;
                lf_vertex_triangles(f, i, [&](int n1, int n2) {
                    const Pt p1 = lf_pt(f, n1), p2 = lf_pt(f, n2);
                    a3 += tri_area3(pi.x, pi.y, zi, p1.x, p1.y, z[n1], p2.x, p2.y, z[n2]);
                    if (want2d) a2 += tri_area3(pi.x, pi.y, 0.0, p1.x, p1.y, 0.0, p2.x, p2.y, 0.0);
                    if (out && (int)offs < p.tri_cap) { out[3 * offs] = perm[i]; out[3 * offs + 1] = perm[n1]; out[3 * offs + 2] = perm[n2]; }
                    ++offs;
                });
            }
        }
        total = base;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { a3 += __shfl_down_sync(0xffffffffu, a3, o); a2 += __shfl_down_sync(0xffffffffu, a2, o); }
    if (lane == 0) {
        if (st & ~GTA_B200_ST_TOO_FEW_POINTS) { a3 = __longlong_as_double(0x7ff8000000000000ll); a2 = a3; }
        if (p.area) p.area[fr] = a3;
        if (p.area2d && (p.flags & GTA_B200_2D)) p.area2d[fr] = a2;
        if (p.status) p.status[fr] = st;
        if (p.ntri) p.ntri[fr] = st ? 0 : (int)total;
    }
}

}  // namespace gt
