// gta_prepass.cuh -- first stage of the throughput path: load, -corr, sort -> the frame's sorted point records.
//
// Replaces, per frame, the head of the reference's loop body:
//   -corr box-edge point generation   src/gta_tri.c:112-272   (K0; same arithmetic and emission order as gta_kernel.cuh)
//   qsort by compareVerts             src/delaunay_tri.c:115-123, :436
//   duplicate check                   src/delaunay_tri.c:440-452 (detect and flag, see include/gta_b200.h)
// What is B200-specific:
//   * the frame's coordinate tile (natoms x 24 B, AoS as the caller has it) is brought into shared memory by ONE bulk
//     asynchronous copy (cp.async.bulk + mbarrier: the TMA unit, no register staging), double-buffered: a persistent CTA
//     walks frames blockIdx.x, blockIdx.x + gridDim.x, ... and the copy of its next frame runs while it sorts this one;
//   * the sort is a block radix (bucket) sort on the x coordinate instead of a bitonic network: points are binned by a
//     monotone map of x into 2,048 buckets (shared-memory histogram + scan + scatter), and each bucket -- a handful of
//     points, or one lattice column of equal x -- is put into exact (x, y) order by one thread. 3 barriers instead of 66
//     compare-exchange passes, and no data-dependent bank conflicts on 16-byte point reads. A frame whose points pile
//     up in one bucket falls back to the bitonic network (same result).
#pragma once
#include "gta_kernel.cuh"
#include "gta_corr.cuh"

namespace gt {

constexpr int PP_BUCKETS = 2048;
constexpr int PP_BIG_BUCKET = 512;         // more points than this in one bucket: bitonic fallback for the frame

struct PrepassSmem {
    size_t tile, corr, bkt, perm, perm2, hist, pts, ctl, total;
    __host__ __device__ PrepassSmem(int cap, int natoms, int in_dim) {
        size_t o = 0;
        const size_t tile_bytes = (((size_t)natoms * in_dim * 8) + 15) & ~(size_t)15;
        tile = o; o += 2 * tile_bytes;                                      // two coordinate tiles (bulk-copy destinations)
        corr = o; o += 24 * (size_t)(cap > natoms ? cap - natoms : 0); o = (o + 15) & ~(size_t)15;      // appended -corr points (x, y, z)
        pts = o; o += sizeof(Pt) * (size_t)cap;                             // sorted (x, y)
        hist = o; o += 4 * (size_t)(PP_BUCKETS + 1);                        // bucket counters / cursors
        int cap_p2 = 2; while (cap_p2 < cap) cap_p2 <<= 1;
        bkt = o; o += 2 * (size_t)cap_p2;
        perm = o; o += 2 * (size_t)cap_p2;
        perm2 = o; o += 2 * (size_t)cap_p2;
        o = (o + 15) & ~(size_t)15;
        ctl = o; o += 160 * 4 + 2 * 8;                                      // status words, reduction scratch, two mbarriers
        total = (o + 15) & ~(size_t)15;
    }
};

__device__ __forceinline__ uint32_t pp_smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void pp_mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(pp_smem_addr(bar)), "r"(count) : "memory");
}
// one thread: expect `bytes` on the barrier and start the bulk copy global -> shared that will deliver them
__device__ __forceinline__ void pp_bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(pp_smem_addr(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(pp_smem_addr(dst)), "l"(src), "r"(bytes), "r"(pp_smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void pp_mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done)                                            // (try_wait suspends the thread in hardware for a while: not a busy poll)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(pp_smem_addr(bar)), "r"(parity) : "memory");
}

// -DGT_PHASE_TIMING: thread 0 of every CTA adds clock64() deltas per section into p.dbg[0..9]
// (0 tile wait, 1 finite check, 2 -corr, 3 min/max + histogram, 4 scan, 5 scatter + bucket order, 6 gather, 7 near-tie + duplicates, 8 records, 9 frames)
#ifdef GT_PHASE_TIMING
#define PP_TICK(slot) do { if (tid == 0 && p.dbg) { const long long t_ = clock64(); atomicAdd(&p.dbg[slot], (unsigned long long)(t_ - tick_)); tick_ = t_; } } while (0)
#else
#define PP_TICK(slot) do { } while (0)
#endif

template <int T>
__global__ void __launch_bounds__(T) prepass_kernel(const KParams p) {
    extern __shared__ __align__(128) unsigned char psm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cap = p.cap, natoms = p.natoms, in_dim = p.in_dim;
    const PrepassSmem L(cap, natoms, in_dim);
    const size_t tile_bytes = (((size_t)natoms * in_dim * 8) + 15) & ~(size_t)15;
    double* corr3 = reinterpret_cast<double*>(psm + L.corr);
    Pt* pt = reinterpret_cast<Pt*>(psm + L.pts);
    uint32_t* hist = reinterpret_cast<uint32_t*>(psm + L.hist);
    u16* bkt = reinterpret_cast<u16*>(psm + L.bkt);
    u16* perm = reinterpret_cast<u16*>(psm + L.perm);
    u16* perm2 = reinterpret_cast<u16*>(psm + L.perm2);
    uint32_t* ctl = reinterpret_cast<uint32_t*>(psm + L.ctl);               // [0] status, [1] big bucket seen, [8..8+2*W) reduction scratch (as doubles)
    double* red = reinterpret_cast<double*>(ctl + 8);
    uint64_t* mbar = reinterpret_cast<uint64_t*>(ctl + 160);
    unsigned long long* ckey = reinterpret_cast<unsigned long long*>(psm + L.pts);       // -corr scratch: 12 B per slot in the (not yet filled) sorted-xy region

    const bool corr = (p.flags & GTA_B200_CORRECT) != 0;
    const uint32_t frame_bytes = (uint32_t)((size_t)natoms * in_dim * 8);
    // the bulk copy wants 16-byte aligned source, destination and size; otherwise the tile is loaded with plain loads
    const bool bulk = (frame_bytes % 16 == 0) && ((reinterpret_cast<uintptr_t>(p.xyz) & 15) == 0) && frame_bytes > 0;

    if (tid == 0) { pp_mbar_init(&mbar[0], 1); pp_mbar_init(&mbar[1], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    const int first = blockIdx.x, stride = gridDim.x;
    // (fallback of the star path: the frames list[0 .. *list_count) of the batch; workspace slot = position in the list)
    const int nwork = p.list ? (int)*p.list_count : p.nframes;
    auto frame_of = [&](int i) -> int { return p.list ? p.list[i] : i; };
    if (bulk && tid == 0 && first < nwork)
        pp_bulk_load(psm + L.tile, p.xyz + (size_t)frame_of(first) * natoms * in_dim, frame_bytes, &mbar[0]);

    int it = 0;
    for (int wi = first; wi < nwork; wi += stride, ++it) {
        const int fr = frame_of(wi);
        const int buf = it & 1;
        const double* tile = reinterpret_cast<const double*>(psm + L.tile + (size_t)buf * tile_bytes);
        // ---- the next frame's tile starts loading now (its buffer was last read two frames ago, before a barrier)
        const int nxt = wi + stride;
        if (bulk && tid == 0 && nxt < nwork) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            pp_bulk_load(psm + L.tile + (size_t)(buf ^ 1) * tile_bytes, p.xyz + (size_t)frame_of(nxt) * natoms * in_dim, frame_bytes, &mbar[buf ^ 1]);
        }
#ifdef GT_PHASE_TIMING
        long long tick_ = clock64();
#endif
        if (tid == 0) { ctl[0] = 0; ctl[1] = 0; }
        if (bulk) pp_mbar_wait(&mbar[buf], (uint32_t)((it >> 1) & 1));
        else {
            const double* src = p.xyz + (size_t)fr * natoms * in_dim;
            double* dstt = const_cast<double*>(tile);
            for (int k = tid; k < natoms * in_dim; k += T) dstt[k] = src[k];
        }
        __syncthreads();
        PP_TICK(0);
        bool bad = false;
        for (int k = tid; k < natoms * in_dim; k += T) bad |= !isfinite(tile[k]);
        if (bad) atomicOr(&ctl[0], GTA_B200_ST_NONFINITE);
        auto ax = [&](int a) -> double { return tile[(size_t)a * in_dim]; };
        auto ay = [&](int a) -> double { return tile[(size_t)a * in_dim + 1]; };
        auto az = [&](int a) -> double { return in_dim == 3 ? tile[(size_t)a * 3 + 2] : 0.0; };
        double bx = 0.0, by = 0.0;
        if (p.box) { bx = p.box[(size_t)fr * p.box_stride]; by = p.box[(size_t)fr * p.box_stride + p.by_off]; }
        int n = natoms, ncorr_out = 0;
        __syncthreads();
        PP_TICK(1);

        // ------------------------------------------------------------ K0: -corr points (gta_tri.c:112-272; gta_corr.cuh)
        if (corr) {
            n = corr_points<T>(tile, in_dim, natoms, cap, bx, by, p.espace, ckey, red, ctl, corr3, p.corr, p.corr_cap, fr, &ncorr_out);
        } else if (natoms > cap) {
            if (tid == 0) atomicOr(&ctl[0], GTA_B200_ST_CAPACITY);
        }
        if (n < 2 && tid == 0) atomicOr(&ctl[0], GTA_B200_ST_TOO_FEW_POINTS);    // delaunay_tri.c:416-421
        __syncthreads();
        PP_TICK(2);
        auto px = [&](int i) -> double { return i < natoms ? ax(i) : corr3[3 * (i - natoms)]; };
        auto py = [&](int i) -> double { return i < natoms ? ay(i) : corr3[3 * (i - natoms) + 1]; };
        auto pz = [&](int i) -> double { return i < natoms ? az(i) : corr3[3 * (i - natoms) + 2]; };

        if (ctl[0] == 0) {
            // -------------------------------------------------------- sort positions by (x, y): bucket by x, then exact order inside the buckets
            double lo = 1.7976931348623157e308, hi = -1.7976931348623157e308;
            for (int i = tid; i < n; i += T) { const double x = px(i); lo = fmin(lo, x); hi = fmax(hi, x); }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o)); hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o)); }
            if (lane == 0) { red[2 * warp] = lo; red[2 * warp + 1] = hi; }
            for (int b = tid; b <= PP_BUCKETS; b += T) hist[b] = 0;
            __syncthreads();
#pragma unroll
            for (int w = 0; w < T / 32; ++w) { lo = fmin(lo, red[2 * w]); hi = fmax(hi, red[2 * w + 1]); }
            // monotone map x -> bucket: x1 <= x2 => (x1 - lo) <= (x2 - lo) => products and truncations in the same order
            const double scale = hi > lo ? (double)(PP_BUCKETS - 1) / (hi - lo) : 0.0;
            for (int i = tid; i < n; i += T) {
                int b = (int)((px(i) - lo) * scale);
                b = b < 0 ? 0 : (b > PP_BUCKETS - 1 ? PP_BUCKETS - 1 : b);
                bkt[i] = (u16)b;
                atomicAdd(&hist[b], 1u);
            }
            __syncthreads();
            PP_TICK(3);
            // exclusive scan of the bucket counts (PP_BUCKETS / T per thread)
            {
                constexpr int PER = PP_BUCKETS / T;
                uint32_t c[PER], sum = 0;
#pragma unroll
                for (int k = 0; k < PER; ++k) { c[k] = hist[tid * PER + k]; sum += c[k]; if (c[k] > PP_BIG_BUCKET) ctl[1] = 1; }
                uint32_t inc = sum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
                uint32_t* wsum = reinterpret_cast<uint32_t*>(red);
                __syncthreads();
                if (lane == 31) wsum[warp] = inc;
                __syncthreads();
                uint32_t base = 0;
#pragma unroll
                for (int w = 0; w < T / 32; ++w) if (w < warp) base += wsum[w];
                uint32_t run = base + inc - sum;
#pragma unroll
                for (int k = 0; k < PER; ++k) { hist[tid * PER + k] = run; run += c[k]; }
                if (tid == T - 1) hist[PP_BUCKETS] = run;
            }
            __syncthreads();
            PP_TICK(4);
            const bool big = ctl[1] != 0;
            if (!big) {
                // scatter into the buckets (cursor = the bucket's start, advanced by atomics); then every point finds its rank among
                // the members of its bucket by exact (x, y) comparison (equal points -- duplicates, flagged below -- by index): most
                // buckets hold one or two points, a lattice column or the -corr points of a box edge (equal x) a few dozen, and
                // a thread per POINT keeps those parallel where a thread per bucket would walk them alone
                for (int i = tid; i < n; i += T) perm2[atomicAdd(&hist[bkt[i]], 1u)] = (u16)i;
                __syncthreads();
                // after the scatter hist[b] = end of bucket b = start of bucket b + 1
                for (int a = tid; a < n; a += T) {
                    const int i = perm2[a], b = bkt[i];
                    const int e = (int)hist[b], s = b ? (int)hist[b - 1] : 0;
                    int rank = 0;
                    if (e - s > 1) {
                        const double xi = px(i), yi = py(i);
                        for (int q = s; q < e; ++q) {
                            const int j = perm2[q];
                            const double xj = px(j), yj = py(j);
                            rank += (xj < xi) || (xj == xi && (yj < yi || (yj == yi && j < i)));
                        }
                    }
                    perm[s + rank] = (u16)i;
                }
            } else {
                // a frame with a crowded bucket (thousands of points on a few distinct x): the bitonic network, any distribution
                int np2 = 2; while (np2 < n) np2 <<= 1;
                for (int i = tid; i < np2; i += T) perm[i] = (i < n) ? (u16)i : (u16)GT_NIL;
                __syncthreads();
                for (int k = 2; k <= np2; k <<= 1) {
                    for (int j = k >> 1; j > 0; j >>= 1) {
                        for (int t = tid; t < (np2 >> 1); t += T) {
                            const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), l = i | j;
                            const unsigned a = perm[i], b = perm[l];
                            auto greater = [&](unsigned u, unsigned v) -> bool {
                                if (u == GT_NIL) return v != GT_NIL;
                                if (v == GT_NIL) return false;
                                const double ux = px(u), vx = px(v);
                                return (ux > vx) || (ux == vx && py(u) > py(v));
                            };
                            const bool up = ((i & k) == 0);
                            if (up ? greater(a, b) : greater(b, a)) { perm[i] = (u16)b; perm[l] = (u16)a; }
                        }
                        __syncthreads();
                    }
                }
            }
            __syncthreads();
            PP_TICK(5);
            for (int i = tid; i < n; i += T) { const int o = perm[i]; pt[i].x = px(o); pt[i].y = py(o); }
            __syncthreads();
            PP_TICK(6);
            // The reference comparator (delaunay_tri.c:115-123) treats |dx| < 1e-12 as an x-tie and then orders by y: runs of
            // near-equal x that are not exact ties are re-ordered by y alone (see gta_kernel.cuh for the full story).
            for (int i = tid; i < n; i += T) {
                auto tie = [&](int a, int b) { const double d = pt[a].x - pt[b].x; return d < 1e-12 && d > -1e-12; };
                if ((i == 0 || !tie(i, i - 1)) && i + 1 < n && tie(i + 1, i)) {
                    int e = i + 1; bool inexact = false;
                    while (e < n && tie(e, e - 1)) { inexact |= (pt[e].x != pt[e - 1].x); ++e; }
                    if (inexact) {
                        if (!tie(e - 1, i)) atomicOr(&ctl[0], GTA_B200_ST_NEAR_TIE_X);
                        for (int a = i + 1; a < e; ++a) {
                            const Pt pa = pt[a]; const u16 oa = perm[a];
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
                const double dx = pt[i].x - pt[i - 1].x, dy = pt[i].y - pt[i - 1].y;
                if (dx < 1e-12 && dx > -1e-12 && dy < 1e-12 && dy > -1e-12) atomicOr(&ctl[0], GTA_B200_ST_DUPLICATE);
            }
            __syncthreads();
        }
        PP_TICK(7);
        // ------------------------------------------------------------ the frame's records
        const size_t slot = (size_t)wi;
        if (ctl[0] == 0) {
            LPt* wpt = p.ws.pt + slot * p.ws.cap;
            for (int i = tid; i < n; i += T) {
                const int o = perm[i];
                LPt r; r.x = pt[i].x; r.y = pt[i].y; r.z = pz(o);
                r.head = (u16)GT_NIL; r.perm = (u16)o; r.pad = 0;
                wpt[i] = r;
            }
        }
        if (tid == 0) {
            p.ws.n[slot] = n; p.ws.st[slot] = (int)ctl[0]; p.ws.err[slot] = 0;
            if (p.areabox) p.areabox[fr] = bx * by;
            if (p.ncorr) p.ncorr[fr] = ncorr_out;
        }
        __syncthreads();                                                   // the tile, the scratch and ctl are free for the next frame
        PP_TICK(8);
#ifdef GT_PHASE_TIMING
        if (tid == 0 && p.dbg) atomicAdd(&p.dbg[9], 1ull);
#endif
    }
}

}  // namespace gt
