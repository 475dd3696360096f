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
    if (bulk && tid == 0 && first < p.nframes)
        pp_bulk_load(psm + L.tile, p.xyz + (size_t)first * natoms * in_dim, frame_bytes, &mbar[0]);

    int it = 0;
    for (int fr = first; fr < p.nframes; fr += stride, ++it) {
        const int buf = it & 1;
        const double* tile = reinterpret_cast<const double*>(psm + L.tile + (size_t)buf * tile_bytes);
        // ---- the next frame's tile starts loading now (its buffer was last read two frames ago, before a barrier)
        const int nxt = fr + stride;
        if (bulk && tid == 0 && nxt < p.nframes) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            pp_bulk_load(psm + L.tile + (size_t)(buf ^ 1) * tile_bytes, p.xyz + (size_t)nxt * natoms * in_dim, frame_bytes, &mbar[buf ^ 1]);
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

        // ------------------------------------------------------------ K0: -corr points (gta_tri.c:112-272; as gta_kernel.cuh)
        if (corr) {
            const int nx = (int)(bx / p.espace), ny = (int)(by / p.espace);      // gta_tri.c:112-113
            const bool boxok = isfinite(bx) && isfinite(by) && nx >= 0 && ny >= 0 && natoms >= 1;
            const bool toomany = boxok && (nx > (1 << 20) || ny > (1 << 20));
            const int ncorr = (boxok && !toomany) ? 4 + 2 * nx + 2 * ny : 0;
            ncorr_out = ncorr;
            const int K = ncorr + 4;                                             // 4 corners + 2(nx+1) + 2(ny+1) slots
            if (!boxok) { if (tid == 0) atomicOr(&ctl[0], natoms < 1 ? GTA_B200_ST_TOO_FEW_POINTS : GTA_B200_ST_NONFINITE); }
            else if (toomany || natoms + ncorr > cap) { if (tid == 0) atomicOr(&ctl[0], GTA_B200_ST_CAPACITY); }
            else {
                const int oymin = 4, oymax = 4 + (nx + 1), oxmin = 4 + 2 * (nx + 1), oxmax = oxmin + (ny + 1);
                const unsigned long long kmin0 = okey((double)FLT_MAX), kmax0 = okey((double)FLT_MIN);   // gta_tri.c:119-120,137-144
                // Per-interval extremes on order-preserving 64-bit keys, found with NATIVE 32-bit shared atomics in two steps (high
                // word, then low word among the atoms that tie on the high word): 64-bit shared atomicMin / atomicMax are
                // compare-and-swap loops and were a quarter of the pre-pass. khi / klo / cidx: K words each.
                uint32_t* khi = reinterpret_cast<uint32_t*>(ckey);
                uint32_t* klo = khi + K;
                int* cidx = reinterpret_cast<int*>(klo + K);
                auto is_min = [&](int s) -> bool { return (s == 0 || s == 2) || (s >= oymin && s < oymax) || (s >= oxmin && s < oxmax); };
                for (int s = tid; s < K; s += T) { khi[s] = (uint32_t)((is_min(s) ? kmin0 : kmax0) >> 32); cidx[s] = 0x7fffffff; }
                __syncthreads();
                // pass 1: extreme values per slot (gta_tri.c:153-202). The four corner slots see every atom: they are reduced in
                // registers and by shuffles; the per-interval slots (a few dozen atoms each) use the atomics.
                const int NONE = 0x7fffffff;
                double bv[4] = {(double)FLT_MAX, (double)FLT_MIN, (double)FLT_MAX, (double)FLT_MIN};      // running extremes, initial values of the reference
                int bi[4] = {NONE, NONE, NONE, NONE};
                for (int j = tid; j < natoms; j += T) {                       // ascending j: the strict comparisons keep the first atom
                    const double X = ax(j), Y = ay(j);
                    const double d0 = X * X + Y * Y, dY = by - Y, d1 = X * X + dY * dY;
                    const double fx = (X / bx) * nx, fy = (Y / by) * ny;
                    if (!(fx > -1.0 && fx < (double)(nx + 1) && fy > -1.0 && fy < (double)(ny + 1))) { atomicOr(&ctl[0], GTA_B200_ST_OUT_OF_BOX); continue; }
                    const int xi = (int)fx, yi = (int)fy;
                    if (d0 < bv[0]) { bv[0] = d0; bi[0] = j; }
                    if (d0 > bv[1]) { bv[1] = d0; bi[1] = j; }
                    if (d1 < bv[2]) { bv[2] = d1; bi[2] = j; }
                    if (d1 > bv[3]) { bv[3] = d1; bi[3] = j; }
                    const uint32_t yh = (uint32_t)(okey(Y) >> 32), xh = (uint32_t)(okey(X) >> 32);
                    atomicMin(&khi[oymin + xi], yh); atomicMax(&khi[oymax + xi], yh);
                    atomicMin(&khi[oxmin + yi], xh); atomicMax(&khi[oxmax + yi], xh);
                }
                {
                    // is candidate b better than a for corner slot c? better value wins (c odd: a maximum), equal values keep the
                    // lower index, "no candidate" loses -- the same atom the reference's sequential scan ends with
                    auto better = [NONE](int c, double va, int ia, double vb, int ib) -> bool {
                        if (ib == NONE) return false;
                        if (ia == NONE) return true;
                        if (vb != va) return (c & 1) ? vb > va : vb < va;
                        return ib < ia;
                    };
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) {
                            const double vo = __shfl_xor_sync(0xffffffffu, bv[c], o); const int io = __shfl_xor_sync(0xffffffffu, bi[c], o);
                            if (better(c, bv[c], bi[c], vo, io)) { bv[c] = vo; bi[c] = io; }
                        }
                    }
                    double* rv = red;                                         // (reduction scratch: 4 x 8 doubles + 4 x 8 ints)
                    int* ri = reinterpret_cast<int*>(rv + 4 * (T / 32));
                    if (lane == 0) {
#pragma unroll
                        for (int c = 0; c < 4; ++c) { rv[4 * warp + c] = bv[c]; ri[4 * warp + c] = bi[c]; }
                    }
                    __syncthreads();
                    if (tid < 4) {
                        double v = rv[tid]; int ix = ri[tid];
                        for (int w = 1; w < T / 32; ++w) if (better(tid, v, ix, rv[4 * w + tid], ri[4 * w + tid])) { v = rv[4 * w + tid]; ix = ri[4 * w + tid]; }
                        cidx[tid] = ix;
                    }
                    // low words start from the initial value's when the high word still is the initial value's, else from the neutral element
                    for (int s = 4 + tid; s < K; s += T) {
                        const unsigned long long k0 = is_min(s) ? kmin0 : kmax0;
                        klo[s] = khi[s] == (uint32_t)(k0 >> 32) ? (uint32_t)k0 : (is_min(s) ? 0xffffffffu : 0u);
                    }
                }
                __syncthreads();
                const bool inbox = !(ctl[0] & GTA_B200_ST_OUT_OF_BOX);
                // pass 1b: low words, among the atoms whose high word is the slot's extreme
                if (inbox) {
                    for (int j = tid; j < natoms; j += T) {
                        const double X = ax(j), Y = ay(j);
                        const int xi = (int)((X / bx) * nx), yi = (int)((Y / by) * ny);
                        const unsigned long long ky = okey(Y), kx = okey(X);
                        const uint32_t yh = (uint32_t)(ky >> 32), xh = (uint32_t)(kx >> 32);
                        if (yh == khi[oymin + xi]) atomicMin(&klo[oymin + xi], (uint32_t)ky);
                        if (yh == khi[oymax + xi]) atomicMax(&klo[oymax + xi], (uint32_t)ky);
                        if (xh == khi[oxmin + yi]) atomicMin(&klo[oxmin + yi], (uint32_t)kx);
                        if (xh == khi[oxmax + yi]) atomicMax(&klo[oxmax + yi], (uint32_t)kx);
                    }
                }
                __syncthreads();
                // pass 2: lowest atom index attaining the extreme of its intervals (strict < / > keeps the first one)
                if (inbox) {
                    for (int j = tid; j < natoms; j += T) {
                        const double X = ax(j), Y = ay(j);
                        const int xi = (int)((X / bx) * nx), yi = (int)((Y / by) * ny);
                        const double fmx = (double)FLT_MAX, fmn = (double)FLT_MIN;
                        const unsigned long long ky = okey(Y), kx = okey(X);
                        auto slot_key = [&](int s) -> unsigned long long { return ((unsigned long long)khi[s] << 32) | klo[s]; };
                        if (Y < fmx && ky == slot_key(oymin + xi)) atomicMin(&cidx[oymin + xi], j);
                        if (Y > fmn && ky == slot_key(oymax + xi)) atomicMin(&cidx[oymax + xi], j);
                        if (X < fmx && kx == slot_key(oxmin + yi)) atomicMin(&cidx[oxmin + yi], j);
                        if (X > fmn && kx == slot_key(oxmax + yi)) atomicMin(&cidx[oxmax + yi], j);
                    }
                }
                __syncthreads();
                for (int s = tid; s < K; s += T) if (cidx[s] == 0x7fffffff) cidx[s] = 0;      // index arrays start at 0 (gta_tri.c:146-149)
                __syncthreads();
                // corner z (gta_tri.c:209-212) and the appended points in the reference's order (:220-272)
                const double zc = (az(cidx[0]) + az(cidx[1]) + az(cidx[2]) + az(cidx[3])) / 4.0;
                for (int s = tid; s < ncorr; s += T) {
                    double X, Y, Z;
                    if (s < 4) {
                        X = (s == 1 || s == 2) ? bx : 0.0; Y = (s >= 2) ? by : 0.0; Z = zc;
                    } else if (s < 4 + 2 * nx) {
                        const int j = (s - 4) >> 1;
                        const int imin = cidx[oymin + j], imax = cidx[oymax + j];
                        const double dist1 = ay(imin), dist2 = by - ay(imax), dist = dist1 + dist2;
                        const double zmin = az(imin), zmax = az(imax);
                        Z = zmin - (dist1 / dist) * zmin + zmax - (dist2 / dist) * zmax;
                        X = j * p.espace + p.espace / 2; Y = ((s - 4) & 1) ? by : 0.0;
                    } else {
                        const int j = (s - 4 - 2 * nx) >> 1;
                        const int imin = cidx[oxmin + j], imax = cidx[oxmax + j];
                        const double dist1 = ax(imin), dist2 = bx - ax(imax), dist = dist1 + dist2;
                        const double zmin = az(imin), zmax = az(imax);
                        Z = zmin - (dist1 / dist) * zmin + zmax - (dist2 / dist) * zmax;
                        Y = j * p.espace + p.espace / 2; X = ((s - 4 - 2 * nx) & 1) ? bx : 0.0;
                    }
                    corr3[3 * s] = X; corr3[3 * s + 1] = Y; corr3[3 * s + 2] = Z;
                    if (p.corr && s < p.corr_cap) {
                        double* o = p.corr + ((size_t)fr * p.corr_cap + s) * 3;
                        o[0] = X; o[1] = Y; o[2] = Z;
                    }
                }
                n = natoms + ncorr;
            }
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
        const size_t slot = (size_t)fr;
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
