// gta_corr.cuh -- K0: the -corr box-edge points of one frame, by one CTA (shared by the pre-pass and the star kernel).
//
// Reference src/gta_tri.c:112-272: the box is cut into nx x ny intervals of width espace; for every interval of every box
// edge the nearest atom (lowest y / highest y per x interval, lowest x / highest x per y interval; FLT_MAX / FLT_MIN start
// values, strict comparisons so the first atom wins ties, index 0 for an empty interval) lends its z, interpolated
// between the two opposite edges; the four corners take the mean z of the atoms nearest to / farthest from two of them.
// Same arithmetic and the reference's emission order (appended points bit-identical).
#pragma once
#include "gta_kernel.cuh"

namespace gt {

// tile: the frame's coordinates [natoms][in_dim]; ckey: scratch of >= 12 * (ncorr + 4) bytes; red: >= 6 * (T / 32) doubles;
// ctl0: the frame's status word; corr3: [cap - natoms][3] out. Returns the frame's point count n (natoms when the frame is
// refused: the status word says why). Every thread of the block must call it (it has barriers).
template <int T>
__device__ __forceinline__ int corr_points(const double* tile, int in_dim, int natoms, int cap, double bx, double by, double espace,
                                           unsigned long long* ckey, double* red, uint32_t* ctl0, double* corr3,
                                           double* corr_out, int corr_cap, int fr, int* ncorr_out) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t* ctl = ctl0;
    auto ax = [&](int a) -> double { return tile[(size_t)a * in_dim]; };
    auto ay = [&](int a) -> double { return tile[(size_t)a * in_dim + 1]; };
    auto az = [&](int a) -> double { return in_dim == 3 ? tile[(size_t)a * 3 + 2] : 0.0; };
    int n = natoms;
    *ncorr_out = 0;
    const int nx = (int)(bx / espace), ny = (int)(by / espace);      // gta_tri.c:112-113
    const bool boxok = isfinite(bx) && isfinite(by) && nx >= 0 && ny >= 0 && natoms >= 1;
    const bool toomany = boxok && (nx > (1 << 20) || ny > (1 << 20));
    const int ncorr = (boxok && !toomany) ? 4 + 2 * nx + 2 * ny : 0;
    *ncorr_out = ncorr;
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
                X = j * espace + espace / 2; Y = ((s - 4) & 1) ? by : 0.0;
            } else {
                const int j = (s - 4 - 2 * nx) >> 1;
                const int imin = cidx[oxmin + j], imax = cidx[oxmax + j];
                const double dist1 = ax(imin), dist2 = bx - ax(imax), dist = dist1 + dist2;
                const double zmin = az(imin), zmax = az(imax);
                Z = zmin - (dist1 / dist) * zmin + zmax - (dist2 / dist) * zmax;
                Y = j * espace + espace / 2; X = ((s - 4 - 2 * nx) & 1) ? bx : 0.0;
            }
            corr3[3 * s] = X; corr3[3 * s + 1] = Y; corr3[3 * s + 2] = Z;
            if (corr_out && s < corr_cap) {
                double* o = corr_out + ((size_t)fr * corr_cap + s) * 3;
                o[0] = X; o[1] = Y; o[2] = Z;
            }
        }
        n = natoms + ncorr;
    }
    return n;
}

}  // namespace gt
