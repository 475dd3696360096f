// gta_lpf.cuh -- lane-per-frame execution of the triangulation (see lpf_core.cuh) and its area epilogue.
//
// Pipeline for a large batch (>= tens of thousands of frames; smaller batches stay on the CTA-per-frame
// kernel, whose per-frame latency is 20x lower):
//   1. tessellate_kernel in pre-pass mode: load, -corr, sort -> workspace          (gta_kernel.cuh)
//   2. lpf_wave_kernel: every frame is one slot of a phase-sorted scheduler; a lane runs one iteration of the D&C
//      state machine (lpf_core.cuh) for one frame at a time                              (this file)
//   3. lpf_area_kernel: one warp per frame enumerates the triangles and sums areas  (this file)
// Replaces the same reference loops as the fused kernel (src/gta_tri.c:106-296).
#pragma once
#include "gta_kernel.cuh"

namespace gt {

struct LpfParams {
    LpfWs ws;
    int nframes;
    unsigned int* counter;                 // next frame to hand out
    unsigned long long* dbg;               // -DGT_PHASE_TIMING builds: [3*ph] cycles, [3*ph+1] chunks, [3*ph+2] lanes; [24] barrier wait, [25] rounds x warps, [26..30] rounds of block 0 by chunk count
};

__device__ __forceinline__ LFrame lpf_frame(const LpfWs& ws, int fr) {
    LFrame f;
    const size_t s = (size_t)fr;
    f.pt = ws.pt + s * ws.cap;
    f.rec = ws.rec + s * 2 * (size_t)ws.cap_e;
    f.n = ws.n[s];
    f.cap_e = ws.cap_e;
    return f;
}

// One iteration of phase PH for one slot, out of line so that every phase gets a register allocation of its own
// (inlined into one kernel, the 14 instances share a 128-register budget and spill ~1 KB per thread). The state
// is unpacked straight from shared memory -- loads of words the phase never reads are dead code -- and only the
// words the phase may modify (LsDirty<PH>) are written back. Returns the packed phase word's new phase.
// LPF_MAX_SLOTS = slot capacity of a block = stride of the word-major state array (a constant: the state accesses of a
// step become base + immediate offset). Deliberately far below what shared memory could hold:
//  - what is not shared memory is L1, and the steps' loads want it (704 slots in use: 180 k frames/s when the block
//    needs the 228 KB carve-out and leaves 28 KB of L1, 257 k with 196 KB, 281 k with 164 KB or less);
//  - a round serves 16 warps; more than ~448 slots per block only queue (measured when 14 warps ran it: 66 k frames per
//    launch 300 k frames/s, 100 k in one launch 281 k) -- the host splits larger batches into even launches instead
//    (gta_b200.cu).
constexpr int LPF_MAX_SLOTS = 576;

template <int PH, bool AP>
__device__ __noinline__ int lpf_wave_phase(uint32_t* st, int sl, LPt* pt, LRec* rec, int n, int cap_e, uint32_t* err_out) {
    LFrame f; f.pt = pt; f.rec = rec; f.n = n; f.cap_e = cap_e;
    LState s;
    uint32_t* mine = st + sl;
    ls_unpack(s, [&](int w) { return mine[w * LPF_MAX_SLOTS]; });
    lpf_step_t<PH, AP>(f, s);
    ls_writeback<PH>(s, [&](int w, uint32_t v) { mine[w * LPF_MAX_SLOTS] = v; });
    if (s.phase == LP_DONE) *err_out = s.err;
    return s.phase;
}

// Phase-sorted ("wavefront") scheduler. A block keeps SLOTS frames in flight; their packed states live in
// shared memory (word-major: st[w * SLOTS + slot]) and every slot is filed in the bin of its phase. Every round
// the warps take tickets for chunks of up to 32 slots that are in the SAME phase, run that phase's straight-line
// step for them and file each slot under its new phase for the next round (double-buffered bins): no divergence
// between lanes, 64 dependent-load chains in flight per warp. A slot whose frame is finished goes to the
// "needs a frame" bin and takes the next frame of the batch from a global counter.
template <int THREADS, bool AP>
__global__ void __launch_bounds__(THREADS, 1) lpf_wave_kernel(const LpfParams p, const int SLOTS) {
    extern __shared__ __align__(16) unsigned char lsm[];
    constexpr int NB = LP_DONE + 1;                                        // one bin per phase + one for slots that need a frame
    uint32_t* st = reinterpret_cast<uint32_t*>(lsm);                       // [LS_WORDS][LPF_MAX_SLOTS]
    int* slot_fr = reinterpret_cast<int*>(st + (size_t)LS_WORDS * LPF_MAX_SLOTS);  // [SLOTS] frame of the slot
    u16* slot_n = reinterpret_cast<u16*>(slot_fr + SLOTS);                 // [SLOTS] its point count
    u16* bins = slot_n + SLOTS;                                            // [2][NB][SLOTS] slots of each phase, this round / next round
    int* cnt = reinterpret_cast<int*>(bins + 2 * (size_t)NB * SLOTS);      // [3][16] slots per bin: this round / next round / being cleared
    int* ctl = cnt + 48;                                                   // [0] batch exhausted, [1..3] chunk tickets (same rotation)
    const int tid = threadIdx.x, lane = tid & 31;

    if (tid < 48) cnt[tid] = 0;
    if (tid < 4) ctl[tid] = 0;
    __syncthreads();
    // every slot starts in the "needs a frame" bin
    for (int sl = tid; sl < SLOTS; sl += THREADS) bins[(size_t)LP_DONE * SLOTS + sl] = (u16)sl;
    if (tid == 0) cnt[LP_DONE] = SLOTS;
    __syncthreads();

    for (int round = 0;; ++round) {
        const int b = round & 1, r3 = round % 3, r3n = (round + 1) % 3, r3c = (round + 2) % 3;
        const int* c_now = cnt + 16 * r3;
        int* c_next = cnt + 16 * r3n;
        // the third set of counters / tickets was last read a round ago: clear it for the round after this one
        if (tid < 16) cnt[16 * r3c + tid] = 0;
        if (tid == 0) ctl[1 + r3c] = 0;
        const u16* bin_now = bins + (size_t)b * NB * SLOTS;
        u16* bin_next = bins + (size_t)(b ^ 1) * NB * SLOTS;
        // ---- chunk table, computed by every warp for itself: lane q holds bin q's count and first chunk
        const bool exhausted = ctl[0] != 0;
        // (position q of the table is bin ORDER[q]: the phases with the longest steps first, so that the round's last
        // chunks are its shortest and the warps reach the barrier closer together)
        constexpr unsigned ORDER = (unsigned)LP_INS | (unsigned)LP_SCAN << 4 | (unsigned)LP_C << 8 | (unsigned)LP_TAN << 12 | (unsigned)LP_V << 16
                                 | (unsigned)LP_LEAF << 20 | (unsigned)LP_TINIT << 24 | (unsigned)LP_DONE << 28;
        const int qbin = (int)((ORDER >> (4 * (lane & 7))) & 15u);
        int myc = lane < NB ? c_now[qbin] : 0;
        if (qbin == LP_DONE && exhausted) myc = 0;                        // nothing left to fetch: empty slots stay empty
        int incl = (myc + 31) >> 5;
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }   // (lanes >= 8 hold 0)
        const int first_chunk = incl - ((myc + 31) >> 5);
        const int nchunks = __shfl_sync(0xffffffffu, incl, 7);
        if (nchunks == 0) break;                                          // nothing in flight and nothing left to fetch
#ifdef GT_PHASE_TIMING
        if (p.dbg && tid == 0 && blockIdx.x == 0) atomicAdd(p.dbg + 26 + min(max(nchunks - 16, 0), 4), 1ull);      // rounds with <= 16, 17, 18, 19, >= 20 chunks
#endif
        // ---- one iteration for every live slot: a warp takes the next chunk of <= 32 slots that share a phase and files
        // every slot under its new phase for the next round
        // (a warp's first chunk is the one with its own number; tickets hand out the rest)
        int* ticket = &ctl[1 + r3];
        for (int c = tid >> 5;;) {
            if (c >= nchunks) break;
            // first_chunk is non-decreasing over the bins: the last bin that starts at or before c owns chunk c
            const int q = __popc(__ballot_sync(0xffffffffu, lane < NB && first_chunk <= c)) - 1;      // warp-uniform
            const int bin = (int)((ORDER >> (4 * q)) & 15u);
            const int idx = (c - __shfl_sync(0xffffffffu, first_chunk, q)) * 32 + lane;
            const int nph = __shfl_sync(0xffffffffu, myc, q);
            if (idx < nph) {
                const int sl = bin_now[(size_t)bin * SLOTS + idx];
                int next = LP_DONE;
                if (bin == LP_DONE) {
                    // the slot's frame is finished (or it never had one): take the next frame of the batch that passed the pre-pass
                    for (;;) {
                        const int fr = (int)atomicAdd(p.counter, 1u);
                        if (fr >= p.nframes) { ctl[0] = 1; break; }
                        if (p.ws.st[fr] != 0) continue;
                        LFrame f = lpf_frame(p.ws, fr);
                        if (f.n < 2) continue;
                        LState s = {};
                        lp_begin(f, s);
                        if (s.phase == LP_DONE) continue;
                        ls_pack(s, [&](int w, uint32_t v) { st[w * LPF_MAX_SLOTS + sl] = v; });
                        slot_fr[sl] = fr; slot_n[sl] = (u16)f.n;
                        next = s.phase;
                        break;
                    }
                } else {
                    const int fr = slot_fr[sl];
                    LPt* fpt = p.ws.pt + (size_t)fr * p.ws.cap; LRec* frec = p.ws.rec + (size_t)fr * 2 * (size_t)p.ws.cap_e;
                    const int fn = slot_n[sl], fce = p.ws.cap_e;
                    uint32_t* eo = p.ws.err + fr;
#ifdef GT_PHASE_TIMING
                    const long long tk0 = clock64();
#endif
                    switch (bin) {                                        // every lane of the chunk is in this phase
                    case LP_LEAF: next = lpf_wave_phase<LP_LEAF, AP>(st, sl, fpt, frec, fn, fce, eo); break;
                    case LP_TINIT: next = lpf_wave_phase<LP_TINIT, AP>(st, sl, fpt, frec, fn, fce, eo); break;
                    case LP_TAN: next = lpf_wave_phase<LP_TAN, AP>(st, sl, fpt, frec, fn, fce, eo); break;
                    case LP_INS: next = lpf_wave_phase<LP_INS, AP>(st, sl, fpt, frec, fn, fce, eo); break;
                    case LP_SCAN: next = lpf_wave_phase<LP_SCAN, AP>(st, sl, fpt, frec, fn, fce, eo); break;
                    case LP_V: next = lpf_wave_phase<LP_V, AP>(st, sl, fpt, frec, fn, fce, eo); break;
                    default: next = lpf_wave_phase<LP_C, AP>(st, sl, fpt, frec, fn, fce, eo); break;
                    }
#ifdef GT_PHASE_TIMING
                    if (p.dbg) {
                        const long long tk1 = clock64(); const unsigned am = __activemask();
                        if (lane == __ffs(am) - 1) { atomicAdd(p.dbg + 3 * bin, (unsigned long long)(tk1 - tk0)); atomicAdd(p.dbg + 3 * bin + 1, 1ull); atomicAdd(p.dbg + 3 * bin + 2, (unsigned long long)__popc(am)); }
                    }
#endif
                }
                bin_next[(size_t)next * SLOTS + atomicAdd(&c_next[next], 1)] = (u16)sl;
            }
            if (nchunks <= THREADS / 32) break;                            // every chunk had a warp of its own
            if (lane == 0) c = atomicAdd(ticket, 1) + THREADS / 32;
            c = __shfl_sync(0xffffffffu, c, 0);
        }
#ifdef GT_PHASE_TIMING
        const long long tb0 = clock64();
        __syncthreads();
        if (p.dbg && lane == 0) { atomicAdd(p.dbg + 24, (unsigned long long)(clock64() - tb0)); atomicAdd(p.dbg + 25, 1ull); }
#else
        __syncthreads();
#endif
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
                const LPt pi = f.pt[i]; const double zi = pi.z;
                int* out = p.tri ? p.tri + ((size_t)fr * p.tri_cap) * 3 : nullptr;
                lf_vertex_triangles(f, i, [&](int n1, int n2) {
                    const LPt p1 = f.pt[n1], p2 = f.pt[n2];
                    a3 += tri_area3(pi.x, pi.y, zi, p1.x, p1.y, p1.z, p2.x, p2.y, p2.z);
                    if (want2d) a2 += tri_area3(pi.x, pi.y, 0.0, p1.x, p1.y, 0.0, p2.x, p2.y, 0.0);
                    if (out && (int)offs < p.tri_cap) { out[3 * offs] = pi.perm; out[3 * offs + 1] = p1.perm; out[3 * offs + 2] = p2.perm; }
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
