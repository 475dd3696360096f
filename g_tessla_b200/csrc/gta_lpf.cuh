// gta_lpf.cuh -- the throughput path: walkers over (frame, subtree) slots, phase-sorted.
//
// Pipeline for a batch:
//   1. tessellate_kernel in pre-pass mode: load, -corr, sort -> workspace                      (gta_kernel.cuh)
//   2. lpf_wave_kernel: one persistent block per SM; every live walker is one slot of a phase-sorted scheduler, a lane
//      runs one iteration of the D&C state machine (lpf_core.cuh) for one walker at a time         (this file)
//   3. lpf_area_kernel: one warp per frame enumerates the triangles and sums areas               (this file)
// A frame is not one lane for its whole life: its recursion tree is cut at depth `split` into independent subtrees that
// start in parallel and hand over to their parents pairwise (lp_finish_node), so a frame is in flight for ~2,800 rounds
// instead of ~28,000, ten times fewer frames are live at any time, and a batch of a few thousand frames already fills
// the GPU.
// Replaces the same reference loops as the fused kernel (src/gta_tri.c:106-296).
#pragma once
#include "gta_kernel.cuh"

namespace gt {

struct LpfParams {
    LpfWs ws;
    int nframes;
    const unsigned int* nframes_dev;       // != nullptr: the number of frames in the workspace is read from the device (fallback of the star path)
    unsigned int* counter;                 // next frame of the batch to hand out
    int split;                             // a frame starts as 2^split subtree walkers (fewer if it has too few points)
    int fslots;                            // frames a block can have in flight
    int defer_below;                       // a bin of the phases in defer_phases with fewer slots than this waits while (round & defer_rounds) != 0 (0: never)
    int defer_phases, defer_rounds;
    unsigned long long* dbg;               // -DGT_PHASE_TIMING builds: [3*ph] cycles, [3*ph+1] chunks, [3*ph+2] lanes; [27] live walkers summed over rounds,
                                           // [28] rounds, [30] barrier wait, [31] rounds x warps
};

__device__ __forceinline__ LFrame lpf_frame(const LpfWs& ws, int fr) {
    LFrame f;
    const size_t s = (size_t)fr;
    f.pt = ws.pt + s * ws.cap;
    f.rec = ws.rec + s * 2 * (size_t)ws.cap_e;
    f.m = nullptr;
    f.n = ws.n[s];
    f.cap_e = ws.cap_e;
    return f;
}

// LPF_MAX_SLOTS = slot capacity of a block = stride of the word-major state array (a constant: the state accesses of a
// step become base + immediate offset). A round gives every warp one chunk of <= 32 same-phase slots; with 16 warps and
// ~25 slots per chunk more than ~352 slots only add second chunks to a round that already lasts as long as its slowest
// one (measured on cfg 3, 100k frames: 320 / 352 / 384 / 416 / 480 / 576 slots -> 286 / 277 / 276 / 299 / 290 / 294 ms).
#ifndef GT_LPF_STRIDE
#define GT_LPF_STRIDE 384
#endif
constexpr int LPF_MAX_SLOTS = GT_LPF_STRIDE;
constexpr int LPF_DEFAULT_SLOTS = 384;
constexpr int LPF_MAX_SPLIT = 5;           // the "arrived" bits of a frame's top nodes fit one word (lpf_core.cuh LMeta), and a
                                           // frame's walkers start from one chunk of free slots (<= 32)

// shared-memory layout of a block, computed identically on the host
struct LpfSmem {
    size_t st, meta, fs_frame, fs_n, fs_mask, slot_fs, freeq, bins, cnt, ctl, total;
    __host__ __device__ LpfSmem(int slots, int fslots) {
        constexpr int NB = LP_DONE + 1;
        size_t o = 0;
        st = o; o += (size_t)LS_WORDS * LPF_MAX_SLOTS * 4;
        meta = o; o += sizeof(LMeta) * (size_t)fslots;
        fs_frame = o; o += 4 * (size_t)fslots;
        fs_n = o; o += 4 * (size_t)fslots;
        fs_mask = o; o += 4 * (size_t)((fslots + 31) / 32);
        cnt = o; o += 48 * 4;
        ctl = o; o += 8 * 4;
        slot_fs = o; o += 2 * (size_t)slots;
        freeq = o; o += 2 * (size_t)slots;
        bins = o; o += 2 * (size_t)2 * (NB - 1) * slots;
        total = (o + 15) & ~(size_t)15;
    }
};

// One iteration of phase PH for one slot, out of line so that every phase gets a register allocation of its own
// (inlined into one kernel, the instances share a 128-register budget and spill ~1 KB per thread). The state
// is unpacked straight from shared memory -- loads of words the phase never reads are dead code -- and only the
// words the phase may modify (LsDirty<PH>) are written back. Returns the walker's new phase.
template <int PH, bool AP>
__device__ __noinline__ int lpf_wave_phase(uint32_t* st, int sl, LPt* pt, LRec* rec, LMeta* m, int n, int cap_e) {
    LFrame f; f.pt = pt; f.rec = rec; f.m = m; f.n = n; f.cap_e = cap_e;
    LState s;
    uint32_t* mine = st + sl;
    ls_unpack(s, [&](int w) { return mine[w * LPF_MAX_SLOTS]; });
    lpf_step_t<PH, AP>(f, s);
    ls_writeback<PH>(s, [&](int w, uint32_t v) { mine[w * LPF_MAX_SLOTS] = v; });
    // errors (never on contract input): the scheduler records them for the frame and gives the subtree up after a hard one
    return s.err ? (s.phase | 16) : s.phase;
}
// A walker came back with error bits (they are in its packed state): record them for the frame and give the subtree up
// -- the frame still completes, its merges above are skipped. (A range error is hard too: its predicate's sign was forced
// to 0, and a ring scan over inconsistent signs need not end.)
// One out-of-line copy for all phases.
__device__ __noinline__ int lpf_wave_error(uint32_t* st, int sl, LMeta* m, int n, int cap_e) {
    LFrame f; f.pt = nullptr; f.rec = nullptr; f.m = m; f.n = n; f.cap_e = cap_e;
    LState s;
    uint32_t* mine = st + sl;
    ls_unpack(s, [&](int w) { return mine[w * LPF_MAX_SLOTS]; });
    if (s.phase >= LP_DONE || !(s.err & (GT_ERR_POOL | GT_ERR_RING | GT_ERR_RANGE))) { atomicOr(&m->err, s.err); return s.phase; }
    lp_abort(f, s);
    ls_pack(s, [&](int w, uint32_t v) { mine[w * LPF_MAX_SLOTS] = v; });
    return s.phase;
}

// frame slots of a block (what a frame's walkers share lives there): a bit mask of free ones
__device__ __forceinline__ int lpf_fs_pop(uint32_t* mask, int words) {
    for (int w = 0; w < words; ++w) {
        uint32_t m = *(volatile uint32_t*)&mask[w];
        while (m) {
            const uint32_t bit = m & (0u - m);
            const uint32_t old = atomicAnd(&mask[w], ~bit);
            if (old & bit) return w * 32 + __ffs(bit) - 1;
            m = old & ~bit;
        }
    }
    return -1;
}

// A chunk of free slots admits `nadm` frames: lane a < nadm takes a frame slot and the next frame of the batch that passed
// the pre-pass, then lanes a * 2^K ... a * 2^K + 2^kk - 1 start the walkers of its subtrees. Returns the lane's new phase.
__device__ __noinline__ int lpf_admit(const LpfParams& p, uint32_t* st, LMeta* meta, int* fs_frame, int* fs_n, uint32_t* fs_mask, int FSW,
                                      u16* slot_fs, int* exhausted_next, int nadm, bool live, int sl, int lane) {
    const int K = p.split;
    int fs = -1, kk = 0;
    if (lane < nadm) {
        fs = lpf_fs_pop(fs_mask, FSW);
        if (fs >= 0) {
            for (;;) {
                const int fr = (int)atomicAdd(p.counter, 1u);
                if (fr >= (p.nframes_dev ? (int)*p.nframes_dev : p.nframes)) { *exhausted_next = 1; atomicOr(&fs_mask[fs >> 5], 1u << (fs & 31)); fs = -1; break; }
                const int n = p.ws.n[fr];
                if (p.ws.st[fr] != 0 || n < 2) continue;          // rejected by the pre-pass: the area kernel reports it
                fs_frame[fs] = fr; fs_n[fs] = n;
                meta[fs].bump = 0; meta[fs].err = 0; meta[fs].ready = 0;
                kk = lp_split_level(n, K);
                break;
            }
        }
    }
    __syncwarp();
    const int a = lane >> K, k = lane & ((1 << K) - 1);
    const int fs_a = __shfl_sync(0xffffffffu, fs, a), kk_a = __shfl_sync(0xffffffffu, kk, a);
    if (!(live && a < nadm && fs_a >= 0 && k < (1 << kk_a))) return LP_DONE;      // (kk < K: a frame with too few points for 2^K subtrees)
    LFrame f; f.pt = nullptr; f.rec = nullptr; f.m = meta + fs_a; f.n = fs_n[fs_a]; f.cap_e = p.ws.cap_e;
    LState s = {};
    lp_begin(f, s, kk_a, k);
    ls_pack(s, [&](int w, uint32_t v) { st[w * LPF_MAX_SLOTS + sl] = v; });
    slot_fs[sl] = (u16)fs_a;
    return s.phase;
}

// Phase-sorted ("wavefront") scheduler. A block keeps up to SLOTS walkers in flight; their packed states live in
// shared memory (word-major: st[w * LPF_MAX_SLOTS + slot]) and every slot is filed in the bin of its phase. Every round
// the warps take tickets for chunks of up to 32 slots that are in the SAME phase, run that phase's straight-line step
// for them and file each slot under its new phase for the next round (double-buffered bins, one barrier per round): no
// divergence between lanes, 64 dependent-load chains in flight per warp. A retired walker's slot goes to the "free"
// bin; a chunk of free slots admits as many whole frames as it can start all subtree walkers of.
template <int THREADS, bool AP>
__global__ void __launch_bounds__(THREADS, 1) lpf_wave_kernel(const LpfParams p, const int SLOTS) {
    extern __shared__ __align__(16) unsigned char lsm[];
    constexpr int NB = LP_DONE + 1;                                        // one bin per phase + one for free slots
    const int FS = p.fslots, FSW = (FS + 31) / 32;
    const LpfSmem L(SLOTS, FS);
    uint32_t* st = reinterpret_cast<uint32_t*>(lsm + L.st);                // [LS_WORDS][LPF_MAX_SLOTS]
    LMeta* meta = reinterpret_cast<LMeta*>(lsm + L.meta);                  // [FS] what a frame's walkers share
    int* fs_frame = reinterpret_cast<int*>(lsm + L.fs_frame);              // [FS] frame of the batch
    int* fs_n = reinterpret_cast<int*>(lsm + L.fs_n);                      // [FS] its point count
    uint32_t* fs_mask = reinterpret_cast<uint32_t*>(lsm + L.fs_mask);      // free frame slots
    u16* slot_fs = reinterpret_cast<u16*>(lsm + L.slot_fs);                // [SLOTS] frame slot of the walker
    u16* freeq = reinterpret_cast<u16*>(lsm + L.freeq);                    // [SLOTS] ring of free slots (retired walkers), see below
    u16* bins = reinterpret_cast<u16*>(lsm + L.bins);                      // [2][NB - 1][SLOTS] slots of each phase, this round / next round
    int* cnt = reinterpret_cast<int*>(lsm + L.cnt);                        // [3][16] slots per bin ([15]: batch exhausted): this round / next round / being cleared
    int* ctl = reinterpret_cast<int*>(lsm + L.ctl);                        // [1..3] chunk tickets (same rotation as cnt)
    const int tid = threadIdx.x, lane = tid & 31;
    const int K = p.split;

    if (tid < 48) cnt[tid] = 0;
    if (tid < 8) ctl[tid] = 0;
    for (int w = tid; w < FSW; w += THREADS) fs_mask[w] = (w * 32 + 32 <= FS) ? 0xffffffffu : ((1u << (FS - w * 32)) - 1u);
    __syncthreads();
    // Free slots wait in a ring, not in a bin: a bin costs a chunk (a warp) every round, and with 2^K slots needed per
    // admission there nearly always are a few free ones waiting. Retiring walkers append behind q_tail (their count is
    // this round's c_next[LP_DONE]); the round's one admission chunk takes from q_head. Both cursors are advanced by every
    // warp for itself from values that are stable across the round barrier, so all warps see the same chunk table.
    for (int sl = tid; sl < SLOTS; sl += THREADS) freeq[sl] = (u16)sl;
    int q_head = 0, q_tail = SLOTS;
    bool exhausted = false;                                                // the batch has no frame left to admit
    __syncthreads();

    for (int round = 0;; ++round) {
        const int b = round & 1, r3 = round % 3, r3n = (round + 1) % 3, r3c = (round + 2) % 3;
        const int* c_now = cnt + 16 * r3;
        int* c_next = cnt + 16 * r3n;
        // the third set of counters / tickets was last read a round ago: clear it for the round after this one
        if (tid < 16) cnt[16 * r3c + tid] = 0;
        if (tid == 0) ctl[1 + r3c] = 0;
        const u16* bin_now = bins + (size_t)b * (NB - 1) * SLOTS;
        u16* bin_next = bins + (size_t)(b ^ 1) * (NB - 1) * SLOTS;
        q_tail += c_now[LP_DONE];                                         // walkers that retired in the previous round
        // this round's admission: up to 32 free slots, whole frames (2^K slots each) only
        exhausted |= c_now[15] != 0;                                      // (raised through the round's counters: stable while the round runs)
        const int take = exhausted ? 0 : (min(32, q_tail - q_head) & ~((1 << K) - 1));
        const int q_take0 = q_head;
        q_head += take;
        // ---- chunk table, computed by every warp for itself: lane q holds bin q's count and first chunk
        // (position q of the table is bin ORDER[q]: the phases with the longest steps first, so that the round's last
        // chunks are its shortest and the warps reach the barrier closer together)
        constexpr unsigned ORDER = (unsigned)LP_INS | (unsigned)LP_SCAN << 4 | (unsigned)LP_C << 8 | (unsigned)LP_TAN << 12 | (unsigned)LP_V << 16
                                 | (unsigned)LP_LEAF << 20 | (unsigned)LP_TINIT << 24 | (unsigned)LP_DONE << 28;
        const int qbin = (int)((ORDER >> (4 * (lane & 7))) & 15u);
        int myc = (lane < NB && qbin < NB) ? (qbin == LP_DONE ? take : c_now[qbin]) : 0;
        int incl = (myc + 31) >> 5;
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }   // (lanes >= 8 hold 0)
        const int first_chunk = incl - ((myc + 31) >> 5);
        const int nchunks = __shfl_sync(0xffffffffu, incl, 7);
        if (nchunks == 0) break;                                          // no walker in flight and nothing left to start
#ifdef GT_PHASE_TIMING
        if (p.dbg && tid == 0) { int lv = 0; for (int q = 0; q < LP_DONE; ++q) lv += c_now[q]; atomicAdd(p.dbg + 27, (unsigned long long)lv); atomicAdd(p.dbg + 28, 1ull); }
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
            const bool live = idx < nph;
            const int sl = !live ? 0 : (bin == LP_DONE ? freeq[(q_take0 + idx) % SLOTS] : bin_now[(size_t)bin * SLOTS + idx]);
            int next = LP_DONE;
            if (bin == LP_DONE) {
                // free slots (lanes 0 .. na-1 of the chunk): start as many whole frames as there are slots for all of a
                // frame's 2^K subtree walkers (out of line: it runs for a few chunks per frame, the loop around it all the time)
                next = lpf_admit(p, st, meta, fs_frame, fs_n, fs_mask, FSW, slot_fs, c_next + 15, take >> K, live, sl, lane);
            } else if (((p.defer_phases >> bin) & 1) && nph < p.defer_below && (round & p.defer_rounds)) {
                // a sparse bin (leaves and merge starts: 6-9 slots per round on cfg 3) waits: its slots are re-filed unchanged, the
                // warp is free for another chunk at once, and the chunk this phase gets every fourth round is four times as full.
                // A round lasts as long as its slowest chunk whatever the number of chunks, as long as every chunk has a warp:
                // fewer chunks per round = more live walkers per round (cfg 3, 100k frames: 273 -> 265 ms)
                next = bin;
            } else if (live) {
                const int fs = slot_fs[sl];
                const int fr = fs_frame[fs];
                LPt* fpt = p.ws.pt + (size_t)fr * p.ws.cap; LRec* frec = p.ws.rec + (size_t)fr * 2 * (size_t)p.ws.cap_e;
                LMeta* fm = meta + fs;
                const int fn = fs_n[fs], fce = p.ws.cap_e;
#ifdef GT_PHASE_TIMING
                const long long tk0 = clock64();
#endif
                switch (bin) {                                            // every lane of the chunk is in this phase
                case LP_LEAF: next = lpf_wave_phase<LP_LEAF, AP>(st, sl, fpt, frec, fm, fn, fce); break;
                case LP_TINIT: next = lpf_wave_phase<LP_TINIT, AP>(st, sl, fpt, frec, fm, fn, fce); break;
                case LP_TAN: next = lpf_wave_phase<LP_TAN, AP>(st, sl, fpt, frec, fm, fn, fce); break;
                case LP_INS: next = lpf_wave_phase<LP_INS, AP>(st, sl, fpt, frec, fm, fn, fce); break;
                case LP_SCAN: next = lpf_wave_phase<LP_SCAN, AP>(st, sl, fpt, frec, fm, fn, fce); break;
                case LP_V: next = lpf_wave_phase<LP_V, AP>(st, sl, fpt, frec, fm, fn, fce); break;
                default: next = lpf_wave_phase<LP_C, AP>(st, sl, fpt, frec, fm, fn, fce); break;
                }
                if (next & 16) next = lpf_wave_error(st, sl, fm, fn, fce);
#ifdef GT_PHASE_TIMING
                if (p.dbg) {
                    const long long tk1 = clock64(); const unsigned am = __activemask();
                    if (lane == __ffs(am) - 1) { atomicAdd(p.dbg + 3 * bin, (unsigned long long)(tk1 - tk0)); atomicAdd(p.dbg + 3 * bin + 1, 1ull); atomicAdd(p.dbg + 3 * bin + 2, (unsigned long long)__popc(am)); }
                }
#endif
                if (next == LP_ROOT) {
                    // the frame's triangulation is complete (every other walker of it has retired): publish its error bits
                    // for the area kernel and free the frame slot
                    p.ws.err[fr] = *(volatile uint32_t*)&fm->err;
                    __threadfence_block();
                    atomicOr(&fs_mask[fs >> 5], 1u << (fs & 31));
                    next = LP_DONE;
                }
            }
            if (live) {
                const int pos = atomicAdd(&c_next[next], 1);
                if (next == LP_DONE) freeq[(q_tail + pos) % SLOTS] = (u16)sl;         // retired (or not started): behind this round's tail
                else bin_next[(size_t)next * SLOTS + pos] = (u16)sl;
            }
            if (nchunks <= THREADS / 32) break;                            // every chunk had a warp of its own
            if (lane == 0) c = atomicAdd(ticket, 1) + THREADS / 32;
            c = __shfl_sync(0xffffffffu, c, 0);
        }
#ifdef GT_PHASE_TIMING
        const long long tb0 = clock64();
        __syncthreads();
        if (p.dbg && lane == 0) { atomicAdd(p.dbg + 30, (unsigned long long)(clock64() - tb0)); atomicAdd(p.dbg + 31, 1ull); }
#else
        __syncthreads();
#endif
    }
}

// One warp per frame: triangles owned by each sorted vertex, area of each (reference area_tri,
// include/gta_tri.h:91-99), fixed-order warp reduction.
struct LpfAreaParams {
    LpfWs ws; int nframes; int natoms; unsigned flags;
    const int* list; const unsigned int* list_count;      // fallback of the star path: workspace slot i = frame list[i] of the batch
    double* area; double* area2d; int* status; int* ntri; int* tri; int tri_cap;
};
__global__ void __launch_bounds__(128) lpf_area_kernel(const LpfAreaParams p) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= (p.list ? (int)*p.list_count : p.nframes)) return;
    const int fr = warp;                                                   // workspace slot
    const int ofr = p.list ? p.list[warp] : warp;                          // frame of the batch the results belong to
    int st = p.ws.st[fr];
    const uint32_t e = p.ws.err[fr];
    if (e & GT_ERR_POOL) st |= GTA_B200_ST_POOL;
    if (e & GT_ERR_RANGE) st |= GTA_B200_ST_PRECISION_RANGE;
    if (e & GT_ERR_RING) st |= GTA_B200_ST_INTERNAL;
    double a3 = 0.0, a2 = 0.0; unsigned total = 0;
    if (st == 0) {
        const LFrame f = lpf_frame(p.ws, fr);
        const bool want2d = (p.flags & GTA_B200_2D) != 0;
        unsigned base = 0, mine = 0;
        for (int i0 = 0; i0 < f.n; i0 += 32) {
            const int i = i0 + lane;
            unsigned offs = 0;
            if (p.tri) {                                                   // (parity tests) positions in the frame's triangle list: a count pass first
                const unsigned cnt = (i < f.n) ? (unsigned)lf_vertex_triangles(f, i, [](int, int) {}) : 0u;
                unsigned inc = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { unsigned t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
                offs = base + inc - cnt;
                base += __shfl_sync(0xffffffffu, inc, 31);
            }
            if (i < f.n) {
                const LPt pi = f.pt[i]; const double zi = pi.z;
                int* out = p.tri ? p.tri + ((size_t)ofr * p.tri_cap) * 3 : nullptr;
                mine += (unsigned)lf_vertex_triangles_pts(f, i, pi, [&](int, int, const LPt& p1, const LPt& p2) {
                    a3 += tri_area3(pi.x, pi.y, zi, p1.x, p1.y, p1.z, p2.x, p2.y, p2.z);
                    if (want2d) a2 += tri_area3(pi.x, pi.y, 0.0, p1.x, p1.y, 0.0, p2.x, p2.y, 0.0);
                    if (out && (int)offs < p.tri_cap) { out[3 * offs] = pi.perm; out[3 * offs + 1] = p1.perm; out[3 * offs + 2] = p2.perm; }
                    ++offs;
                });
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mine += __shfl_down_sync(0xffffffffu, mine, o);
        base = __shfl_sync(0xffffffffu, mine, 0);
        total = base;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { a3 += __shfl_down_sync(0xffffffffu, a3, o); a2 += __shfl_down_sync(0xffffffffu, a2, o); }
    if (lane == 0) {
        if (st & ~GTA_B200_ST_TOO_FEW_POINTS) { a3 = __longlong_as_double(0x7ff8000000000000ll); a2 = a3; }
        if (p.area) p.area[ofr] = a3;
        if (p.area2d && (p.flags & GTA_B200_2D)) p.area2d[ofr] = a2;
        if (p.status) p.status[ofr] = st;
        if (p.ntri) p.ntri[ofr] = st ? 0 : (int)total;
    }
}

}  // namespace gt
