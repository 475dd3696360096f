#include "../g_tessla_b200/csrc/dt_coop.cuh"
using namespace gt;
__global__ void __launch_bounds__(128, 3) k(Pt* pt, u16* head, u16* dst, u16* nxt, u16* prv, uint32_t* ctl, int n, int mid) {
    extern __shared__ unsigned char sm[];
    Frame f; f.pt = (Pt*)sm; f.head = (u16*)(sm + 20000); f.dst = (u16*)(sm + 24000); f.nxt = (u16*)(sm + 40000); f.prv = (u16*)(sm + 56000);
    f.bump = (uint32_t*)(sm + 72000); f.stack = f.bump + 1; f.err = f.bump + 2; f.n = n; f.cap_e = 3 * n;
    Lane ln; ln.free_head = GT_NIL;
    coop_merge(f, ln, mid);
    if (ln.free_head != GT_NIL) ctl[0] = ln.free_head;
}
