#include <cstdint>
#include "../g_tessla_b200/csrc/gta_lpf.cuh"
using namespace gt;
#define K(PH) extern "C" __global__ void k_##PH(uint32_t* st, Pt* pt, LRec* rec, u16* head, int n, int cap_e, uint32_t* eo, int* out) { \
    extern __shared__ uint32_t sm[]; out[threadIdx.x] = lpf_wave_phase<PH>(sm, threadIdx.x, pt, rec, head, n, cap_e, eo); }
K(LP_LEAF) K(LP_TINIT) K(LP_TAN) K(LP_INS) K(LP_SCAN) K(LP_V) K(LP_C)
