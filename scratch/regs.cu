// One wrapper kernel per walker phase: `nvcc -cubin -Xptxas -v` prints each phase's register need, and
// `cuobjdump -sass` its instruction count (the phase functions are out of line in the product kernel).
#include <cstdint>
#include "../g_tessla_b200/csrc/gta_lpf.cuh"
using namespace gt;
#define K(PH) extern "C" __global__ void k_##PH(LPt* pt, LRec* rec, LMeta* m, int n, int cap_e, int* out) { \
    extern __shared__ uint32_t sm[]; out[threadIdx.x] = lpf_wave_phase<PH, false>(sm, threadIdx.x, pt, rec, m, n, cap_e); }
K(LP_LEAF) K(LP_TINIT) K(LP_TAN) K(LP_INS) K(LP_SCAN) K(LP_V) K(LP_C)
