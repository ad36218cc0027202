// hs_tick_exact.cu -- the control tick (hs_tick_kernel) compiled a second time with IEEE arithmetic.
//
// Same source as the product kernel (hs_common.cuh / hs_stages.cuh / hs_tick.cuh) with HS_EXACT_MATH=1: divisions,
// square roots and exp are round-to-nearest IEEE operations, nothing is contracted to FMA (this translation unit is
// built with -fmad=false), and the ill-conditioned stages (evader force, line-of-sight test, stats division) follow
// the reference's own operation order.  Selected at run time with hs_set_option(h, HS_OPT_EXACT_MATH, 1); ~2x slower
// than the product kernel.  Purpose: the parity tests run every fixture through BOTH builds - whatever the fast build
// gets "wrong" must be right here, i.e. be a rounding-level difference amplified by a discontinuity of the task
// (evader velocity = v * f / (|f| + 1e-5) per component, reward indicators), not a defect.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include <math.h>

#include "hs_b200.h"

#define HS_EXACT_MATH 1
#ifndef HS_USE_BULK_STORE
#define HS_USE_BULK_STORE 1
#endif
#include "hs_common.cuh"
#include "hs_stages.cuh"
#include "hs_tick.cuh"
#include "hs_tick_wide.cuh"

// kparams: the caller's KParams (same header, same layout), passed by address because types in anonymous namespaces
// are private to their translation unit
cudaError_t hs_launch_tick_exact(const void* kparams, size_t bytes, int num_agents, int reset, int small_c, unsigned grid,
                                 unsigned block, cudaStream_t s) {
    KParams P;
    if (bytes != sizeof(P)) return cudaErrorInvalidValue;
    memcpy(&P, kparams, sizeof(P));
#define HS_X(AA, RR) do { if (small_c) hs_tick_kernel<AA, RR, 5><<<grid, block, 0, s>>>(P); \
                          else hs_tick_kernel<AA, RR, CMAX><<<grid, block, 0, s>>>(P); } while (0)
    switch (num_agents) {
        case 1: if (reset) HS_X(1, true); else HS_X(1, false); break;
        case 2: if (reset) HS_X(2, true); else HS_X(2, false); break;
        case 3: if (reset) HS_X(3, true); else HS_X(3, false); break;
        default: return cudaErrorInvalidValue;
    }
#undef HS_X
    return cudaGetLastError();
}

// the one-lane-per-env mapping in IEEE arithmetic (A = 3 and 4: what the parity tests exercise)
cudaError_t hs_wide_exact_set_smem(int num_agents, int small_c, int bytes) {
    cudaError_t e = cudaSuccess;
#define HS_S(AA, CC) do { e = cudaFuncSetAttribute(hs_tick_wide_kernel<AA, CC, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes); \
                          if (e == cudaSuccess) e = cudaFuncSetAttribute(hs_tick_wide_kernel<AA, CC, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes); } while (0)
    if (num_agents == 3) { if (small_c) HS_S(3, 5); else HS_S(3, CMAX); }
    else if (num_agents == 4) { if (small_c) HS_S(4, 5); else HS_S(4, CMAX); }
#undef HS_S
    return e;                          // other agent counts: no exact build (hs_launch_tick_wide_exact reports it)
}
cudaError_t hs_launch_tick_wide_exact(const void* kparams, size_t bytes, const void* maps3, int num_agents, int reset, int small_c,
                                      unsigned grid, size_t smem, cudaStream_t s) {
    KParams P;
    if (bytes != sizeof(P)) return cudaErrorInvalidValue;
    memcpy(&P, kparams, sizeof(P));
    const CUtensorMap* tm = static_cast<const CUtensorMap*>(maps3);
#define HS_W(AA, CC, RR) hs_tick_wide_kernel<AA, CC, RR><<<grid, ((AA) + 1) * 32, smem, s>>>(P, tm[0], tm[1])
#define HS_WC(AA, RR) do { if (small_c) HS_W(AA, 5, RR); else HS_W(AA, CMAX, RR); } while (0)
    if (num_agents == 3) { if (reset) HS_WC(3, true); else HS_WC(3, false); }
    else if (num_agents == 4) { if (reset) HS_WC(4, true); else HS_WC(4, false); }
    else return cudaErrorInvalidValue;
#undef HS_WC
#undef HS_W
    return cudaGetLastError();
}
