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
