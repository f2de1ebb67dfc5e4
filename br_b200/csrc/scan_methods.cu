// scan_methods.cu — one translation unit per (correction method, variant): compiled ten times by
// br_b200/build.py with -DBRGPU_METHOD=0..4 and -DBRGPU_VARIANT=cnt|fast (-DBRGPU_COUNT_GETS=1|0), so
// that the five methods' scan kernels build in parallel and the product path carries no profiling
// bookkeeping.  correct_kernels.cu dispatches on (method, ctx->profiling).
#ifndef BRGPU_METHOD
#error "compile with -DBRGPU_METHOD=0..4"
#endif
#include "scan_device.cuh"

namespace brgpu {
namespace BRGPU_VARIANT {

#define BRGPU_CAT2(a, b) a##b
#define BRGPU_CAT(a, b) BRGPU_CAT2(a, b)

void BRGPU_CAT(launch_scan_m, BRGPU_METHOD)(const ScanArgs &a) {
    if (a.p.k == 17)
        launch_scan_method<BRGPU_METHOD, 17>(a);
    else
        launch_scan_method<BRGPU_METHOD, 0>(a);
}

} // namespace BRGPU_VARIANT
} // namespace brgpu
