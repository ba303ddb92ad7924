// precond.cu -- K4: exact mean-based preconditioner for a constant-coefficient 5-point mean block.
// (round-1 placeholder: the handle type exists, the fast-diagonalisation kernels land next)
#include "precond.h"

using namespace spuq;

extern "C" int spuq_sg_precond_mean_fd(spuq_sg_precond** out, int device, int32_t nx, int32_t ny,
                                       double d, double ox, double oy) {
    (void)out; (void)device; (void)nx; (void)ny; (void)d; (void)ox; (void)oy;
    set_error("precond_mean_fd: not built yet");
    return SPUQ_E_UNSUPPORTED;
}

namespace spuq {
int mean_fd_destroy(spuq_sg_precond*) { return SPUQ_OK; }
int64_t mean_fd_work_bytes(const spuq_sg_precond*, int64_t, int32_t) { return 0; }
int mean_fd_apply(spuq_sg_precond*, int64_t, int32_t, const double*, double*, void*, cudaStream_t,
                  const double*) {
    set_error("precond_mean_fd: not built yet");
    return SPUQ_E_UNSUPPORTED;
}
}
