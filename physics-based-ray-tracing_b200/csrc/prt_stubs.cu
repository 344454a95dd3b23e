// prt_stubs.cu -- entry points declared in include/prt_b200.h whose kernels are not built yet.
#include "prt_internal.h"
using namespace prt;
extern "C" {
int prt_das_beamform(prt_context *, const prt_das_params *, const float *, const float *, const double *, const float *, const float *,
                     float *, float *) {
    set_error("prt_das_beamform: not implemented in this build");
    return PRT_ERR_UNSUPPORTED;
}
}
