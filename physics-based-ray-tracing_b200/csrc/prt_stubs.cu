// prt_stubs.cu -- entry points declared in include/prt_b200.h whose kernels are not built yet.
#include "prt_internal.h"
using namespace prt;
extern "C" {
int prt_render_path(prt_scene *, const prt_render_params *, uint64_t, uint32_t, uint32_t, uint32_t, float *, prt_render_stats *) {
    set_error("prt_render_path: not implemented in this build");
    return PRT_ERR_UNSUPPORTED;
}
int prt_render_path_dev(prt_scene *, const prt_render_params *, uint64_t, uint32_t, uint32_t, uint32_t, float *, uint64_t *, void *) {
    set_error("prt_render_path_dev: not implemented in this build");
    return PRT_ERR_UNSUPPORTED;
}
int prt_das_beamform(prt_context *, const prt_das_params *, const float *, const float *, const double *, const float *, const float *,
                     float *, float *) {
    set_error("prt_das_beamform: not implemented in this build");
    return PRT_ERR_UNSUPPORTED;
}
}
