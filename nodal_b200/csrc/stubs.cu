// Entry points that are declared in include/nodal_b200.h but not implemented yet.
#include "common.cuh"

#define NOT_YET(name)                                        \
    nodal_set_error(name ": not implemented in this build"); \
    return NODAL_BAD_ARG

extern "C" int nodal_dist_unique_id(uint8_t*) { NOT_YET("nodal_dist_unique_id"); }
extern "C" int nodal_dist_create(nodal_ctx*, const uint8_t*, int32_t, int32_t, nodal_dist**) {
    NOT_YET("nodal_dist_create");
}
extern "C" int nodal_dist_destroy(nodal_dist*) { return NODAL_OK; }
extern "C" int nodal_dist_pcg(nodal_ctx*, nodal_dist*, int32_t, int32_t, int32_t, const int32_t*,
                              const int32_t*, const double*, const double*, double*, double,
                              int32_t, int32_t*, double*, double*, void*) {
    NOT_YET("nodal_dist_pcg");
}
