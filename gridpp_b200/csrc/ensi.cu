// Ensemble optimal interpolation (EnSI) -- placeholder until the kernel lands.
#include "oi.cuh"

extern "C" int gpp_optimal_interpolation_ensi_host(const gpp_points*, const float*, int, const gpp_points*, const float*,
                                                   const float*, const float*, const gpp_structure*, int, int, float*, int*) {
    return gpp::fail(GPP_ERR_NOT_IMPLEMENTED, "optimal_interpolation_ensi is not implemented on the device yet");
}
