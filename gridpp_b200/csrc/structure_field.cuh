// Spatially varying structure-function scales: h, v, w on the nodes of a grid (gridpp::BarnesStructure(Grid, vec2 h,
// vec2 v, vec2 w, min_rho), structure.cpp:168-184, and the Soar / Toar / Powerlaw / Linear siblings). A point uses the
// scales of its nearest node (structure.cpp:189-199).
#pragma once

#include <vector>

#include "points.cuh"

struct gpp_structure_field {
    const gpp_points* grid = nullptr;   // not owned: must outlive the field
    std::vector<float> h, v, w;         // one value per grid node, row-major
};

namespace gpp {
int spatial_loc_constants(int type, float min_rho, float* loc_c, double* loc_d);            // capi.cu
float spatial_loc_dist_host(int type, float h, float loc_c, double loc_d);
}
