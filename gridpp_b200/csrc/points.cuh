// gpp_points: the device-resident replacement for gridpp::Points / gridpp::KDTree / (flattened) gridpp::Grid.
#pragma once

#include <mutex>

#include "common.cuh"

namespace gpp {

// Uniform bucket grid over a point set ("K4" in SURVEY.md): points sorted by cell, cell_start offsets.
// Device arrays. The geometry is a POD so it can be passed to kernels by value.
struct CellGeom {
    float lo[3];     // lower corner of the bounding box
    float inv[3];    // cells per unit length (0 for a degenerate dimension)
    float edge[3];   // cell edge length
    int n[3];        // cells per dimension
};

__host__ __device__ __forceinline__ int cell_coord(const CellGeom& g, int d, float v) {
    // monotone non-decreasing in v: subtraction, multiplication by a non-negative constant and floor all are
    float t = floorf((v - g.lo[d]) * g.inv[d]);
    if(!(t > 0.f)) return 0;            // also catches NaN
    if(t > (float) (g.n[d] - 1)) return g.n[d] - 1;
    return (int) t;
}

struct CellIndex {
    CellGeom geom;
    DeviceBuffer<int> cell_start;   // ncells + 1
    DeviceBuffer<int> order;        // point index per sorted slot
    DeviceBuffer<float> sx, sy, sz; // coordinates in sorted order
    int ncells = 0;
    bool built = false;
};

}  // namespace gpp

struct gpp_points {
    int n = 0;
    int type = GPP_GEODETIC;
    int shape_ny = 0, shape_nx = 0;                        // set when the points are a flattened ny x nx grid
    bool has_elevs = false, has_lafs = false;              // given at creation (else all NaN: points.cpp:23-30, grid.cpp:41-54)
    std::vector<float> lats, lons, elevs, lafs, x, y, z;   // host copies
    float lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};             // bounding box of x/y/z
    int device = -1;
    std::mutex mutex;
    // device copies in original order (uploaded on first use)
    gpp::DeviceBuffer<float> dx, dy, dz, delev, dlaf;
    bool on_device = false;
    gpp::CellIndex index;

    int ensure_on_device();   // uploads x/y/z/elev/laf
    int ensure_index();       // builds the bucket grid on the device
};
