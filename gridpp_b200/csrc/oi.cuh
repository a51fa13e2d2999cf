// Observation-side state shared by the OI kernels (deterministic OI in oi.cu, EnSI in ensi.cu).
#pragma once

#include "points.cuh"
#include "structure.cuh"

namespace gpp {

// Device view of the observation table: VALID observations only, sorted by bucket-grid cell (stable, so
// ascending original index inside a cell). Passed to kernels by value.
struct ObsView {
    CellGeom geom;
    const int* cell_start;   // ncells + 1
    const float *x, *y, *z, *elev, *laf;
    const double* innov;     // (double) pobs - (double) pbackground            (oi.cpp:301-302,316)
    const float* ratio;      // obs_variance / bvariance_at_points, float       (oi.cpp:192-195)
    const int* orig;         // original observation index (selection tie-break, EnSI member lookup)
    int n;                   // number of valid observations
};

// Gathers up to `cap` candidates for one background point into per-warp shared memory, keeping only the
// `k` best by (rho descending, original index ascending) -- the streaming form of oi.cpp:233-273.
// Warp-synchronous; every lane returns the same count. On return the buffers hold the selected candidates
// sorted best-first.
struct CandBuf {
    float* rho;   // [cap]
    int* pos;     // [cap] slot in the sorted observation table
    int* orig;    // [cap] original index
};

__device__ __forceinline__ bool cand_better(float r1, int o1, float r2, int o2) { return r1 > r2 || (r1 == r2 && o1 < o2); }

// Rank-sort the first n entries (n <= 64) best-first and keep min(n, k). Each lane owns entries lane, lane+32.
__device__ __forceinline__ int cand_prune(const CandBuf& b, int n, int k) {
    unsigned lane = lane_id();
    float r0 = 0.f, r1 = 0.f;
    int p0 = 0, p1 = 0, o0 = 0, o1 = 0;
    bool h0 = (int) lane < n, h1 = (int) lane + 32 < n;
    if(h0) { r0 = b.rho[lane]; p0 = b.pos[lane]; o0 = b.orig[lane]; }
    if(h1) { r1 = b.rho[lane + 32]; p1 = b.pos[lane + 32]; o1 = b.orig[lane + 32]; }
    int rank0 = 0, rank1 = 0;
    for(int j = 0; j < n; j++) {
        float rj = b.rho[j];
        int oj = b.orig[j];
        rank0 += cand_better(rj, oj, r0, o0) ? 1 : 0;
        rank1 += cand_better(rj, oj, r1, o1) ? 1 : 0;
    }
    __syncwarp();
    if(h0 && rank0 < k) { b.rho[rank0] = r0; b.pos[rank0] = p0; b.orig[rank0] = o0; }
    if(h1 && rank1 < k) { b.rho[rank1] = r1; b.pos[rank1] = p1; b.orig[rank1] = o1; }
    __syncwarp();
    return min(n, k);
}

// need_pbackground_valid is folded into the table (invalid observations are not in it).
// Returns the number of selected observations (<= k <= 32).
__device__ __forceinline__ int gather_candidates(const ObsView& obs, const gpp_structure& s, const Pt& p1, float R, int k,
                                                 const CandBuf& b) {
    unsigned lane = lane_id();
    // kdtree.cpp:46-47: box corners are computed in float
    float lo0 = __fsub_rn(p1.x, R), lo1 = __fsub_rn(p1.y, R), lo2 = __fsub_rn(p1.z, R);
    float hi0 = __fadd_rn(p1.x, R), hi1 = __fadd_rn(p1.y, R), hi2 = __fadd_rn(p1.z, R);
    if(!(lo0 < hi0 && lo1 < hi1 && lo2 < hi2)) return 0;
    int cx0 = cell_coord(obs.geom, 0, lo0), cx1 = cell_coord(obs.geom, 0, hi0);
    int cy0 = cell_coord(obs.geom, 1, lo1), cy1 = cell_coord(obs.geom, 1, hi1);
    int cz0 = cell_coord(obs.geom, 2, lo2), cz1 = cell_coord(obs.geom, 2, hi2);
    int n = 0;
    for(int cz = cz0; cz <= cz1; cz++)
        for(int cy = cy0; cy <= cy1; cy++) {
            int base = (cz * obs.geom.n[1] + cy) * obs.geom.n[0];
            int s0 = obs.cell_start[base + cx0], s1 = obs.cell_start[base + cx1 + 1];
            for(int chunk = s0; chunk < s1; chunk += 32) {
                int i = chunk + (int) lane;
                float rho = 0.f;
                bool ok = i < s1;
                if(ok) {
                    float ox = obs.x[i], oy = obs.y[i], oz = obs.z[i];
                    // Boost within(): strictly inside the box (kdtree.cpp:46-53)
                    ok = ox > lo0 && ox < hi0 && oy > lo1 && oy < hi1 && oz > lo2 && oz < hi2;
                    if(ok) {
                        float dist = straight_distance(ox, oy, oz, p1.x, p1.y, p1.z);
                        ok = dist <= R;   // within_radius, kdtree.cpp:247-260
                        if(ok) {
                            Pt p2 = {ox, oy, oz, obs.elev[i], obs.laf[i]};
                            rho = structure_corr_background(s, p1, p2, dist);   // oi.cpp:250
                            ok = rho > 0.f;                                      // oi.cpp:253
                        }
                    }
                }
                unsigned mask = __ballot_sync(0xffffffffu, ok);
                if(ok) {
                    int slot = n + __popc(mask & ((1u << lane) - 1u));
                    b.rho[slot] = rho;
                    b.pos[slot] = i;
                    b.orig[slot] = obs.orig[i];
                }
                n += __popc(mask);
                __syncwarp();
                if(n > 32) n = cand_prune(b, n, k);   // keep room for the next 32
            }
        }
    if(n > 0) n = cand_prune(b, n, k);   // final selection, and canonical best-first order
    return n;
}

}  // namespace gpp

// Host-side owner of the observation table.
struct gpp_oi_obs {
    int n_total = 0;          // observations given
    int n_valid = 0;          // observations in the table
    float loc_dist = 0;       // localization distance the bucket grid was sized for
    gpp::CellGeom geom;
    int ncells = 0;
    gpp::DeviceBuffer<int> cell_start, orig;
    gpp::DeviceBuffer<float> x, y, z, elev, laf, ratio;
    gpp::DeviceBuffer<double> innov;
    gpp::ObsView view() const {
        gpp::ObsView v;
        v.geom = geom;
        v.cell_start = cell_start.ptr;
        v.x = x.ptr; v.y = y.ptr; v.z = z.ptr; v.elev = elev.ptr; v.laf = laf.ptr;
        v.innov = innov.ptr;
        v.ratio = ratio.ptr;
        v.orig = orig.ptr;
        v.n = n_valid;
        return v;
    }
};

namespace gpp {
// Builds the table from host arrays. `valid[i]` selects the observations that enter the table.
int build_obs_table(const gpp_points* opoints, const std::vector<char>& valid, const std::vector<double>& innov,
                    const std::vector<float>& ratio, float loc_dist, gpp_oi_obs* out);
}
