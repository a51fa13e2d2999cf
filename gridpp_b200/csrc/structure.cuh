// Device evaluation of the gridpp structure functions from the POD descriptor gpp_structure.
// Follows src/api/structure.cpp of the reference statement by statement: which operations run in float and
// which in double is part of the contract (the observation selection compares the resulting float rho).
#pragma once

#include "common.cuh"

namespace gpp {

struct Pt {
    float x, y, z, elev, laf;
};

// exp(x) for x <= 0 in fp64: Cody-Waite reduction by ln2 and a degree-13 Taylor polynomial on |r| <= ln2/2
// (truncation 4e-18, total error ~2 ulp). About 20 instructions against ~56 for the library exp(); the result is
// only ever rounded to float (barnes_rho, structure.cpp:32-33), where it differs from the correctly rounded value
// in ~1e-8 of cases -- the same class of deviation as between two libm implementations.
__device__ __forceinline__ double exp_nonpos(double x) {
    const double magic = 6755399441055744.0;   // 1.5 * 2^52: the add rounds to the nearest integer
    double t = fma(x, 1.4426950408889634, magic);
    int n = __double2loint(t);
    double fn = t - magic;
    double r = fma(fn, -6.93147180369123816490e-01, x);
    r = fma(fn, -1.90821492927058770002e-10, r);
    double p = 1.6059043836821613e-10;
    p = fma(p, r, 2.08767569878681e-09);
    p = fma(p, r, 2.505210838544172e-08);
    p = fma(p, r, 2.755731922398589e-07);
    p = fma(p, r, 2.7557319223985893e-06);
    p = fma(p, r, 2.48015873015873e-05);
    p = fma(p, r, 1.984126984126984e-04);
    p = fma(p, r, 1.388888888888889e-03);
    p = fma(p, r, 8.333333333333333e-03);
    p = fma(p, r, 4.1666666666666664e-02);
    p = fma(p, r, 1.6666666666666666e-01);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    double y = __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));
    return x < -700.0 ? 0.0 : y;
}

// exp() of a float argument whose reference evaluation is the float overload (soar_rho/toar_rho,
// structure.cpp:53,63). Going through the double exp gives the correctly rounded float in all but ~1e-9 of
// cases, which is what glibc's expf (< 0.51 ulp) returns too.
__device__ __forceinline__ float exp_as_float(float x) { return (float) exp((double) x); }

// structure.cpp:26-87
__device__ __forceinline__ float term_rho(int type, float dist, float length) {
    if(type == GPP_STRUCT_LINEAR) {
        // linear_rho(diff, min_corr), structure.cpp:76-87
        if(!is_valid(length) || length < 0) return 1.f;
        if(!is_valid(dist)) return 0.f;
        float absdiff = fabsf(dist);
        if(absdiff > 1.f) absdiff = 1.f;
        return __fsub_rn(1.f, __fmul_rn(__fsub_rn(1.f, length), absdiff));
    }
    if(!is_valid(length) || length == 0.f) return 1.f;   // disabled
    if(!is_valid(dist)) return 0.f;
    if(type == GPP_STRUCT_CRESSMAN) {
        // structure.cpp:35-44
        if(dist >= length) return 0.f;
        float l2 = __fmul_rn(length, length), d2 = __fmul_rn(dist, dist);
        return __fdiv_rn(__fsub_rn(l2, d2), __fadd_rn(l2, d2));
    }
    float v = __fdiv_rn(dist, length);
    if(type == GPP_STRUCT_BARNES) {
        // structure.cpp:32-33: exp(-0.5 * v * v) in double (the products of two floats are exact in double)
        double dv = (double) v;
        return (float) exp_nonpos(__dmul_rn(__dmul_rn(-0.5, dv), dv));
    }
    if(type == GPP_STRUCT_SOAR) {
        // structure.cpp:52-53: (1 + v) * exp(-v) in float
        return __fmul_rn(__fadd_rn(1.f, v), exp_as_float(-v));
    }
    if(type == GPP_STRUCT_TOAR) {
        // structure.cpp:62-63: (1 + v + (v*v)/3) * exp(-v) in float
        float poly = __fadd_rn(__fadd_rn(1.f, v), __fdiv_rn(__fmul_rn(v, v), 3.f));
        return __fmul_rn(poly, exp_as_float(-v));
    }
    // GPP_STRUCT_POWERLAW, structure.cpp:72-73: 1 / (1 + 0.5 * v * v) in double
    double dv = (double) v;
    return (float) __ddiv_rn(1.0, __dadd_rn(1.0, __dmul_rn(__dmul_rn(0.5, dv), dv)));
}

// <Family>Structure::corr, non-spatial branch (Barnes structure.cpp:214-228, Soar :388-402, Toar :538-552,
// Powerlaw :689-703, Linear :836-850) and CressmanStructure::corr (:298-309, no localization test).
// hdist is calc_straight_distance(p1, p2), passed in because callers already have it.
__device__ __forceinline__ float term_corr_typed(int type, const gpp_structure_term& t, float hdist, float elev1, float laf1,
                                                 float elev2, float laf2) {
    if(type != GPP_STRUCT_CRESSMAN && hdist > t.loc_dist) return 0.f;
    float rho = term_rho(type, hdist, t.h);
    if(is_valid(elev1) && is_valid(elev2)) rho = __fmul_rn(rho, term_rho(type, __fsub_rn(elev1, elev2), t.v));
    if(is_valid(laf1) && is_valid(laf2)) rho = __fmul_rn(rho, term_rho(type, __fsub_rn(laf1, laf2), t.w));
    return rho;
}
__device__ __forceinline__ float term_corr(const gpp_structure_term& t, float hdist, float elev1, float laf1, float elev2,
                                           float laf2) {
    return term_corr_typed(t.type, t, hdist, elev1, laf1, elev2, laf2);
}

// Compile-time specialisation used by the OI kernels: SMODE 1 = a single Barnes term without cross-validation
// (the common case; all family dispatch folds away), SMODE 0 = any descriptor.
template <int SMODE>
__device__ __forceinline__ float corr_mode(const gpp_structure& s, const Pt& p1, const Pt& p2, float hdist);
template <int SMODE>
__device__ __forceinline__ float corr_background_mode(const gpp_structure& s, const Pt& p1, const Pt& p2, float hdist);

// StructureFunction::corr for the descriptor: a plain term, or MultipleStructure::corr (structure.cpp:98-112)
// where the horizontal term sees p2 with p1's elevation/laf, the vertical term sees p1's position with p2's
// elevation, and the land/sea term p1's position and elevation with p2's laf.
__device__ __forceinline__ float structure_corr(const gpp_structure& s, const Pt& p1, const Pt& p2, float hdist) {
    if(s.n_terms != 3) return term_corr(s.term[0], hdist, p1.elev, p1.laf, p2.elev, p2.laf);
    float corr_h = term_corr(s.term[0], hdist, p1.elev, p1.laf, p1.elev, p1.laf);
    float corr_v = term_corr(s.term[1], 0.f, p1.elev, p1.laf, p2.elev, p1.laf);
    float corr_w = term_corr(s.term[2], 0.f, p1.elev, p1.laf, p1.elev, p2.laf);
    return __fmul_rn(__fmul_rn(corr_h, corr_v), corr_w);
}
__device__ __forceinline__ float structure_corr(const gpp_structure& s, const Pt& p1, const Pt& p2) {
    return structure_corr(s, p1, p2, straight_distance(p1.x, p1.y, p1.z, p2.x, p2.y, p2.z));
}
// corr_background: base class structure.cpp:20-25, CrossValidation :919-935
__device__ __forceinline__ float structure_corr_background(const gpp_structure& s, const Pt& p1, const Pt& p2, float hdist) {
    if(s.has_cv && is_valid(s.cv_dist) && hdist <= s.cv_dist) return 0.f;
    return structure_corr(s, p1, p2, hdist);
}

template <>
__device__ __forceinline__ float corr_mode<0>(const gpp_structure& s, const Pt& p1, const Pt& p2, float hdist) {
    return structure_corr(s, p1, p2, hdist);
}
template <>
__device__ __forceinline__ float corr_mode<1>(const gpp_structure& s, const Pt& p1, const Pt& p2, float hdist) {
    return term_corr_typed(GPP_STRUCT_BARNES, s.term[0], hdist, p1.elev, p1.laf, p2.elev, p2.laf);
}
template <>
__device__ __forceinline__ float corr_background_mode<0>(const gpp_structure& s, const Pt& p1, const Pt& p2, float hdist) {
    return structure_corr_background(s, p1, p2, hdist);
}
template <>
__device__ __forceinline__ float corr_background_mode<1>(const gpp_structure& s, const Pt& p1, const Pt& p2, float hdist) {
    return term_corr_typed(GPP_STRUCT_BARNES, s.term[0], hdist, p1.elev, p1.laf, p2.elev, p2.laf);
}
// Out-of-line forms for the OI kernels: the evaluation is ~60 instructions and has several call sites per kernel;
// inlining every one of them made the hot kernel 119 KB of code, which stalled on instruction fetch.
template <int SMODE>
__device__ __noinline__ float corr_call(const gpp_structure& s, const Pt& p1, const Pt& p2, float hdist) {
    return corr_mode<SMODE>(s, p1, p2, hdist);
}
template <int SMODE>
__device__ __noinline__ float corr_background_call(const gpp_structure& s, const Pt& p1, const Pt& p2, float hdist) {
    return corr_background_mode<SMODE>(s, p1, p2, hdist);
}
// (mode 1 also requires an active horizontal scale: the kernels' inlined Barnes evaluation skips barnes_rho's
// "length invalid or zero" early return, structure.cpp:27-29)
inline int structure_mode(const gpp_structure& s) {
    return (s.n_terms == 1 && !s.has_cv && s.term[0].type == GPP_STRUCT_BARNES && is_valid(s.term[0].h) && s.term[0].h > 0.f) ? 1 : 0;
}

// A descriptor whose horizontal scale is NaN is the placeholder the host layers build for a spatially varying
// <Family>Structure(Grid, h, v, w) (the scales live in a gpp_structure_field, not in the descriptor). Entry points that
// take a plain descriptor must refuse it instead of analysing with a localization distance of NaN / 0.
inline int reject_unset_scales(const gpp_structure* s) {
    for(int t = 0; s && t < s->n_terms && t < 3; t++)
        if(isnan(s->term[t].h))
            return fail(GPP_ERR_NOT_IMPLEMENTED, "spatially varying structure functions are only supported by optimal_interpolation / optimal_interpolation_full "
                                                 "(gpp_optimal_interpolation_spatial_host)");
    return GPP_OK;
}

// True when corr(p1, p2) == corr(p2, p1) for every pair, so that P + R is symmetric (positive definite) and
// the symmetric elimination kernel applies. Cressman, Soar and Toar are NOT even functions of the elevation /
// laf difference in the reference (structure.cpp:41-43,52-53,62-63 use the signed difference), so any of them
// with an active vertical or land/sea scale makes the matrix non-symmetric.
inline bool structure_is_symmetric(const gpp_structure& s) {
    auto odd = [](int type) { return type == GPP_STRUCT_CRESSMAN || type == GPP_STRUCT_SOAR || type == GPP_STRUCT_TOAR; };
    auto active = [](float len) { return is_valid(len) && len != 0.f; };
    if(s.n_terms != 3) {
        const gpp_structure_term& t = s.term[0];
        return !(odd(t.type) && (active(t.v) || active(t.w)));
    }
    // MultipleStructure: term[1] only ever sees an elevation difference, term[2] only a laf difference
    if(odd(s.term[1].type) && active(s.term[1].v)) return false;
    if(odd(s.term[2].type) && active(s.term[2].w)) return false;
    return true;
}

}  // namespace gpp
