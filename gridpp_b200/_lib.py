"""ctypes binding of libgridpp_b200.so (the C ABI declared in include/gridpp_b200.h).

The shared library is the product; this module only loads it and declares the prototypes. There is no CPU
fallback: if the library is missing, import fails with instructions to build it.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# GPP_B200_LIB: developer override used to A/B kernel builds (profiles/variants.sh); the default is the in-tree library
LIB_PATH = os.environ.get("GPP_B200_LIB") or os.path.join(HERE, "libgridpp_b200.so")

OK, ERR_INVALID_ARGUMENT, ERR_RUNTIME, ERR_NOT_IMPLEMENTED, ERR_CUDA = range(5)


class StructureTerm(C.Structure):
    _fields_ = [("type", C.c_int), ("h", C.c_float), ("v", C.c_float), ("w", C.c_float),
                ("min_rho", C.c_float), ("loc_dist", C.c_float)]


class StructureDesc(C.Structure):
    """gpp_structure"""
    _fields_ = [("n_terms", C.c_int), ("term", StructureTerm * 3), ("has_cv", C.c_int), ("cv_dist", C.c_float)]


fp = C.POINTER(C.c_float)
ip = C.POINTER(C.c_int)
vp = C.c_void_p
sp = C.POINTER(StructureDesc)

if not os.path.exists(LIB_PATH):
    raise ImportError(
        "gridpp_b200: %s is missing. Build it with `python -m gridpp_b200.build` (needs nvcc; targets sm_100a). "
        "There is no CPU fallback." % LIB_PATH)

lib = C.CDLL(LIB_PATH)


def _proto(name, restype, *argtypes):
    fn = getattr(lib, name)
    fn.restype = restype
    fn.argtypes = list(argtypes)
    return fn


_proto("gpp_version", C.c_char_p)
_proto("gpp_last_error", C.c_char_p)
_proto("gpp_device_count", C.c_int, ip)
_proto("gpp_set_device", C.c_int, C.c_int)
_proto("gpp_device_synchronize", C.c_int)
_proto("gpp_kernel_launch_count", C.c_ulonglong)
_proto("gpp_measure_fp64_fma_peak", C.c_int, C.POINTER(C.c_double))
_proto("gpp_structure_init", C.c_int, sp, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float)
_proto("gpp_structure_init_min_rho", C.c_int, sp, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float)
_proto("gpp_structure_multiple", C.c_int, sp, sp, sp, sp)
_proto("gpp_structure_cross_validation", C.c_int, sp, sp, C.c_float)
_proto("gpp_structure_corr_host", C.c_int, sp, fp, fp, C.c_int, C.c_int, fp)
_proto("gpp_points_create", C.c_int, fp, fp, fp, fp, C.c_int, C.c_int, C.POINTER(vp))
_proto("gpp_points_destroy", None, vp)
_proto("gpp_points_set_shape", C.c_int, vp, C.c_int, C.c_int)
_proto("gpp_points_size", C.c_int, vp)
_proto("gpp_points_coordinate_type", C.c_int, vp)
_proto("gpp_points_get_xyz", C.c_int, vp, fp, fp, fp)
_proto("gpp_convert_coordinates", C.c_int, fp, fp, C.c_int, C.c_int, fp, fp, fp)
_proto("gpp_points_nearest_host", C.c_int, vp, fp, fp, C.c_int, C.c_int, ip)
_proto("gpp_points_neighbours_host", C.c_int, vp, fp, fp, fp, C.c_int, C.c_int, C.c_int, ip, fp, ip)
_proto("gpp_points_closest_host", C.c_int, vp, fp, fp, C.c_int, C.c_int, C.c_int, ip)
_proto("gpp_nearest_host", C.c_int, vp, fp, fp, C.c_int, fp, C.c_int, fp)
_proto("gpp_optimal_interpolation_host", C.c_int, vp, fp, fp, vp, fp, fp, fp, fp, sp, C.c_int, C.c_int, fp, fp)
_proto("gpp_oi_obs_create", C.c_int, vp, fp, fp, fp, fp, sp, C.POINTER(vp))
_proto("gpp_oi_obs_destroy", None, vp)
_proto("gpp_optimal_interpolation_device", C.c_int, vp, C.c_int, C.c_int, vp, vp, vp, sp, C.c_int, C.c_int, vp, vp, vp)
_proto("gpp_oi_workspace_bytes", C.c_size_t)
_proto("gpp_optimal_interpolation_device_ws", C.c_int, vp, C.c_int, C.c_int, vp, vp, vp, sp, C.c_int, C.c_int, vp, vp, vp, C.c_size_t, vp)
_proto("gpp_ensi_obs_create", C.c_int, vp, fp, fp, fp, C.c_int, ip, sp, C.POINTER(vp))
_proto("gpp_ensi_obs_destroy", None, vp)
_proto("gpp_ensi_valid_members_device", C.c_int, vp, C.c_longlong, C.c_int, ip, vp)
_proto("gpp_optimal_interpolation_ensi_device", C.c_int, vp, C.c_int, C.c_int, vp, C.c_int, vp, sp, C.c_int, C.c_int, vp, vp, vp)
_proto("gpp_optimal_interpolation_multi_gpu_host", C.c_int, C.c_int, vp, fp, fp, vp, fp, fp, fp, fp, sp, C.c_int, C.c_int, fp, fp)
_proto("gpp_halo_pull_device", C.c_int, vp, C.c_int, C.c_int, C.c_int, vp, vp, vp)
_proto("gpp_optimal_interpolation_ensi_host", C.c_int, vp, fp, C.c_int, vp, fp, fp, fp, sp, C.c_int, C.c_int, fp, ip)
_proto("gpp_optimal_interpolation_ensi_multi_ebe_host", C.c_int, vp, fp, fp, fp, C.c_int, vp, fp, fp, fp, fp, sp, C.c_int, C.c_int, fp)
_proto("gpp_optimal_interpolation_ensi_multi_ebesc_host", C.c_int, vp, fp, fp, C.c_int, vp, fp, fp, fp, sp, C.c_int, C.c_int, fp)
_proto("gpp_optimal_interpolation_ensi_multi_utem_host", C.c_int, vp, fp, fp, fp, C.c_int, vp, fp, fp, fp, fp, sp, C.c_int, C.c_int, fp, ip)
_proto("gpp_staticcorr_points_host", C.c_int, vp, vp, sp, C.c_int, fp)
_proto("gpp_neighbourhood_search_host", C.c_int, fp, fp, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, ip, fp)
_proto("gpp_calc_gradient_host", C.c_int, fp, fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, fp)
_proto("gpp_neighbourhood_host", C.c_int, fp, C.c_int, C.c_int, C.c_int, C.c_int, fp)
_proto("gpp_neighbourhood_device", C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp)
_proto("gpp_neighbourhood_quantile_fast_host", C.c_int, fp, C.c_int, C.c_int, C.c_float, fp, C.c_int, fp, C.c_int, fp)
_proto("gpp_neighbourhood_quantile_fast_device", C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, vp, C.c_int,
       fp, C.c_int, vp, vp)

_proto("gpp_neighbourhood_ens_host", C.c_int, fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, fp)
_proto("gpp_neighbourhood_ens_device", C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp)
_proto("gpp_neighbourhood_quantile_fast_ens_host", C.c_int, fp, C.c_int, C.c_int, C.c_int, C.c_float, fp, C.c_int, fp, C.c_int, fp)
_proto("gpp_neighbourhood_quantile_fast_ens_device", C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_float, vp, C.c_int, fp, C.c_int, vp, vp)

_proto("gpp_gridding_host", C.c_int, vp, vp, fp, C.c_float, C.c_int, C.c_int, fp)
_proto("gpp_gridding_nearest_host", C.c_int, vp, vp, fp, C.c_int, C.c_int, fp)
_proto("gpp_count_host", C.c_int, vp, vp, C.c_float, fp)
_proto("gpp_distance_host", C.c_int, vp, vp, C.c_int, C.c_int, fp)
_proto("gpp_fill_host", C.c_int, vp, fp, vp, fp, C.c_float, C.c_int, fp)
_proto("gpp_fill_missing_host", C.c_int, fp, C.c_int, C.c_int, fp)
_proto("gpp_doping_square_host", C.c_int, vp, fp, vp, fp, ip, C.c_float, fp)
_proto("gpp_doping_circle_host", C.c_int, vp, fp, vp, fp, fp, C.c_float, fp)
_proto("gpp_calc_statistic_host", C.c_int, fp, C.c_longlong, C.c_int, C.c_int, fp)
_proto("gpp_calc_statistic_device", C.c_int, vp, C.c_longlong, C.c_int, C.c_int, vp, vp)
_proto("gpp_neighbourhood_brute_force_host", C.c_int, fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, fp)
_proto("gpp_neighbourhood_brute_force_device", C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, vp, vp)
_proto("gpp_calc_quantile_host", C.c_int, fp, C.c_longlong, C.c_int, C.c_float, fp, fp)
_proto("gpp_interpolate_host", C.c_int, fp, C.c_longlong, fp, fp, C.c_int, fp)
_proto("gpp_structure_field_create", C.c_int, vp, fp, fp, fp, C.POINTER(vp))
_proto("gpp_structure_field_destroy", None, vp)
_proto("gpp_structure_field_lookup_host", C.c_int, vp, fp, fp, C.c_int, fp, fp, fp)
_proto("gpp_structure_field_localization_distance", C.c_int, vp, C.c_int, C.c_float, C.c_float, C.c_float, fp)
_proto("gpp_optimal_interpolation_spatial_host", C.c_int, vp, fp, fp, vp, fp, fp, fp, fp, C.c_int, vp, C.c_float, C.c_int, C.c_int, fp, fp)
_proto("gpp_get_neighbourhood_thresholds_host", C.c_int, fp, C.c_longlong, C.c_int, fp, ip)

EXPORTS = [
    "gpp_version", "gpp_last_error", "gpp_device_count", "gpp_set_device", "gpp_device_synchronize",
    "gpp_kernel_launch_count", "gpp_measure_fp64_fma_peak", "gpp_structure_init", "gpp_structure_init_min_rho", "gpp_structure_multiple", "gpp_structure_cross_validation",
    "gpp_structure_corr_host", "gpp_points_create", "gpp_points_destroy", "gpp_points_set_shape", "gpp_points_size",
    "gpp_points_coordinate_type", "gpp_points_get_xyz", "gpp_convert_coordinates", "gpp_points_nearest_host", "gpp_points_neighbours_host",
    "gpp_points_closest_host", "gpp_nearest_host", "gpp_optimal_interpolation_host", "gpp_oi_obs_create",
    "gpp_oi_obs_destroy", "gpp_optimal_interpolation_device", "gpp_optimal_interpolation_ensi_host",
    "gpp_oi_workspace_bytes", "gpp_optimal_interpolation_device_ws", "gpp_optimal_interpolation_multi_gpu_host", "gpp_halo_pull_device", "gpp_ensi_obs_create", "gpp_ensi_obs_destroy",
    "gpp_ensi_valid_members_device", "gpp_optimal_interpolation_ensi_device",
    "gpp_optimal_interpolation_ensi_multi_ebe_host", "gpp_optimal_interpolation_ensi_multi_ebesc_host",
    "gpp_optimal_interpolation_ensi_multi_utem_host", "gpp_staticcorr_points_host",
    "gpp_neighbourhood_host", "gpp_neighbourhood_device", "gpp_neighbourhood_quantile_fast_host",
    "gpp_neighbourhood_quantile_fast_device", "gpp_neighbourhood_ens_host", "gpp_neighbourhood_ens_device",
    "gpp_neighbourhood_quantile_fast_ens_host", "gpp_neighbourhood_quantile_fast_ens_device",
    "gpp_gridding_host", "gpp_gridding_nearest_host", "gpp_count_host", "gpp_distance_host", "gpp_fill_host", "gpp_fill_missing_host",
    "gpp_doping_square_host", "gpp_doping_circle_host",
    "gpp_calc_statistic_host", "gpp_calc_statistic_device", "gpp_calc_quantile_host", "gpp_interpolate_host",
    "gpp_neighbourhood_brute_force_host", "gpp_neighbourhood_brute_force_device",
    "gpp_neighbourhood_search_host", "gpp_calc_gradient_host",
    "gpp_get_neighbourhood_thresholds_host", "gpp_structure_field_create", "gpp_structure_field_destroy",
    "gpp_structure_field_lookup_host", "gpp_structure_field_localization_distance", "gpp_optimal_interpolation_spatial_host",
]


class NotImplementedOnDevice(RuntimeError):
    pass


def check(rc):
    """Maps status codes to the exceptions the reference's SWIG layer raises (swig/gridpp.i:21-40)."""
    if rc == OK:
        return
    msg = lib.gpp_last_error().decode(errors="replace")
    if rc == ERR_INVALID_ARGUMENT:
        raise ValueError(msg)
    if rc == ERR_NOT_IMPLEMENTED:
        raise NotImplementedOnDevice(msg)
    raise RuntimeError(msg)
