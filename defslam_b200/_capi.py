"""ctypes mirror of include/defslam_b200.h.

This module only *describes* the C ABI (struct layouts + prototypes) and loads the
in-tree CUDA library.  There is no CPU fallback: if ``libdefslam_b200.so`` is missing,
``load()`` raises, and every entry point of the library itself returns
``DEFSLAM_ECUDA`` when no CUDA device is usable.
"""
from __future__ import annotations

import ctypes as C
import os

c_double_p = C.POINTER(C.c_double)
c_float_p = C.POINTER(C.c_float)
c_int32_p = C.POINTER(C.c_int32)
c_uint8_p = C.POINTER(C.c_uint8)

OK, EBADARG, ECUDA, ENUMERIC, ETOOLARGE, ENOTIMPL = 0, -1, -2, -3, -4, -5
c_int64_p = C.POINTER(C.c_int64)


class TemplateDesc(C.Structure):
    _fields_ = [
        ("n_nodes", C.c_int32),
        ("n_edges", C.c_int32),
        ("n_facets", C.c_int32),
        ("node_rest_xyz", c_double_p),
        ("node_boundary", c_uint8_p),
        ("nbr_ptr", c_int32_p),
        ("nbr_idx", c_int32_p),
        ("nbr_w", c_double_p),
        ("node_kappa0", c_double_p),
        ("edge_ab", c_int32_p),
        ("edge_len0", c_double_p),
        ("facets", c_int32_p),
        ("edge_median_len", C.c_double),
    ]


class SftProblem(C.Structure):
    _fields_ = [
        ("tmpl", C.c_void_p),
        ("tmpl_desc", C.POINTER(TemplateDesc)),
        ("node_xyz", c_double_p),
        ("n_matches", C.c_int32),
        ("n_frame_keypoints", C.c_int32),
        ("match_nodes", c_int32_p),
        ("match_bary", c_double_p),
        ("match_uv", c_float_p),
        ("match_inv_sigma2", c_float_p),
        ("fx", C.c_double),
        ("fy", C.c_double),
        ("cx", C.c_double),
        ("cy", C.c_double),
        ("T_cw", C.c_float * 16),
        ("reg_lap", C.c_double),
        ("reg_inex", C.c_double),
        ("reg_temp", C.c_double),
        ("neighbour_layers", C.c_int32),
        ("max_iterations", C.c_int32),
        ("matches_given", C.c_int32),
        ("curv_edge_len", C.c_double),
    ]


class SftResult(C.Structure):
    _fields_ = [
        ("node_xyz_out", c_double_p),
        ("outlier_out", c_uint8_p),
        ("node_role_out", c_uint8_p),
        ("T_cw_out", C.c_float * 16),
        ("rep_error", C.c_float),
        ("n_inliers", C.c_int32),
        ("lm_iterations", C.c_int32),
        ("lm_trials", C.c_int32),
        ("chi2_initial", C.c_double),
        ("chi2_final", C.c_double),
        ("lambda_final", C.c_double),
        ("trace", c_double_p),
        ("trace_capacity", C.c_int32),
        ("status", C.c_int32),
    ]


class Bbs(C.Structure):
    _fields_ = [
        ("umin", C.c_double),
        ("umax", C.c_double),
        ("nptsu", C.c_int32),
        ("vmin", C.c_double),
        ("vmax", C.c_double),
        ("nptsv", C.c_int32),
        ("valdim", C.c_int32),
    ]


class SchwarpProblem(C.Structure):
    _fields_ = [
        ("bbs", Bbs),
        ("n_matches", C.c_int32),
        ("kp1", c_float_p),
        ("kp2", c_float_p),
        ("inv_sigma", c_float_p),
        ("lambda_", C.c_double),
        ("fx", C.c_double),
        ("fy", C.c_double),
        ("px_fx", C.c_double),
        ("px_fy", C.c_double),
        ("max_iterations", C.c_int32),
        ("initialize", C.c_int32),
        ("x", c_double_p),
    ]


class DiffProp(C.Structure):
    _fields_ = [
        ("warp_uv", c_float_p),
        ("J12", c_float_p),
        ("J21", c_float_p),
        ("H12", c_float_p),
        ("keep", c_uint8_p),
        ("cost_initial", C.c_double),
        ("cost_final", C.c_double),
        ("iterations", C.c_int32),
        ("accepted", C.c_int32),
    ]


class NormalsProblem(C.Structure):
    _fields_ = [
        ("n_points", C.c_int32),
        ("pair_ptr", c_int32_p),
        ("J12", c_float_p),
        ("J21", c_float_p),
        ("H12", c_float_p),
        ("I1", c_float_p),
        ("I2", c_float_p),
        ("pair_from_ref", c_uint8_p),
        ("k_first", c_float_p),
        ("k_init", c_double_p),
        ("ref_uv", c_float_p),
        ("max_iterations", C.c_int32),
        ("corrected_t2", C.c_int32),
    ]


class SfnProblem(C.Structure):
    _fields_ = [
        ("bbs", Bbs),
        ("n_normals", C.c_int32),
        ("uv", c_float_p),
        ("normals", c_float_p),
        ("bending", C.c_double),
        ("mean_depth", C.c_double),
        ("n_eval", C.c_int32),
        ("eval_uv", c_float_p),
        ("ctrl_out", c_double_p),
        ("xyz_out", c_float_p),
    ]


class Sim3Problem(C.Structure):
    _fields_ = [
        ("n_points", C.c_int32),
        ("pts1", c_float_p),
        ("pts2", c_float_p),
        ("rot", C.c_double * 4),
        ("trans", C.c_double * 3),
        ("scale", C.c_double),
        ("chi", C.c_double),
        ("huber", C.c_double),
        ("max_iterations", C.c_int32),
    ]


class Sim3Result(C.Structure):
    _fields_ = [
        ("rot", C.c_double * 4),
        ("trans", C.c_double * 3),
        ("scale", C.c_double),
        ("chi2", C.c_double),
        ("inliers", C.c_int32),
        ("acceptable", C.c_int32),
        ("iterations", C.c_int32 * 2),
    ]


class NewPointsProblem(C.Structure):
    _fields_ = [
        ("n_keypoints", C.c_int32),
        ("rows", C.c_int32),
        ("cols", C.c_int32),
        ("kp_xy", c_float_p),
        ("kp_state", c_uint8_p),
        ("surf_xyz", c_float_p),
        ("T_wc", c_float_p),
    ]


class ProjSearchProblem(C.Structure):
    _fields_ = [
        ("n_last", C.c_int32), ("n_cur", C.c_int32), ("n_levels", C.c_int32),
        ("last_state", c_uint8_p), ("last_has_obs", c_uint8_p), ("last_world_xyz", c_float_p), ("last_desc", c_uint8_p),
        ("last_octave", c_int32_p), ("last_angle", c_float_p),
        ("cur_xy", c_float_p), ("cur_octave", c_int32_p), ("cur_angle", c_float_p), ("cur_desc", c_uint8_p),
        ("cur_uright", c_float_p), ("cur_taken", c_uint8_p), ("scale_factors", c_float_p),
        ("T_cw", C.c_float * 16), ("T_lw", C.c_float * 16),
        ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float), ("mb", C.c_float), ("mbf", C.c_float),
        ("min_x", C.c_float), ("max_x", C.c_float), ("min_y", C.c_float), ("max_y", C.c_float),
        ("grid_width_inv", C.c_float), ("grid_height_inv", C.c_float), ("th", C.c_float),
        ("mono", C.c_int32), ("th_high", C.c_int32), ("check_orientation", C.c_int32),
    ]


class WarpSearchProblem(C.Structure):
    _fields_ = [
        ("bbs", Bbs), ("x", c_double_p), ("n1", C.c_int32), ("n2", C.c_int32),
        ("kp1_norm", c_float_p), ("kp1_state", c_uint8_p), ("kp1_desc", c_uint8_p),
        ("kp2_xy", c_float_p), ("kp2_has_mp", c_uint8_p), ("kp2_desc", c_uint8_p),
        ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
        ("min_x", C.c_float), ("max_x", C.c_float), ("min_y", C.c_float), ("max_y", C.c_float),
        ("grid_width_inv", C.c_float), ("grid_height_inv", C.c_float), ("radius", C.c_float), ("th_low", C.c_int32),
    ]


NORMALS_ARGS = [C.POINTER(NormalsProblem), c_double_p, c_double_p, c_float_p, c_uint8_p, c_int32_p, c_float_p,
                c_uint8_p]
POLY_ARGS = [C.c_int32, c_float_p, c_float_p, c_float_p, c_float_p, c_double_p, c_double_p]

# name -> (restype, argtypes); kept in one table so the "every declared symbol is
# exported" test can iterate it next to the header.
PROTOTYPES = {
    "defslam_template_create": (C.c_int, [C.POINTER(TemplateDesc), C.c_int, C.POINTER(C.c_void_p)]),
    "defslam_template_destroy": (None, [C.c_void_p]),
    "defslam_template_info": (C.c_int, [C.c_void_p, c_int32_p, c_int32_p, c_int32_p, c_int32_p, c_int32_p]),
    "defslam_sft_batch_create": (C.c_int, [C.c_int32, C.POINTER(SftProblem), C.c_int, C.POINTER(C.c_void_p)]),
    "defslam_sft_batch_run": (C.c_int, [C.c_void_p]),
    "defslam_sft_batch_fetch": (C.c_int, [C.c_void_p, C.POINTER(SftResult)]),
    "defslam_sft_batch_info": (
        C.c_int, [C.c_void_p, c_int32_p, c_int32_p, c_int32_p, c_int64_p, c_int64_p, c_double_p]),
    "defslam_sft_batch_destroy": (None, [C.c_void_p]),
    "defslam_mesh_laplacian": (
        C.c_int,
        [C.c_int32, c_double_p, C.c_int32, c_int32_p, C.c_int32, c_int32_p, c_int32_p, c_double_p,
         c_uint8_p, c_double_p, c_int32_p, c_int32_p, c_double_p, c_double_p],
    ),
    "defslam_sft_solve": (C.c_int, [C.POINTER(SftProblem), C.POINTER(SftResult)]),
    "defslam_sft_solve_batched": (C.c_int, [C.c_int32, C.POINTER(SftProblem), C.POINTER(SftResult), C.c_int]),
    "defslam_sft_normal_equations": (C.c_int, [C.POINTER(SftProblem), c_double_p, c_double_p, c_double_p]),
    "defslam_mappoints_recalculate": (
        C.c_int, [C.c_int32, c_double_p, C.c_int32, c_int32_p, c_double_p, c_float_p]),
    "defslam_embed_points": (
        C.c_int,
        [C.c_int32, c_double_p, C.c_int32, c_int32_p, C.c_int32, c_float_p, c_int32_p, c_int32_p, c_float_p],
    ),
    "defslam_bbs_eval": (
        C.c_int, [C.POINTER(Bbs), c_double_p, C.c_int32, c_double_p, c_double_p, C.c_int32, C.c_int32, c_double_p]),
    "defslam_bbs_eval6": (C.c_int, [C.POINTER(Bbs), c_double_p, C.c_int32, c_double_p, c_double_p, c_double_p]),
    "defslam_bbs_coloc": (
        C.c_int, [C.POINTER(Bbs), C.c_int32, c_double_p, c_double_p, C.c_int32, C.c_int32, c_double_p]),
    "defslam_bbs_bending": (C.c_int, [C.POINTER(Bbs), c_double_p]),
    "defslam_schwarp_fit": (C.c_int, [C.POINTER(SchwarpProblem), C.POINTER(DiffProp)]),
    "defslam_schwarp_fit_batched": (
        C.c_int, [C.c_int32, C.POINTER(SchwarpProblem), C.POINTER(DiffProp), C.c_int32]),
    "defslam_schwarp_evaluate": (C.c_int, [C.POINTER(SchwarpProblem), c_double_p, c_double_p]),
    "defslam_schwarp_initial": (C.c_int, [C.POINTER(SchwarpProblem), c_uint8_p, c_double_p]),
    "defslam_normals_batched": (C.c_int, NORMALS_ARGS),
    "defslam_polysolver_coefficients": (C.c_int, POLY_ARGS),
    "defslam_sfn_solve": (C.c_int, [C.POINTER(SfnProblem)]),
    "defslam_sfn_solve_batched": (C.c_int, [C.c_int32, C.POINTER(SfnProblem), c_int32_p, C.c_int32]),
    "defslam_sfn_system": (C.c_int, [C.POINTER(SfnProblem), c_double_p, c_double_p]),
    "defslam_sim3_register_batched": (
        C.c_int, [C.c_int32, C.POINTER(Sim3Problem), C.POINTER(Sim3Result), C.c_int32]),
    "defslam_scale_min_median": (C.c_int, [C.c_int32, c_float_p, c_float_p, C.c_uint64, c_float_p]),
    "defslam_new_map_points": (C.c_int, [C.POINTER(NewPointsProblem), c_uint8_p, c_float_p, c_int32_p]),
    "defslam_search_by_projection": (C.c_int, [C.POINTER(ProjSearchProblem), c_int32_p, c_int32_p]),
    "defslam_search_by_schwarp": (C.c_int, [C.POINTER(WarpSearchProblem), c_int32_p, c_int32_p]),
    "defslam_surface_vertices": (C.c_int, [C.POINTER(Bbs), c_double_p, C.c_int32, C.c_int32, c_float_p]),
    "defslam_version": (C.c_char_p, []),
    "defslam_kernel_launch_count": (C.c_int64, []),
    "defslam_device_count": (C.c_int, []),
    "defslam_last_kernel_ms": (C.c_double, []),
}

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DEFSLAM_LIB", os.path.join(_PKG_DIR, "libdefslam_b200.so"))
_lib = None


def bind(lib):
    """Attach restype/argtypes for every prototype to an opened library."""
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


def load():
    """Open the in-tree CUDA library.  Raises if it has not been built: there is
    deliberately no fallback implementation."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  defslam_b200 has no CPU fallback."
            )
        _lib = bind(C.CDLL(LIB_PATH))
    return _lib


def as_ptr(arr, ctype):
    return arr.ctypes.data_as(C.POINTER(ctype))
