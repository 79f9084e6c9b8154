"""ctypes binding of libadmm_b200.so (the C ABI declared in include/admm_b200.h).

There is NO CPU fallback: if the shared library is missing or no CUDA device is present the
product path raises.  Host-only entry points (grid loading / refinement) work without a GPU.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libadmm_b200.so")


class AdmmB200Error(RuntimeError):
    pass


class GmgDesc(C.Structure):
    _fields_ = [("smoother", C.c_int), ("pre_smooth", C.c_int), ("post_smooth", C.c_int), ("base_level", C.c_int),
                ("rap", C.c_int), ("max_iterations", C.c_int), ("abs_tol", C.c_double), ("red_tol", C.c_double),
                ("verbose", C.c_int), ("cheb_ratio", C.c_double), ("jacobi_damp", C.c_double)]


_P = C.c_void_p
_PP = C.POINTER(C.c_void_p)
_D = C.c_double
_DP = C.POINTER(C.c_double)
_I = C.c_int
_IP = C.POINTER(C.c_int)
_I32P = C.POINTER(C.c_int32)
_I64P = C.POINTER(C.c_int64)
_S = C.c_char_p

# name -> argtypes; every function returns int status except ab_last_error / ab_version
PROTOTYPES = {
    "ab_context_create": [_I, _P, _PP],
    "ab_context_destroy": [_P],
    "ab_context_synchronize": [_P],
    "ab_context_launch_count": [_P, _I64P],
    "ab_context_set_tuning": [_P, _S, _I],
    "ab_context_init_comm": [_P, _I, _I, _P],
    "ab_nccl_unique_id": [_P],
    "ab_context_allreduce_host": [_P, _DP, _I, _I],
    "ab_domain_load_ugx": [_P, _S, _PP],
    "ab_domain_create": [_P, _I, _I, _DP, _I, _I32P, _I, C.POINTER(_S), _I32P, _I32P, _I, _I32P, _I32P, _I, _I32P, _I32P, _PP],
    "ab_domain_destroy": [_P],
    "ab_domain_refine": [_P, _I],
    "ab_domain_set_interface": [_P, _I, _I, _I32P, _I32P, _I32P, C.POINTER(C.c_ubyte)],
    "ab_domain_set_gather": [_P, _I, _P, _I32P, _I32P, _I64P, _I32P],
    "ab_domain_level_pattern": [_P, _I, _I64P, _I32P, _I32P],
    "ab_domain_level_incidence": [_P, _I, _I32P, _I32P],
    "ab_domain_set_block_interface": [_P, _I, _I, _I32P, _I32P, _I32P, _I, _I32P, _I32P, _I32P],
    "ab_domain_p2p_export": [_P, _P, _I64P, _I32P],
    "ab_domain_p2p_connect": [_P, _P, _I64P, _I64P],
    "ab_domain_p2p_status": [_P, _IP, _IP],
    "ab_domain_num_levels": [_P, _IP],
    "ab_domain_level_info": [_P, _I, _IP, _IP, _IP, _IP, _IP],
    "ab_domain_get_level": [_P, _I, _DP, _I32P, _I32P, _I32P, _I32P],
    "ab_domain_subset_index": [_P, _S, _IP],
    "ab_domain_subset_name": [_P, _I, C.c_char_p, _I],
    "ab_domain_special_info": [_P, _I, _IP, _IP, _IP],
    "ab_domain_get_special": [_P, _I, _I32P, _I32P, _I32P, _I32P, _I32P],
    "ab_transform_domain_by_displacement": [_P, _P],
    "ab_space_create": [_P, _I, _I, _PP],
    "ab_space_destroy": [_P],
    "ab_space_num_dofs": [_P, _I64P],
    "ab_vector_create": [_P, _PP],
    "ab_vector_destroy": [_P],
    "ab_vector_set": [_P, _D],
    "ab_vector_upload": [_P, _DP, _I],
    "ab_vector_download": [_P, _DP],
    "ab_vector_device_ptr": [_P, _PP, _I64P],
    "ab_vector_storage": [_P, _IP],
    "ab_vector_change_storage": [_P, _I],
    "ab_vec_scale_assign": [_P, _D, _P],
    "ab_vec_scale_add2": [_P, _D, _P, _D, _P],
    "ab_vec_prod": [_P, _P, _DP],
    "ab_vec_prod_multi": [_I, _PP, _P, _DP],
    "ab_vec_norm": [_P, _DP],
    "ab_l2norm": [_P, _I, _DP],
    "ab_l2norm_all": [_P, _DP],
    "ab_elemdisc_create": [_P, _I, _PP],
    "ab_elemdisc_destroy": [_P],
    "ab_elemdisc_set_param": [_P, _I, _D],
    "ab_elemdisc_get_param": [_P, _I, _DP],
    "ab_elemdisc_bind": [_P, _I, _P],
    "ab_domaindisc_create": [_P, _PP],
    "ab_domaindisc_destroy": [_P],
    "ab_domaindisc_add_elemdisc": [_P, _P],
    "ab_domaindisc_add_dirichlet": [_P, _S, _I, _D],
    "ab_domaindisc_assemble_jacobian": [_P, _P, _P],
    "ab_domaindisc_assemble_defect": [_P, _P, _P],
    "ab_domaindisc_adjust_solution": [_P, _P],
    "ab_operator_create": [_P, _PP],
    "ab_operator_destroy": [_P],
    "ab_operator_apply": [_P, _P, _P],
    "ab_operator_info": [_P, _IP, _I64P, _I64P],
    "ab_operator_download": [_P, _I32P, _I32P, _DP],
    "ab_solver_create_bicgstab_gmg": [_P, C.POINTER(GmgDesc), _PP],
    "ab_solver_create_cg_jacobi": [_P, _D, _I, _D, _D, _I, _PP],
    "ab_solver_destroy": [_P],
    "ab_solver_init": [_P, _P, _P],
    "ab_solver_apply": [_P, _P, _P, _IP],
    "ab_solver_apply_return_defect": [_P, _P, _P, _IP],
    "ab_solver_step": [_P, _IP],
    "ab_solver_last_defect": [_P, _DP],
    "ab_solver_vcycle": [_P, _P, _P],
    "ab_solver_level_info": [_P, _I, _I64P, _I64P],
    "ab_project_frobenius": [_P, _P, _D],
    "ab_project_spectral": [_P, _P, _D],
    "ab_max_frobenius_norm": [_P, _DP],
    "ab_max_spectral_norm": [_P, _DP],
    "ab_volume_defect": [_P, _D, _DP],
    "ab_barycenter_defect": [_P, _DP],
    "ab_set_zero_away_from_subset": [_P, _S],
}

_lib = None


def load():
    """Load the shared library (raises AdmmB200Error when it was not built: no silent fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise AdmmB200Error("libadmm_b200.so is missing at %s -- run `python -c 'import __graft_entry__ as g; g.build()'`; "
                            "there is no CPU fallback" % LIB_PATH)
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, args in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = C.c_int
    lib.ab_last_error.restype = C.c_char_p
    lib.ab_last_error.argtypes = []
    lib.ab_version.restype = C.c_int
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise AdmmB200Error("libadmm_b200 error %d: %s" % (rc, load().ab_last_error().decode()))


def call(name, *args):
    check(getattr(load(), name)(*args))
