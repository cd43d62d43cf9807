"""ctypes binding of the C ABI declared in include/batrack_ba.h."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# BATRACK_B200_LIB: another build of the same library (the sanitizer verification build, tools/racecheck_verify.sh)
LIB_PATH = os.environ.get("BATRACK_B200_LIB") or os.path.join(_HERE, "libbatrack_ba.so")
_lib = None

STAGES = ("zero", "edge_pass", "track_q", "schur", "solve", "backsub", "pose_retr")   # BA_STAGE_*
LOSS_IDS = {"trivial": 0, "huber": 1, "cauchy": 2}      # compute_kernel_weight, ba.py:81-100
# BA_OPT_* keys of include/batrack_ba.h
OPTIONS = {"solver": 1, "stream": 2, "stream_smem_kb": 3, "schur_tile": 4, "twist_min": 5, "spin_cap": 6,
           "solver_trace": 7, "schur": 8, "schur_acc": 9}
SOLVERS = {"auto": 0, "mma": 1, "window": 2, "dense": 3, "tiles": 4, "diag": 5}


class BaPlanInfo(C.Structure):
    _fields_ = [("n_edges", C.c_int64), ("n_poses", C.c_int32), ("n_patches", C.c_int32),
                ("n_total", C.c_int32), ("n_tracks", C.c_int32), ("n_groups", C.c_int32),
                ("n_chunks", C.c_int32), ("max_degree", C.c_int32), ("max_slots", C.c_int32),
                ("block_bandwidth", C.c_int32), ("perm_identity", C.c_int32), ("banded", C.c_int32),
                ("workspace_bytes", C.c_int64)]


class BaProblem(C.Structure):
    _fields_ = [("poses", C.c_void_p), ("patches", C.c_void_p), ("monodisp", C.c_void_p),
                ("intrinsics", C.c_void_p), ("targets", C.c_void_p), ("weights", C.c_void_p),
                ("lmbda_vec", C.c_void_p), ("lmbda", C.c_float), ("ep", C.c_float), ("alpha", C.c_float),
                ("bounds", C.c_float * 4), ("fixedp", C.c_int32), ("structure_only", C.c_int32),
                ("loss", C.c_int32), ("targets_stride", C.c_int32),
                ("poses_out", C.c_void_p), ("patches_out", C.c_void_p)]


# name -> (restype, argtypes); every symbol include/batrack_ba.h declares
_P, _I64, _I32 = C.c_void_p, C.c_int64, C.c_int32
SYMBOLS = {
    "ba_plan_create": (C.c_int, [_P, _P, _P, _I64, _I32, _I32, _P, C.POINTER(_P)]),
    "ba_plan_destroy": (None, [_P]),
    "ba_plan_create_capacity": (C.c_int, [_I64, _I32, _I32, _I32, _I64, _I32, _I32, _P, C.POINTER(_P)]),
    "ba_plan_update": (C.c_int, [_P, _P, _P, _P, _I64, _P, _P]),
    "ba_plan_finalize": (C.c_int, [_P]),
    "ba_trajectory": (C.c_int, [_P, _P, _P, _P, _I32, _P, _P, _P, _P]),
    "ba_graph_create": (C.c_int, [_I64, _P, C.POINTER(_P)]),
    "ba_graph_destroy": (None, [_P]),
    "ba_graph_arrays": (C.c_int, [_P, C.POINTER(_P), C.POINTER(_P), C.POINTER(_P), C.POINTER(_P), C.POINTER(_P), C.POINTER(_P),
                                  C.POINTER(_P), C.POINTER(_I64), C.POINTER(_I64)]),
    "ba_graph_append": (C.c_int, [_P, _P, _P, _I64, _P, _P, _P, _P, _P]),
    "ba_graph_remove": (C.c_int, [_P, _I32, _I64, _I64, _P, _P, _P]),
    "ba_graph_count": (C.c_int, [_P, C.POINTER(_I64), _P]),
    "ba_graph_tighten": (C.c_int, [_P, _I64]),
    "ba_plan_info": (C.c_int, [_P, C.POINTER(BaPlanInfo)]),
    "ba_plan_set_layout": (C.c_int, [_P, _I32, _I32]),
    "ba_plan_set_option": (C.c_int, [_P, _I32, _I32]),
    "ba_plan_get_option": (C.c_int, [_P, _I32, C.POINTER(_I32)]),
    "ba_plan_read_trace": (C.c_int, [_P, _P, _I64, _P]),
    "ba_plan_tracks": (C.c_int, [_P, _P, _P]),
    "ba_step": (C.c_int, [_P, C.POINTER(BaProblem), _P]),
    "ba_update": (C.c_int, [_P, C.POINTER(BaProblem), _P, _I32, _P]),
    "ba_assemble": (C.c_int, [_P, C.POINTER(BaProblem), _P]),
    "ba_plan_reduced_system": (C.c_int, [_P, C.POINTER(_P), C.POINTER(_I64)]),
    "ba_solve_update": (C.c_int, [_P, C.POINTER(BaProblem), _P]),
    "ba_plan_debug_dense": (C.c_int, [_P, _I32, _P, _P, _P, _P, _P, _P, _P]),
    "ba_plan_status_ptr": (C.c_int, [_P, C.POINTER(_P)]),
    "ba_plan_enable_timing": (C.c_int, [_P, C.c_int]),
    "ba_plan_last_timing": (C.c_int, [_P, C.POINTER(C.c_float)]),
    "ba_step_host": (C.c_int, [_P, C.POINTER(BaProblem), _P]),
    "ba_step_host_async": (C.c_int, [_P, C.POINTER(BaProblem), _P]),
    "ba_prefetch_host_async": (C.c_int, [_P, C.POINTER(BaProblem), _P]),
    "ba_stage_host_async": (C.c_int, [_P, C.POINTER(BaProblem), C.POINTER(BaProblem), _P]),
    "ba_unstage_host_async": (C.c_int, [_P, C.POINTER(BaProblem), C.POINTER(BaProblem), _P]),
    "ba_host_sync": (C.c_int, [_P, _P, C.c_int]),
    "ba_reproject": (C.c_int, [_P, _P, _P, _P, _P, _P, _I64, _I32, _I32, _I32, _P, _P, _P]),
    "ba_transform": (C.c_int, [_P, _P, _P, _P, _P, _P, _I64, _I32, _I32, _I32, _I32, _P, _P, _P, _P, _P, _P]),
    "ba_point_cloud": (C.c_int, [_P, _P, _P, _P, _I64, _I32, _P, _P]),
    "ba_back_proj": (C.c_int, [_P, _P, _P, _P, _I32, _I64, _P, _P]),
    "ba_proj_to_frames": (C.c_int, [_P, _P, _P, _I32, _I32, _I64, _P, _P]),
    "se3_expm": (C.c_int, [_P, _P, _I64, _P]),
    "se3_logm": (C.c_int, [_P, _P, _I64, _P]),
    "se3_inv": (C.c_int, [_P, _P, _I64, _P]),
    "se3_mul": (C.c_int, [_P, _P, _P, _I64, _P]),
    "se3_adj": (C.c_int, [_P, _P, _P, _I64, _P]),
    "se3_adjT": (C.c_int, [_P, _P, _P, _I64, _P]),
    "se3_act": (C.c_int, [_P, _P, _P, _I64, _P]),
    "se3_act4": (C.c_int, [_P, _P, _P, _I64, _P]),
    "se3_as_matrix": (C.c_int, [_P, _P, _I64, _P]),
    "se3d_expm": (C.c_int, [_P, _P, _I64, _P]),
    "se3d_logm": (C.c_int, [_P, _P, _I64, _P]),
    "se3d_inv": (C.c_int, [_P, _P, _I64, _P]),
    "se3d_mul": (C.c_int, [_P, _P, _P, _I64, _P]),
    "se3d_adj": (C.c_int, [_P, _P, _P, _I64, _P]),
    "se3d_adjT": (C.c_int, [_P, _P, _P, _I64, _P]),
    "se3d_act": (C.c_int, [_P, _P, _P, _I64, _P]),
    "se3d_act4": (C.c_int, [_P, _P, _P, _I64, _P]),
    "se3d_as_matrix": (C.c_int, [_P, _P, _I64, _P]),
    "ba_error_string": (C.c_char_p, [C.c_int]),
    "ba_last_cuda_error": (C.c_char_p, []),
    "ba_version": (C.c_int, []),
    "ba_launch_count": (C.c_int64, []),
}


def load_library():
    """dlopen libbatrack_ba.so and bind every declared symbol. Raises if the library is missing —
    there is no fallback path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} not found: build it with `python __graft_entry__.py` "
                           "(batrack_b200 has no CPU / eager fallback)")
    lib_ = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib_, name)          # AttributeError if the ABI and the header disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib_
    return _lib


def lib():
    return load_library()


def launch_count():
    return int(lib().ba_launch_count())


def check(rc, what=""):
    if rc != 0:
        L = lib()
        msg = L.ba_error_string(rc).decode()
        if rc == -1:
            msg += ": " + L.ba_last_cuda_error().decode()
        raise RuntimeError(f"batrack_b200 {what} failed ({rc}): {msg}")


def ptr(t):
    """device pointer of a torch tensor (None -> NULL)"""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr(device):
    import torch
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda_float(name, t, contiguous=True):
    """float32 or float64 CUDA tensor (the SE3 group ops dispatch both, like lietorch/include/dispatch.h:37-45)"""
    import torch
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name}: expected a torch.Tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise RuntimeError(f"{name}: batrack_b200 runs on CUDA tensors only (got {t.device}); there is no CPU fallback")
    if t.dtype not in (torch.float32, torch.float64):
        raise TypeError(f"{name}: float32 or float64 required (got {t.dtype})")
    if contiguous and not t.is_contiguous():
        raise RuntimeError(f"{name}: input must be contiguous")        # lietorch.cpp:7 CHECK_CONTIGUOUS
    return t


def require_cuda_f32(name, t, contiguous=True):
    import torch
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name}: expected a torch.Tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise RuntimeError(f"{name}: batrack_b200 runs on CUDA tensors only (got {t.device}); "
                           "there is no CPU fallback")
    if t.dtype != torch.float32:
        raise TypeError(f"{name}: float32 required (got {t.dtype})")
    if contiguous and not t.is_contiguous():
        raise RuntimeError(f"{name}: input must be contiguous")        # lietorch.cpp:7 CHECK_CONTIGUOUS
    return t
