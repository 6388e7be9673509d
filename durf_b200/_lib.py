"""ctypes binding of libdurf_b200.so (the C ABI declared in include/durf_b200.h).

The product path has no CPU fallback: if the shared library is missing, or a call is made without a CUDA
device, this module raises.  PyTorch is used only for device memory and streams (tensor.data_ptr(),
torch.cuda.current_stream().cuda_stream).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# DURF_B200_LIB: another build of the library (e.g. one made with `make EXTRA=-DDURF_TRACE=1`)
LIB_PATH = os.environ.get("DURF_B200_LIB") or os.path.join(_HERE, "libdurf_b200.so")
CSRC = os.path.join(_HERE, "csrc")

OK = 0
PREC_FP32 = 0
PREC_BF16 = 1

RM_SAMPLE = 1 << 0
RM_RANDOMIZED = 1 << 1
RM_CONTRACT = 1 << 2
RM_WEIGHTED = 1 << 3
RM_CYLINDER = 1 << 4
RM_NO_INTEGRATE = 1 << 5
RM_OUT_BF16_TILE = 1 << 6
RM_NO_TVALS_OUT = 1 << 7
RM_MULT_IS_NHIT = 1 << 8

LP_STRIDE = 8
LP_NAMES = ("rgb", "depth", "near", "empty", "sky", "distr")


class DurfError(RuntimeError):
    pass


class Camera(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("focal", C.c_float), ("c2w", C.c_float * 12),
                ("near", C.c_float), ("far", C.c_float), ("use_principal_point", C.c_int32), ("cx", C.c_float), ("cy", C.c_float)]


class MlpTopology(C.Structure):
    _fields_ = [("in_dim", C.c_int32), ("width", C.c_int32), ("depth", C.c_int32), ("skip", C.c_int32),
                ("cond_dim", C.c_int32), ("cond_width", C.c_int32)]

    def key(self):
        return (self.in_dim, self.width, self.depth, self.skip, self.cond_dim, self.cond_width)


class RaymarchArgs(C.Structure):
    _fields_ = [("B", C.c_int32), ("N", C.c_int32), ("min_deg", C.c_int32), ("max_deg", C.c_int32),
                ("flags", C.c_uint32), ("alpha", C.c_float),
                ("origins", C.c_void_p), ("dirs", C.c_void_p), ("radii", C.c_void_p), ("near", C.c_void_p),
                ("far", C.c_void_p), ("t_rand", C.c_void_p), ("t_vals", C.c_void_p), ("ray_mult", C.c_void_p),
                ("ray_index", C.c_void_p), ("count", C.c_void_p), ("features", C.c_void_p), ("means", C.c_void_p),
                ("cov_diag", C.c_void_p), ("alpha_dev", C.c_void_p)]


class MlpArgs(C.Structure):
    _fields_ = [("topo", MlpTopology), ("precision", C.c_int32), ("M", C.c_int32), ("N", C.c_int32),
                ("features", C.c_void_p), ("cond", C.c_void_p), ("params", C.c_void_p), ("packed", C.c_void_p),
                ("ray_index", C.c_void_p), ("count", C.c_void_p), ("accumulate", C.c_int32),
                ("raw_rgb", C.c_void_p), ("raw_density", C.c_void_p), ("saved", C.c_void_p),
                ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t), ("fused_raymarch", C.c_void_p),
                ("saved_tile_offset", C.c_int32), ("saved_total_tiles", C.c_int32)]


class CompositeArgs(C.Structure):
    _fields_ = [("B", C.c_int32), ("N", C.c_int32), ("white_bkgd", C.c_int32), ("rand_bkgd", C.c_int32),
                ("activated", C.c_int32), ("density_bias", C.c_float),
                ("raw_rgb", C.c_void_p), ("raw_density", C.c_void_p), ("t_vals", C.c_void_p), ("dirs", C.c_void_p),
                ("comp_rgb", C.c_void_p), ("depth", C.c_void_p), ("acc", C.c_void_p), ("weights", C.c_void_p),
                ("t_mids", C.c_void_p), ("t_dists", C.c_void_p)]


class LossArgs(C.Structure):
    _fields_ = [("B", C.c_int32), ("N", C.c_int32), ("level", C.c_int32), ("num_levels", C.c_int32),
                ("eps", C.c_float), ("coarse_loss_mult", C.c_float), ("box_loss_mult", C.c_float),
                ("depth_loss_mult", C.c_float), ("near_loss_mult", C.c_float), ("empty_loss_mult", C.c_float),
                ("sky_loss_mult", C.c_float), ("distortion_mult", C.c_float),
                ("comp_rgb", C.c_void_p), ("depth", C.c_void_p), ("weights", C.c_void_p), ("t_vals", C.c_void_p),
                ("pixels", C.c_void_p), ("depth_gt", C.c_void_p), ("sky", C.c_void_p), ("lossmult", C.c_void_p),
                ("dyn_mask", C.c_void_p), ("zo", C.c_void_p), ("depth_mask", C.c_void_p), ("partials", C.c_void_p),
                ("d_comp_rgb", C.c_void_p), ("d_depth", C.c_void_p), ("d_weights", C.c_void_p),
                ("reduce_ws", C.c_void_p), ("eps_dev", C.c_void_p)]


class LossFinalizeArgs(C.Structure):
    _fields_ = [("num_levels", C.c_int32), ("coarse_loss_mult", C.c_float), ("depth_loss_mult", C.c_float),
                ("near_loss_mult", C.c_float), ("empty_loss_mult", C.c_float), ("sky_loss_mult", C.c_float),
                ("tv_loss_mult", C.c_float), ("distortion_mult", C.c_float),
                ("partials", C.c_void_p), ("norms", C.c_void_p), ("tv", C.c_void_p), ("weight_l2", C.c_void_p),
                ("stats", C.c_void_p)]


LS_STRIDE = 8

# name -> (restype, argtypes); mirrors include/durf_b200.h one to one
_vp, _i32, _i64, _f = C.c_void_p, C.c_int32, C.c_int64, C.c_float
SIGNATURES = {
    "durf_version": (C.c_char_p, []),
    "durf_last_error": (C.c_char_p, []),
    "durf_launch_count": (_i64, []),
    "durf_reset_launch_count": (None, []),
    "durf_mlp_param_count": (_i64, [C.POINTER(MlpTopology)]),
    "durf_mlp_param_offset": (_i64, [C.POINTER(MlpTopology), _i32, C.POINTER(_i32), C.POINTER(_i32)]),
    "durf_generate_rays": (C.c_int, [_vp, C.POINTER(Camera), _i32, _i32] + [_vp] * 7),
    "durf_aa2matrix_fwd": (C.c_int, [_vp, _i32, _vp, _vp]),
    "durf_obb_frontend_fwd": (C.c_int, [_vp, _i32, _i32] + [_vp] * 13),
    "durf_world2object_fwd": (C.c_int, [_vp, _i32, _i32, _vp, _vp, _vp, _i32, _vp, _i32, _vp, _vp]),
    "durf_ray_box_intersection_fwd": (C.c_int, [_vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "durf_obb_frontend_bwd": (C.c_int, [_vp, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _vp]),
    "durf_compact_hits": (C.c_int, [_vp, _i32, _i32, _i32, _vp, _vp, _vp]),
    "durf_compact_hits_all": (C.c_int, [_vp, _i32, _i32, _vp, _vp, _vp]),
    "durf_mlp_merge_raw": (C.c_int, [_vp, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "durf_raymarch_fwd": (C.c_int, [_vp, C.POINTER(RaymarchArgs)]),
    "durf_raymarch_bwd": (C.c_int, [_vp, C.POINTER(RaymarchArgs), _vp, _vp, _vp]),
    "durf_viewdir_enc_fwd": (C.c_int, [_vp, _i32, _i32, _vp, _vp]),
    "durf_mlp_packed_bytes": (_i64, [C.POINTER(MlpTopology)]),
    "durf_mlp_pack_weights": (C.c_int, [_vp, C.POINTER(MlpTopology), _vp, _vp]),
    "durf_mlp_pack_weights_multi": (C.c_int, [_vp, _i32, C.POINTER(MlpTopology), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "durf_mlp_workspace_bytes": (C.c_size_t, [C.POINTER(MlpTopology), _i32, _i32, _i32, _i32]),
    "durf_mlp_saved_bytes": (C.c_size_t, [C.POINTER(MlpTopology), _i32, _i32, _i32]),
    "durf_mlp_fwd": (C.c_int, [_vp, C.POINTER(MlpArgs)]),
    "durf_mlp_bwd": (C.c_int, [_vp, C.POINTER(MlpArgs), _vp, _vp, _vp, _vp]),
    "durf_mlp_bwd_flags_bytes": (C.c_size_t, [C.POINTER(MlpTopology), _i32]),
    "durf_mlp_bwd_data": (C.c_int, [_vp, C.POINTER(MlpArgs), _vp, _vp, _vp, _vp, _i32]),
    "durf_mlp_bwd_weights": (C.c_int, [_vp, C.POINTER(MlpArgs), _vp, _vp, _vp, _vp, _i32]),
    "durf_composite_fwd": (C.c_int, [_vp, C.POINTER(CompositeArgs)]),
    "durf_composite_bwd": (C.c_int, [_vp, C.POINTER(CompositeArgs), _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "durf_resample_fwd": (C.c_int, [_vp, _i32, _i32, _vp, _vp, _vp, _f, _i32, _i32, _vp]),
    "durf_losses_prepare": (C.c_int, [_vp, C.POINTER(LossArgs), _vp]),
    "durf_losses_fwd_bwd": (C.c_int, [_vp, C.POINTER(LossArgs), _vp]),
    "durf_grad_sanitize": (C.c_int, [_vp, _i64, _vp, _f, _f, _vp]),
    "durf_adam_step": (C.c_int, [_vp, _i64, _vp, _vp, _vp, _vp, _vp, _f, _f, C.c_double, C.c_double, C.c_double, _i32]),
    "durf_adam_step_dev": (C.c_int, [_vp, _i64, _vp, _vp, _vp, _vp, _vp, _f, _vp, _vp, _i32, C.c_double, C.c_double, C.c_double]),
    "durf_losses_reduce_ws_floats": (_i64, [_i32]),
    "durf_losses_finalize": (C.c_int, [_vp, C.POINTER(LossFinalizeArgs)]),
}

_lib: Optional[C.CDLL] = None


def build(verbose: bool = False) -> str:
    """Compile durf_b200/csrc/*.cu for sm_100a into durf_b200/libdurf_b200.so (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-j8", "-C", CSRC], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:])
        print(r.stderr[-4000:])
    if r.returncode != 0:
        raise DurfError("building libdurf_b200.so failed")
    return LIB_PATH


def load() -> C.CDLL:
    """Load the shared library and bind every declared symbol.  Raises if it is missing: there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DurfError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(the CUDA extension is mandatory, there is no CPU path)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != OK:
        msg = load().durf_last_error().decode()
        raise DurfError(f"{what or 'durf call'} failed with code {rc}: {msg}")


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    """Device pointer of a contiguous CUDA tensor (None passes through as NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise DurfError("durf_b200 operates on CUDA tensors only (got a CPU tensor); there is no CPU path")
    if not t.is_contiguous():
        raise DurfError("durf_b200 needs contiguous tensors")
    return t.data_ptr()


def f32(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()
