"""ctypes binding of the C-ABI library (include/pbe_b200.h). Fails loudly when the CUDA library is missing."""
from __future__ import annotations

import ctypes
from ctypes import c_char_p, c_float, c_int, c_int64, c_size_t, c_void_p, POINTER
from pathlib import Path

import os

# PBE_B200_LIB: development override (A/B of two builds of the same library in one GPU session)
_LIB_PATH = Path(os.environ.get("PBE_B200_LIB") or Path(__file__).resolve().parent / "lib" / "libpbe_b200.so")
_lib = None


class PbeError(RuntimeError):
    pass


def lib_path() -> Path:
    return _LIB_PATH


def load(build_if_missing: bool = True) -> ctypes.CDLL:
    """Load libpbe_b200.so (building it in-tree with nvcc if absent). No CPU fallback exists."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.environ.get("PBE_B200_LIB"):
        # the in-tree library must match the in-tree sources: a stale .so would be called through the wrong signatures
        from .build import build_native, library_is_current
        if not library_is_current():
            if not build_if_missing:
                raise PbeError(f"{_LIB_PATH} is missing or older than csrc/: run `python -m pbe_b200.build` (nvcc, sm_100a)")
            build_native()
    elif not _LIB_PATH.exists():
        raise PbeError(f"PBE_B200_LIB={_LIB_PATH} does not exist")
    lib = ctypes.CDLL(str(_LIB_PATH))
    _declare(lib)
    _lib = lib
    return lib


def _declare(lib: ctypes.CDLL) -> None:
    lib.pbe_last_error.restype = c_char_p
    lib.pbe_last_error.argtypes = []
    lib.pbe_op_conv_gemm.restype = c_int
    lib.pbe_op_conv_gemm.argtypes = [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int,
                                     c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                     c_void_p]
    for name, sig in _OPTIONAL_SIGS.items():
        if hasattr(lib, name):
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = sig


_f, _i, _p, _i64 = c_float, c_int, c_void_p, c_int64
_OPTIONAL_SIGS: dict = {
    "pbe_op_self_attention": (c_int, [_p, _p, _p, _i, _i, _i, _i, _p]),
    "pbe_debug_gemm_counters": (c_int, [_p]),
    "pbe_set_operand_format": (c_int, [_i]),
    "pbe_get_operand_format": (c_int, []),
    "pbe_debug_set_gemm_stats_out": (None, [_p]),
    "pbe_op_groupnorm": (c_int, [_p, _i, _p, _i, _i, _i, _p, _p, _f, _i, _p, _p, _p, _p]),
    "pbe_op_groupnorm_workspace_bytes": (c_int64, [_i, _i]),
    "pbe_op_layernorm": (c_int, [_p, _p, _p, _p, _i, _i, _f, _p]),
    "pbe_op_upsample2x": (c_int, [_p, _p, _i, _i, _i, _i, _p]),
    "pbe_op_small_linear": (c_int, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _p, _p, _p]),
    "pbe_sampler_step": (c_int, [_p, _p, _f, _i, _i, _p, _p, _p, _p, _f, _f, _f, _f, _p, _f, _p, _p, _p, _i64, _p]),
    "pbe_build_unet_input": (c_int, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p]),
    "pbe_create": (c_int, [_p, _p]),
    "pbe_destroy": (None, [_p]),
    "pbe_load_weight": (c_int, [_p, c_char_p, _p, _p, _i]),
    "pbe_finalize_weights": (c_int, [_p]),
    "pbe_set_context": (c_int, [_p, _p, _i, _p]),
    "pbe_unet_forward": (c_int, [_p, _p, _p, _p, _i, _i, _i, _p]),
    "pbe_unet_forward_cfg_pair": (c_int, [_p, _p, _p, _p, _i, _i, _i, _p]),
    "pbe_set_use_graph": (c_int, [_p, _i]),
    "pbe_launches_per_forward": (c_int, [_p]),
    "pbe_profile_forward": (c_int, [_p, _p, _p, _p, _i, _i, _i, _p, _p, _i]),
    "pbe_op_info": (c_int, [_p, _i, _p, _p, _p, _p]),
    "pbe_vae_create": (c_int, [_p, _p]),
    "pbe_vae_destroy": (None, [_p]),
    "pbe_vae_load_weight": (c_int, [_p, c_char_p, _p, _p, _i]),
    "pbe_vae_finalize_weights": (c_int, [_p]),
    "pbe_vae_decode": (c_int, [_p, _p, _p, _i, _i, _i, _p]),
    "pbe_vae_encode": (c_int, [_p, _p, _p, _i, _i, _i, _p]),
    "pbe_vae_profile": (c_int, [_p, _i, _p, _p, _i, _i, _i, _p, _p, _i]),
    "pbe_vae_op_info": (c_int, [_p, _i, _p, _p, _p]),
    "pbe_vae_launches_per_decode": (c_int, [_p]),
    "pbe_postprocess_u8": (c_int, [_p, _p, _i, _i, _i, _i, _p]),
    "pbe_normalize_u8": (c_int, [_p, _p, _i, _i, _i, _p, _p, _p]),
    "pbe_prepare_inpaint_u8": (c_int, [_p, _p, _i, _i, _i, _i, _p, _p, _p, _p]),
    "pbe_resize_bilinear": (c_int, [_p, _p, _i, _i, _i, _i, _i, _i, _p]),
    "pbe_clip_create": (c_int, [_p, _p]),
    "pbe_clip_destroy": (None, [_p]),
    "pbe_clip_load_weight": (c_int, [_p, c_char_p, _p, _p, _i]),
    "pbe_clip_finalize_weights": (c_int, [_p]),
    "pbe_clip_encode": (c_int, [_p, _p, _p, _i, _p]),
    "pbe_clip_launches_per_encode": (c_int, [_p]),
}


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().pbe_last_error().decode(errors="replace")
        raise PbeError(f"{what} failed (rc={rc}): {msg}")


def ptr(t) -> int | None:
    """Device pointer of a torch tensor (None -> NULL)."""
    if t is None:
        return None
    assert t.is_contiguous(), "C-ABI expects contiguous tensors"
    return t.data_ptr()


def exported_symbols() -> list[str]:
    """Function names declared in include/pbe_b200.h."""
    import re
    hdr = (Path(__file__).resolve().parent.parent / "include" / "pbe_b200.h").read_text()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(pbe_[a-z0-9_]+)\s*\(", hdr)))
