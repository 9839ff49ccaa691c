"""ctypes binding of libdedflow_b200.so (the C ABI declared in include/dedflow_b200.h).

There is NO fallback: if the shared object is missing, or a call fails, this module raises.  Device memory is
owned by the caller (torch tensors in the tests / bench); only raw pointers cross the boundary.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

PKG = Path(__file__).resolve().parent
# DFB_LIB: load another build of the same library (A/B measurements of kernel variants); there is still no fallback
LIB_PATH = Path(os.environ["DFB_LIB"]) if os.environ.get("DFB_LIB") else PKG / "libdedflow_b200.so"

vp = C.c_void_p
ci = C.c_int


class DfbError(RuntimeError):
    pass


_SIGS = {
    # name: (restype, argtypes)
    "dfb_last_error": (C.c_char_p, []),
    "dfb_version": (ci, []),
    "dfb_launch_count": (C.c_longlong, []),
    "dfb_set_option": (ci, [C.c_char_p, C.c_char_p]),
    "dfb_gmres_profile": (ci, [vp, C.c_char_p, ci]),
    "dfb_pc2_create": (ci, [C.POINTER(vp), ci, vp, vp, vp, ci, ci, vp]),
    "dfb_pc2_info": (ci, [vp, C.POINTER(ci), C.POINTER(ci)]),
    "dfb_pc2_setup": (ci, [vp, vp, vp, vp, vp, vp]),
    "dfb_pc2_apply": (ci, [vp, vp, vp, vp, vp]),
    "dfb_pc2_destroy": (None, [vp]),
    "dfb_gmres_set_pc2": (ci, [vp, vp]),
    "dfb_pattern_rows": (ci, [ci, ci, vp, vp, C.POINTER(ci), vp]),
    "dfb_pattern_cols": (ci, [ci, ci, vp, vp, vp, vp]),
    "dfb_pattern_expand": (ci, [ci, vp, vp, ci, ci, vp, vp, vp]),
    "dfb_color_weights": (ci, [ci, C.c_ulonglong, vp, vp]),
    "dfb_color_jpl": (ci, [ci, ci, vp, vp, ci, vp, C.POINTER(ci), vp]),
    "dfb_color_batches": (ci, [ci, vp, ci, vp, vp, vp]),
    "dfb_plan_create": (ci, [C.POINTER(vp), ci, ci, vp, vp, vp, ci, vp, vp, vp]),
    "dfb_plan_destroy": (None, [vp]),
    "dfb_plan_bytes": (C.c_size_t, [vp]),
    "dfb_assemble_tet": (ci, [vp, vp, vp, vp, vp, vp, vp, vp, vp, ci, ci, vp]),
    "dfb_assemble_face": (ci, [vp, ci, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    "dfb_dirichlet_vec": (ci, [ci, vp, ci, vp, vp, vp]),
    "dfb_dirichlet_mat": (ci, [ci, vp, ci, vp, ci, vp, vp, vp, vp, vp]),
    "dfb_spmv_fs": (ci, [ci, vp, vp, vp, vp, vp, vp, C.c_double, vp, C.c_double, vp, vp]),
    "dfb_pc_setup": (ci, [ci, vp, vp, vp, vp, vp, vp, vp]),
    "dfb_pc_apply": (ci, [ci, vp, vp, vp, vp, vp]),
    "dfb_gmres_create": (ci, [C.POINTER(vp), ci, ci]),
    "dfb_gmres_destroy": (None, [vp]),
    "dfb_gmres_bytes": (C.c_size_t, [vp]),
    "dfb_gmres_set_parallel": (ci, [vp, vp]),
    "dfb_plan_set_rows": (ci, [vp, ci]),
    "dfb_comm_unique_id": (ci, [vp]),
    "dfb_comm_create": (ci, [C.POINTER(vp), ci, ci, vp]),
    "dfb_comm_destroy": (None, [vp]),
    "dfb_comm_set_halo": (ci, [vp, ci, ci, vp, vp, vp, vp, vp]),
    "dfb_comm_allreduce": (ci, [vp, ci, vp, vp]),
    "dfb_comm_halo_begin": (ci, [vp, vp, vp]),
    "dfb_comm_halo_begin_aos": (ci, [vp, vp, vp]),
    "dfb_comm_halo_end": (ci, [vp, vp, vp]),
    "dfb_comm_halo": (ci, [vp, vp, vp]),
    "dfb_genalpha_stage": (ci, [ci, vp, vp, vp, vp, vp, vp]),
    "dfb_newton_update": (ci, [ci, vp, vp, vp]),
    "dfb_genalpha_predict": (ci, [ci, vp, vp]),
    "dfb_genalpha_correct": (ci, [ci, vp, vp, vp, vp]),
    "dfb_block_sumsq": (ci, [ci, ci, vp, vp, vp]),
    "dfb_block_norms": (ci, [ci, vp, vp, vp]),
    "dfb_comm_p2p_alloc": (ci, [vp, vp]),
    "dfb_comm_p2p_connect": (ci, [vp, vp, vp, vp]),
    "dfb_comm_p2p_view": (vp, [vp]),
    "dfb_gmres_solve_pc": (ci, [vp, ci, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, C.c_double, C.c_double, C.POINTER(ci), vp, vp]),
    "dfb_gmres_solve": (ci, [vp, ci, vp, vp, vp, vp, vp, vp, vp, vp, C.c_double, C.c_double, C.POINTER(ci), vp, vp]),
}

_lib = None


def exported_symbols():
    """Every entry point include/dedflow_b200.h declares (used by the CPU-side ABI test)."""
    return sorted(_SIGS)


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise DfbError(f"{LIB_PATH} is missing: build it with `python -m dedflow_b200._build` "
                       "(there is no CPU fallback)")
    lib = C.CDLL(str(LIB_PATH), mode=C.RTLD_LOCAL)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)          # raises AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int, what: str = ""):
    if status != 0:
        msg = load().dfb_last_error().decode(errors="replace")
        raise DfbError(f"{what} failed with status {status}: {msg}")


def set_option(key: str, value) -> None:
    """Switch a library variant (include/dedflow_b200.h dfb_set_option); keys are the DFB_* environment names."""
    check(load().dfb_set_option(key.encode(), str(value).encode()), f"dfb_set_option({key})")


def solve_profile(ws) -> dict:
    """{kernel: {"launches": n, "total_ms": t, "avg_us": a}} of the last solve of a dfb_gmres workspace (DFB_PROFILE != 0)."""
    buf = C.create_string_buffer(4096)
    check(load().dfb_gmres_profile(ws, buf, 4096), "dfb_gmres_profile")
    out = {}
    for item in buf.value.decode().split(";"):
        if item:
            name, n, ms = item.rsplit(":", 2)
            out[name] = {"launches": int(n), "total_ms": float(ms), "avg_us": 1e3 * float(ms) / max(1, int(n))}
    return out


def launch_count() -> int:
    return int(load().dfb_launch_count())
