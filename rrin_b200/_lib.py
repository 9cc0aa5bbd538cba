"""ctypes binding of ``librrin_b200.so`` (C-ABI declared in ``include/rrin_b200.h``).

Fails loudly when the library is missing: there is no Python/torch fallback for any kernel.
"""
from __future__ import annotations

import ctypes as C
import os
import re

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RRIN_LIB") or os.path.join(PKG, "librrin_b200.so")      # RRIN_LIB: an alternative build (A/B timing)
HEADER = os.path.join(os.path.dirname(PKG), "include", "rrin_b200.h")

_lib = None

vp, ci, cs = C.c_void_p, C.c_int, C.c_size_t
_SIGNATURES = {
    "rrin_version": (ci, []),
    "rrin_last_error": (C.c_char_p, []),
    "rrin_num_convs": (ci, []),
    "rrin_conv_info": (ci, [ci, C.c_char_p, ci] + [C.POINTER(ci)] * 5),
    "rrin_packed_weights_bytes": (cs, []),
    "rrin_pack_conv": (ci, [ci, vp, vp, vp, vp]),
    "rrin_pack_conv_ex": (ci, [ci, vp, vp, vp, ci, vp]),
    "rrin_engine_create": (ci, [ci, ci, ci, ci, C.POINTER(vp)]),
    "rrin_engine_create_ex": (ci, [ci, ci, ci, ci, ci, C.POINTER(vp)]),
    "rrin_engine_destroy": (None, [vp]),
    "rrin_engine_workspace_bytes": (cs, [vp]),
    "rrin_engine_num_launches": (ci, [vp]),
    "rrin_engine_forward": (ci, [vp] * 8),
    "rrin_engine_forward_graph": (ci, [vp] * 8),
    "rrin_engine_graph_stats": (ci, [vp] + [C.POINTER(ci)] * 3),
    "rrin_engine_tap": (ci, [vp, vp, ci, vp, vp]),
    "rrin_engine_launch_info": (ci, [vp, ci, C.c_char_p, ci, C.c_char_p, ci, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "rrin_engine_forward_profiled": (ci, [vp] * 8 + [C.POINTER(C.c_float)]),
    "rrin_conv_config_info": (ci, [ci] + [C.POINTER(ci)] * 4),
    "rrin_conv_packed_weight_bytes": (cs, [ci, ci, ci, ci]),
    "rrin_conv_packed_bias_count": (ci, [ci, ci]),
    "rrin_pack_conv_raw": (ci, [ci, vp, vp, ci, ci, ci, ci, vp, vp, vp]),
    "rrin_conv3x3": (ci, [vp, vp, ci, ci, ci, ci, ci, ci, ci, ci, ci, vp, vp, vp, ci, ci, ci, ci, ci, vp, vp]),
    "rrin_pack_conv_raw_ex": (ci, [ci, vp, vp, ci, ci, ci, ci, vp, vp, ci, vp]),
    "rrin_conv3x3_ex": (ci, [vp, vp, ci, ci, ci, ci, ci, ci, ci, ci, ci, vp, vp, vp, ci, ci, ci, ci, ci, vp, ci, ci, vp]),
    "rrin_pack_pair": (ci, [vp, vp, ci, ci, ci, vp, vp]),
    "rrin_flow_tscale_pack": (ci, [vp, vp, vp, vp, ci, ci, ci, ci, vp, vp]),
    "rrin_warp_pack": (ci, [vp, vp, vp, vp, vp, ci, ci, ci, ci, vp, vp, vp]),
    "rrin_blend_pack": (ci, [vp, vp, vp, vp, vp, ci, ci, ci, ci, vp, vp, vp]),
    "rrin_warp": (ci, [vp, vp, ci, ci, ci, ci, vp, vp]),
    "rrin_residue_clamp": (ci, [vp, vp, ci, ci, ci, vp, vp]),
    "rrin_frame_from_u8": (ci, [vp, ci, ci, ci, ci, ci, vp, vp]),
    "rrin_frame_to_u8": (ci, [vp, ci, ci, ci, ci, vp, vp]),
}


def header_symbols():
    """Every function the public header declares (used by the CPU test-suite)."""
    with open(HEADER) as f:
        return sorted(set(re.findall(r"RRIN_API\s+[\w\s\*]+?\b(rrin_\w+)\s*\(", f.read())))


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m rrin_b200.build` "
                "(or __graft_entry__.build()). rrin_b200 has no CPU / torch fallback.")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype, fn.argtypes = res, args
        _lib = l
    return _lib


class RrinError(RuntimeError):
    pass


def check(status: int, what: str = ""):
    if status != 0:
        msg = lib().rrin_last_error().decode(errors="replace")
        raise RrinError(f"{what or 'rrin_b200'} failed (status {status}): {msg}")
