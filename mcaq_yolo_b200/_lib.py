"""ctypes binding of libmcaq_b200.so (declared in include/mcaq_b200.h)."""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_longlong, c_uint, c_ulonglong, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libmcaq_b200.so")

MCAQ_F32, MCAQ_BF16, MCAQ_F16 = 0, 1, 2
MCAQ_EGEOM = -5

# name -> (restype, argtypes): one entry per symbol declared in include/mcaq_b200.h
PROTOTYPES = {
    "mcaq_abi_version": (c_int, []),
    "mcaq_error_string": (c_char_p, [c_int]),
    "mcaq_tile_size": (c_int, [c_int, c_int]),
    "mcaq_selftest_division": (c_int, [c_void_p, c_int, c_uint, c_uint, c_ulonglong, c_void_p, c_void_p]),
    "mcaq_debug_stage_clocks": (None, [c_void_p]),
    "mcaq_debug_cluster_split": (None, [c_int]),
    "mcaq_morph_policy": (None, [c_int]),
    "mcaq_debug_morph_threads": (None, [c_int]),
    "mcaq_ranges_reset": (c_int, [c_void_p, c_int, c_void_p]),
    "mcaq_reduce_planes": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int,
                                   c_void_p, c_void_p, c_void_p, c_void_p]),
    "mcaq_reduce_planes_nhwc": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int,
                                        c_void_p, c_void_p, c_void_p, c_void_p]),
    "mcaq_tile_quantize_ranges_nhwc": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                               c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                               c_void_p]),
    "mcaq_ranges_decode": (c_int, [c_void_p, c_int, c_void_p, c_void_p]),
    "mcaq_ranges_ema": (c_int, [c_void_p, c_int, c_double, c_int, c_void_p, c_void_p, c_void_p]),
    "mcaq_ranges_finish": (c_int, [c_void_p, c_int, c_double, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mcaq_build_qtable": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "mcaq_tile_quantize": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                   c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mcaq_tile_quantize_ranges": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                          c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                          c_void_p, c_void_p]),
    "mcaq_tile_quantize_ranges_tma": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                              c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mcaq_morph_fused": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                 c_void_p, c_void_p, c_int, c_void_p, c_float, c_int, c_int, c_float, c_float,
                                 c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mcaq_tile_quantize_train_fwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                             c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "mcaq_tile_quantize_train_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                             c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                             c_void_p]),
    "mcaq_tile_quantize_train_fwd_kd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                                c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                                c_void_p]),
    "mcaq_tile_quantize_train_bwd_kd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                                c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                                c_void_p, c_void_p, c_void_p]),
    "mcaq_debug_train_scalar": (None, [c_int]),
    "mcaq_debug_k3_chunk": (None, [c_int]),
    "mcaq_debug_mapper_cluster": (None, [c_int]),
    "launch_spatial_quantization": (None, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                           c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "mcaq_spatial_quantization": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                          c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "mcaq_level0_status": (c_int, []),
    "mcaq_morph_phi": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                               c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mcaq_morph_fits": (c_int, [c_int, c_int, c_int, c_int, c_int]),
    "mcaq_morph_planes_workspace": (c_longlong, [c_int, c_int, c_int, c_int, c_int]),
    "mcaq_morph_phi_planes": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_longlong, c_void_p,
                                      c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mcaq_complexity": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_void_p]),
    "mcaq_bit_mapper": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_float, c_int, c_int,
                                c_float, c_float, c_float, c_void_p, c_void_p]),
    "mcaq_mapper_steps": (c_int, [c_void_p, c_float, c_int, c_float, c_float, c_void_p, c_void_p]),
    "mcaq_xchg_bytes": (c_longlong, [c_int, c_int]),
    "mcaq_xchg_alloc": (c_int, [c_longlong, POINTER(c_void_p)]),
    "mcaq_xchg_free": (c_int, [c_void_p]),
    "mcaq_xchg_export": (c_int, [c_void_p, c_void_p]),
    "mcaq_xchg_open": (c_int, [c_void_p, POINTER(c_void_p)]),
    "mcaq_xchg_close": (c_int, [c_void_p]),
    "mcaq_xchg_set_timeout_ms": (None, [c_int]),
    "mcaq_xchg_error": (c_int, [c_void_p, POINTER(c_int)]),
    "mcaq_xchg_publish": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_void_p]),
    "mcaq_xchg_merge": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "mcaq_morph_fused_xchg": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                      c_void_p, c_void_p, c_int, c_void_p, c_float, c_int, c_int, c_float, c_float,
                                      c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                      c_void_p]),
    "mcaq_tile_quantize_xchg": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int,
                                        c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mcaq_cmlp_train_scratch_floats": (c_longlong, [c_int]),
    "mcaq_mapper_train_scratch_floats": (c_longlong, [c_int]),
    "mcaq_complexity_train_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                          c_void_p, c_void_p]),
    "mcaq_mapper_train_fwd": (c_int, [c_void_p, c_int, c_void_p, c_float, c_int, c_float, c_float, c_void_p, c_void_p,
                                      c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                      c_void_p, c_float, c_float, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "mcaq_mapper_train_bwd": (c_int, [c_void_p, c_int, c_void_p, c_float, c_int, c_float, c_float, c_void_p, c_void_p,
                                      c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "mcaq_softmask_act": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "mcaq_softmask_train_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                        c_void_p, c_void_p, c_void_p]),
    "mcaq_bit_stats": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mcaq_bit_losses": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_float, c_void_p, c_void_p, c_void_p,
                                c_void_p]),
    "mcaq_soft_mask": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int,
                               c_void_p, c_void_p, c_void_p, c_void_p]),
}

_lib = None


class McaqLibraryError(RuntimeError):
    pass


def load():
    """Load the CUDA library; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise McaqLibraryError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(the MCAQ B200 path has no CPU fallback)")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)          # AttributeError if the .so is stale
            fn.restype = res
            fn.argtypes = args
        if os.environ.get("MCAQ_K2_THREADS"):      # tuning aid: force the morphology kernel's CTA size
            lib.mcaq_debug_morph_threads(int(os.environ["MCAQ_K2_THREADS"]))
        if os.environ.get("MCAQ_K2_SPLIT"):        # tuning aid: force the morphology kernel's cluster split
            lib.mcaq_debug_cluster_split(int(os.environ["MCAQ_K2_SPLIT"]))
        if os.environ.get("MCAQ_MAPPER_CLUSTER"):
            lib.mcaq_debug_mapper_cluster(int(os.environ["MCAQ_MAPPER_CLUSTER"]))
        if os.environ.get("MCAQ_K3_CHUNK"):
            lib.mcaq_debug_k3_chunk(int(os.environ["MCAQ_K3_CHUNK"]))
        _lib = lib
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().mcaq_error_string(rc)
        raise RuntimeError(f"{what} failed: {msg.decode() if msg else rc} (code {rc})")
