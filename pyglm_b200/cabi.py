"""ctypes binding of the C ABI in include/pyglm_b200.h.

The library is the ONLY compute backend: if it cannot be loaded, or no sm_100 device is present when a
compute entry point is called, this module raises -- there is no CPU or PyTorch fallback.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIBPATH = os.path.join(_HERE, "_lib", "libpyglm_b200.so")

c_int, c_ll, c_ull, c_uint = ctypes.c_int, ctypes.c_longlong, ctypes.c_ulonglong, ctypes.c_uint
ptr, size_t = ctypes.c_void_p, ctypes.c_size_t

# name -> (restype, argtypes); mirrors include/pyglm_b200.h one to one
SIGNATURES = {
    "pyglm_last_error": (ctypes.c_char_p, []),
    "pyglm_abi_version": (c_int, []),
    "pyglm_device_check": (c_int, []),
    "pyglm_filter_spikes": (c_int, [ptr, ptr, c_int, c_int, c_int, c_int, c_int, ptr, c_int, ptr]),
    "pyglm_pack_design": (c_int, [ptr, c_int, c_int, ptr, c_int, ptr]),
    "pyglm_unpack_design": (c_int, [ptr, c_int, c_int, c_int, ptr, ptr]),
    "pyglm_activation": (c_int, [ptr, c_int, ptr, c_int, c_int, c_int, c_int, ptr, c_int, ptr]),
    "pyglm_loglik": (c_int, [ptr, c_int, ptr, c_int, c_int, c_int, c_int, ptr, c_int, c_int, ptr, ptr, ptr]),
    "pyglm_means": (c_int, [ptr, c_int, ptr, c_int, c_int, c_int, c_int, ptr, c_int, ptr]),
    "pyglm_pg_draw": (c_int, [ptr, c_int, c_ll, c_int, ptr, c_int, c_ull, c_uint, c_ll, c_int, c_int, ptr]),
    "pyglm_pg_draw_ws_bytes": (size_t, [c_ll, c_int]),
    "pyglm_pg_draw_ws": (c_int, [ptr, c_int, c_ll, c_int, ptr, c_int, c_ull, c_uint, c_ll, c_int, c_int, ptr, size_t,
                                 ptr]),
    "pyglm_philox_uniforms": (c_int, [c_ull, c_uint, c_ull, c_int, c_int, ptr, ptr]),
    "pyglm_gram_tiles": (c_int, [c_int, c_int, ptr, c_int]),
    "pyglm_gram_slabs": (c_int, [c_int, c_int, c_int]),
    "pyglm_weighted_gram": (c_int, [ptr, c_int, c_int, ptr, c_int, c_int, ptr, c_int, c_int, ptr, c_ll, c_int,
                                    c_int, ptr, ptr]),
    "pyglm_gram_tc_geometry": (c_int, [c_int, c_int, c_ll, c_int, ptr]),
    "pyglm_column_max": (c_int, [ptr, c_int, c_ll, c_int, ptr, ptr, ptr]),
    "pyglm_gram_tc_build_z": (c_int, [ptr, c_int, c_ll, c_int, ptr, c_int, ptr, c_ll, c_ll, ptr]),
    "pyglm_gram_tc_build_z_slab": (c_int, [ptr, c_int, c_ll, c_ll, c_int, ptr, c_int, ptr, c_ll, c_ll, ptr]),
    "pyglm_gram_tc_slice_digits": (c_int, [ptr, c_int, c_ll, c_int, c_int, ptr, ptr, c_int, c_ll, c_int, ptr]),
    "pyglm_gram_tc_slice_omega": (c_int, [ptr, c_int, c_ll, c_int, c_int, ptr, ptr, ptr, c_int, c_ll, c_int, ptr]),
    "pyglm_gram_tc_mma": (c_int, [ptr, ptr, c_int, c_int, c_ll, c_int, ptr, c_ll, c_int, ptr]),
    "pyglm_gram_tc_mma_probe": (c_int, [ptr, ptr, c_int, c_int, c_ll, c_int, ptr, c_ll, ptr]),
    "pyglm_gram_tc_finalize": (c_int, [ptr, c_ll, ptr, ptr, c_int, c_int, c_int, ptr, c_ll, c_int, ptr]),
    "pyglm_gram_tc_finalize_peers": (c_int, [ptr, c_int, c_ll, c_ll, ptr, ptr, c_int, c_int, c_int, ptr, c_ll, c_int, ptr]),
    "pyglm_peer_push": (c_int, [ptr, c_ll, ptr, c_int, c_ll, ptr]),
    "pyglm_gram_tc_stream_tiles": (c_int, [c_int, ptr, c_int]),
    "pyglm_gram_tc_quantize": (c_int, [ptr, c_int, c_ll, c_ll, c_int, ptr, ptr, ptr, c_ll, ptr]),
    "pyglm_gram_tc_mma_stream": (c_int, [ptr, ptr, ptr, c_int, c_int, c_ll, c_int, ptr, c_int, ptr, c_ll, c_int, ptr]),
    "pyglm_generate": (c_int, [ptr, ptr, ptr, c_int, c_int, c_int, c_ll, c_ull, c_uint, ctypes.c_double, ptr, c_int, ptr,
                               ptr, ptr]),
    "pyglm_spike_slab_workspace_doubles": (size_t, [c_int, c_int, c_int]),
    "pyglm_scan_randomness": (c_int, [c_int, c_int, c_int, c_int, c_ull, c_uint, ptr, ptr, ptr, c_int, ptr]),
    "pyglm_spike_slab_update": (c_int, [c_int, c_int, c_int, ptr, c_ll, c_int, ptr, c_int, ptr, ptr, ptr, ptr,
                                        ptr, ptr, ptr, ptr, ptr, c_int, ptr, ptr, ptr, ptr, ptr, ptr, ptr, ptr,
                                        ptr]),
}

_lib = None


class PyglmCudaError(RuntimeError):
    pass


def load():
    """Load libpyglm_b200.so and bind every symbol of the header (raises if any is missing)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIBPATH):
        raise PyglmCudaError(
            "pyglm_b200 CUDA library not built: %s is missing. Run `python -m pyglm_b200.build` "
            "(needs nvcc). There is no CPU fallback." % LIBPATH)
    lib = ctypes.CDLL(LIBPATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status, what):
    if status != 0:
        msg = load().pyglm_last_error().decode("utf-8", "replace")
        raise PyglmCudaError("%s failed (status %d): %s" % (what, status, msg))


def call(name, *args):
    """Call an int-status entry point and raise PyglmCudaError on failure."""
    check(getattr(load(), name)(*args), name)
