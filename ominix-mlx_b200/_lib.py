"""ctypes binding of libomx_attn.so (include/omx_attn.h).

This is the Python twin of what `mlx-sys` is to the reference (bindgen over mlx-c,
mlx-rs/mlx-sys/build.rs:390-405): raw declarations only, no logic.  There is no CPU or
PyTorch fallback: if the library is missing, or no sm_100a device is present, calls fail.
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("OMX_ATTN_LIB") or os.path.join(_HERE, "libomx_attn.so")  # env: A/B builds only
CSRC = os.path.join(_HERE, "csrc")

OMX_MAX_NDIM = 8
OMX_BOOL, OMX_INT32, OMX_FLOAT16, OMX_FLOAT32, OMX_BFLOAT16 = 0, 7, 9, 10, 12


class OmxArray(ctypes.Structure):
    _fields_ = [("data", ctypes.c_void_p), ("dtype", ctypes.c_int32), ("ndim", ctypes.c_int32),
                ("shape", ctypes.c_int64 * OMX_MAX_NDIM), ("strides", ctypes.c_int64 * OMX_MAX_NDIM)]


class OmxOptionalFloat(ctypes.Structure):
    _fields_ = [("value", ctypes.c_float), ("has_value", ctypes.c_bool)]


class OmxKVCache(ctypes.Structure):
    _fields_ = [("ctx", ctypes.c_void_p)]


class OmxPagedKVCache(ctypes.Structure):
    _fields_ = [("ctx", ctypes.c_void_p)]


OMX_MAX_PEERS = 8


class OmxPeerGroup(ctypes.Structure):
    _fields_ = [("world", ctypes.c_int32), ("rank", ctypes.c_int32),
                ("out", ctypes.c_void_p * OMX_MAX_PEERS), ("flags", ctypes.c_void_p * OMX_MAX_PEERS)]


class OmxLLGroup(ctypes.Structure):
    _fields_ = [("world", ctypes.c_int32), ("rank", ctypes.c_int32),
                ("staging", ctypes.c_void_p * OMX_MAX_PEERS), ("seq", ctypes.c_void_p)]


class Exception_(RuntimeError):
    """Counterpart of mlx_rs::error::Exception {what} (mlx-rs/src/error.rs)."""


_AP = ctypes.POINTER(OmxArray)
_SIGS = {
    # name: (restype, argtypes)
    "omx_set_error_handler": (None, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "omx_last_error": (ctypes.c_char_p, []),
    "omx_version": (ctypes.c_int, []),
    "omx_device_check": (ctypes.c_int, [ctypes.POINTER(ctypes.c_int)]),
    "omx_fast_rms_norm": (ctypes.c_int, [_AP, _AP, _AP, ctypes.c_float, ctypes.c_void_p]),
    "omx_fast_rope": (ctypes.c_int, [_AP, _AP, ctypes.c_int, ctypes.c_bool, OmxOptionalFloat, ctypes.c_float,
                                     ctypes.c_int, _AP, ctypes.c_void_p]),
    "omx_fast_rope_dynamic": (ctypes.c_int, [_AP, _AP, ctypes.c_int, ctypes.c_bool, OmxOptionalFloat,
                                             ctypes.c_float, _AP, ctypes.c_int, _AP, ctypes.c_void_p]),
    "omx_fast_scaled_dot_product_attention": (ctypes.c_int, [_AP, _AP, _AP, _AP, ctypes.c_float,
                                                             ctypes.c_char_p, _AP, _AP, ctypes.c_void_p]),
    "omx_kv_cache_new": (ctypes.c_int, [ctypes.POINTER(OmxKVCache), ctypes.c_int]),
    "omx_kv_cache_free": (ctypes.c_int, [OmxKVCache]),
    "omx_kv_cache_offset": (ctypes.c_int, [OmxKVCache, ctypes.POINTER(ctypes.c_int)]),
    "omx_kv_cache_reset": (ctypes.c_int, [OmxKVCache]),
    "omx_kv_cache_update_and_fetch": (ctypes.c_int, [OmxKVCache, _AP, _AP, _AP, _AP, ctypes.c_void_p]),
    "omx_kv_cache_state": (ctypes.c_int, [OmxKVCache, _AP, _AP]),
    "omx_kv_cache_reserve": (ctypes.c_int, [OmxKVCache, ctypes.c_int]),
    "omx_kv_cache_trim": (ctypes.c_int, [OmxKVCache, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]),
    "omx_concat_kv_cache_new": (ctypes.c_int, [ctypes.POINTER(OmxKVCache)]),
    "omx_concat_kv_cache_free": (ctypes.c_int, [OmxKVCache]),
    "omx_concat_kv_cache_offset": (ctypes.c_int, [OmxKVCache, ctypes.POINTER(ctypes.c_int)]),
    "omx_concat_kv_cache_update_and_fetch": (ctypes.c_int, [OmxKVCache, _AP, _AP, _AP, _AP, ctypes.c_void_p]),
    "omx_attn_decode_fused": (ctypes.c_int, [_AP, _AP, _AP, _AP, OmxKVCache, ctypes.c_int, ctypes.c_bool,
                                             OmxOptionalFloat, ctypes.c_float, _AP, ctypes.c_float, _AP, _AP,
                                             ctypes.c_void_p]),
    "omx_attn_decode_fused_norm": (ctypes.c_int, [_AP, _AP, _AP, _AP, OmxKVCache, _AP, _AP, ctypes.c_float,
                                                  ctypes.c_int, ctypes.c_bool, OmxOptionalFloat, ctypes.c_float, _AP,
                                                  ctypes.c_float, _AP, _AP, ctypes.c_void_p]),
    "omx_attn_prefill_fused": (ctypes.c_int, [_AP, _AP, _AP, _AP, OmxKVCache, _AP, _AP, ctypes.c_float, ctypes.c_int,
                                              ctypes.c_bool, OmxOptionalFloat, ctypes.c_float, _AP, ctypes.c_float,
                                              ctypes.c_char_p, _AP, _AP, _AP, ctypes.c_void_p]),
    "omx_kv_cache_prepare_graph": (ctypes.c_int, [OmxKVCache, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]),
    "omx_kv_cache_advance": (ctypes.c_int, [OmxKVCache, ctypes.c_int, ctypes.c_void_p]),
    "omx_attn_decode_fused_dynamic": (ctypes.c_int, [_AP, _AP, _AP, _AP, OmxKVCache, _AP, _AP, ctypes.c_float,
                                                     ctypes.c_int, ctypes.c_bool, OmxOptionalFloat, ctypes.c_float,
                                                     ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p]),
    "omx_device_counter_add": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]),
    "omx_paged_kv_cache_new": (ctypes.c_int, [ctypes.POINTER(OmxPagedKVCache), ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                              ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_int]),
    "omx_paged_kv_cache_free": (ctypes.c_int, [OmxPagedKVCache]),
    "omx_paged_kv_cache_offset": (ctypes.c_int, [OmxPagedKVCache, ctypes.POINTER(ctypes.c_int)]),
    "omx_paged_kv_cache_lengths": (ctypes.c_int, [OmxPagedKVCache, ctypes.POINTER(ctypes.c_int32)]),
    "omx_paged_kv_cache_free_pages": (ctypes.c_int, [OmxPagedKVCache, ctypes.POINTER(ctypes.c_int64)]),
    "omx_paged_kv_cache_reset": (ctypes.c_int, [OmxPagedKVCache, ctypes.c_int, ctypes.c_void_p]),
    "omx_paged_kv_cache_release": (ctypes.c_int, [OmxPagedKVCache, ctypes.c_int, ctypes.c_void_p]),
    "omx_paged_kv_cache_reserve": (ctypes.c_int, [OmxPagedKVCache, ctypes.c_int, ctypes.c_void_p]),
    "omx_paged_kv_cache_sync_lengths": (ctypes.c_int, [OmxPagedKVCache, ctypes.c_void_p]),
    "omx_paged_kv_cache_trim": (ctypes.c_int, [OmxPagedKVCache, ctypes.c_int, ctypes.c_void_p]),
    "omx_paged_kv_cache_update_and_fetch": (ctypes.c_int, [OmxPagedKVCache, _AP, _AP, _AP, _AP, ctypes.c_void_p]),
    "omx_paged_kv_cache_append_slot": (ctypes.c_int, [OmxPagedKVCache, ctypes.c_int, _AP, _AP, ctypes.c_void_p]),
    "omx_paged_kv_cache_fetch": (ctypes.c_int, [OmxPagedKVCache, _AP, _AP, ctypes.c_void_p]),
    "omx_paged_kv_cache_pages": (ctypes.c_int, [OmxPagedKVCache, ctypes.POINTER(ctypes.c_void_p),
                                                ctypes.POINTER(ctypes.c_void_p),
                                                ctypes.POINTER(ctypes.POINTER(ctypes.c_int32)),
                                                ctypes.POINTER(ctypes.c_int)]),
    "omx_attn_decode_fused_paged": (ctypes.c_int, [_AP, _AP, _AP, _AP, OmxPagedKVCache, _AP, _AP, ctypes.c_float,
                                                   ctypes.c_int, ctypes.c_bool, OmxOptionalFloat, ctypes.c_float,
                                                   ctypes.c_float, ctypes.c_void_p]),
    "omx_attn_decode_fused_sharded": (ctypes.c_int, [_AP, _AP, _AP, _AP, OmxKVCache, ctypes.c_int, ctypes.c_bool,
                                                     OmxOptionalFloat, ctypes.c_float, _AP, ctypes.c_float,
                                                     ctypes.POINTER(OmxPeerGroup), ctypes.c_int, ctypes.c_void_p]),
    "omx_attn_decode_fused_sharded_sync": (ctypes.c_int, [_AP, _AP, _AP, _AP, OmxKVCache, ctypes.c_int, ctypes.c_bool,
                                                          OmxOptionalFloat, ctypes.c_float, _AP, ctypes.c_float,
                                                          ctypes.POINTER(OmxPeerGroup), ctypes.c_int, ctypes.c_void_p]),
    "omx_attn_decode_fused_sharded_ll": (ctypes.c_int, [_AP, _AP, _AP, _AP, OmxKVCache, ctypes.c_int, ctypes.c_bool,
                                                        OmxOptionalFloat, ctypes.c_float, _AP, ctypes.c_float,
                                                        ctypes.POINTER(OmxLLGroup), ctypes.c_int, ctypes.c_void_p]),
    "omx_ll_staging_bytes": (ctypes.c_size_t, [ctypes.c_int, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                               ctypes.c_int]),
    "omx_attn_decode_seqshard": (ctypes.c_int, [_AP, _AP, _AP, _AP, OmxKVCache, ctypes.c_int, ctypes.c_bool,
                                                OmxOptionalFloat, ctypes.c_float, ctypes.c_int, ctypes.c_bool,
                                                ctypes.c_float, ctypes.POINTER(OmxPeerGroup), ctypes.c_void_p]),
    "omx_seqshard_merge": (ctypes.c_int, [_AP, _AP, ctypes.POINTER(OmxPeerGroup), ctypes.c_uint32, ctypes.c_void_p]),
    "omx_peer_wait": (ctypes.c_int, [ctypes.POINTER(OmxPeerGroup), ctypes.c_uint32, ctypes.c_void_p]),
    "omx_dit_rope": (ctypes.c_int, [_AP, _AP, _AP, _AP, ctypes.c_void_p]),
    "omx_dit_joint_attention": (ctypes.c_int, [_AP, _AP, _AP, _AP, ctypes.c_float, _AP, ctypes.c_void_p]),
    "omx_dit_attn_fused": (ctypes.c_int, [_AP, ctypes.c_int, ctypes.POINTER(_AP), ctypes.POINTER(_AP),
                                          ctypes.POINTER(_AP), ctypes.POINTER(_AP), ctypes.POINTER(_AP),
                                          ctypes.c_float, _AP, _AP, ctypes.c_float, _AP, ctypes.c_void_p]),
    "omx_last_kernel": (ctypes.c_char_p, []),
    "omx_launch_count": (ctypes.c_int64, [ctypes.c_bool]),
    "omx_force_kernel": (ctypes.c_int, [ctypes.c_char_p]),
}
EXPORTED_SYMBOLS = tuple(_SIGS)

_lib = None


def build(verbose=False):
    """Compile libomx_attn.so for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(["make", "-C", CSRC, "-j8"], stdout=out)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise Exception_(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(make -C ominix-mlx_b200/csrc). There is no CPU or PyTorch fallback for this path.")
        _lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(_lib, name)
            fn.restype = res
            fn.argtypes = args
    return _lib


def check(status):
    """Guarded::try_from_op (mlx-rs/src/utils/guard.rs:24-48): status != 0 -> Exception{what}."""
    if status != 0:
        raise Exception_(lib().omx_last_error().decode("utf-8", "replace"))


def last_kernel():
    return lib().omx_last_kernel().decode()


def launch_count(reset=False):
    return int(lib().omx_launch_count(bool(reset)))


def force_kernel(name):
    check(lib().omx_force_kernel(name.encode() if name else None))
