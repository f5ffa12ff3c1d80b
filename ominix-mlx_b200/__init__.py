"""ominix-mlx_b200 -- B200-native (sm_100a) drop-in for the attention hot path of OminiX-MLX:
fast::rope -> KVCache::update_and_fetch -> fast::scaled_dot_product_attention.

Layout
  csrc/        hand-written CUDA (sm_100a) + the extern "C" boundary (include/omx_attn.h)
  _lib.py      ctypes declarations (what mlx-sys is to the reference)
  fast.py      mlx_rs::fast::{rope, scaled_dot_product_attention}
  nn.py        nn::Rope / RopeBuilder
  cache.py     mlx-rs-core cache.rs: KeyValueCache, KVCache, ConcatKeyValueCache (+ PagedKVCache: page pool)
  utils.py     mlx-rs-core utils.rs: initialize_rope, masks, scaled_dot_product_attention
  attention.py the fused decode step (rope + append + attention in one launch)
  dit.py       FLUX.2-klein / Z-Image joint attention
  parallel.py  batch / kv-head sharding over 1..8 B200 (NCCL only for the head-sharded gather)

The package name has a hyphen (it is the repo's name); import it with
importlib.import_module("ominix-mlx_b200").
"""
from . import _lib, array, attention, cache, dit, fast, nn, parallel, utils  # noqa: F401
from ._lib import Exception_ as Exception  # noqa: A001,F401
from ._lib import EXPORTED_SYMBOLS, LIB_PATH, build, force_kernel, last_kernel, launch_count, lib  # noqa: F401
from .attention import (DecodeLoopGraph, attn_decode_fused, attn_decode_fused_dynamic,  # noqa: F401
                        attn_decode_fused_paged, attn_decode_unfused, attn_prefill_fused, device_counter_add)
from .cache import ConcatKeyValueCache, KeyValueCache, KVCache, PagedKVCache  # noqa: F401
from .utils import (AttentionMask, SdpaMask, create_attention_mask, create_causal_mask,  # noqa: F401
                    initialize_rope, scaled_dot_product_attention)
