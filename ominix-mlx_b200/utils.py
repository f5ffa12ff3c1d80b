"""Mirror of mlx-rs-core/src/utils.rs: initialize_rope, masks, scaled_dot_product_attention."""
import enum

import torch

from . import _lib, fast, nn


def initialize_rope(dims, base, traditional, scaling_config=None, max_position_embeddings=0):
    """utils.rs:52-97: rope_scaling type/rope_type in {default, linear}; linear => scale = 1/factor."""
    rope_type = "default"
    if scaling_config:
        rope_type = scaling_config.get("type", scaling_config.get("rope_type", "default"))
    if rope_type in ("default", "linear"):
        scale = 1.0
        if rope_type == "linear":
            if "factor" not in scaling_config:
                raise _lib.Exception_('key "factor" is not found in scaling config')
            try:
                scale = 1.0 / float(scaling_config["factor"])
            except (TypeError, ValueError):
                raise _lib.Exception_('key "factor" is not a valid float')
        return nn.RopeBuilder(dims).traditional(traditional).base(base).scale(scale).build()
    raise _lib.Exception_(f"Unsupported RoPE type {rope_type!r}")


class SdpaMask(enum.Enum):
    """utils.rs:105-116; SdpaMask::Array(&Array) is passed as the tensor itself."""
    Causal = "causal"


class AttentionMask:
    """utils.rs:118-131."""

    Causal = "causal"

    def __init__(self, array=None):
        self.array = array


def create_causal_mask(N, offset=None, window_size=None, lengths=None, device=None):
    """utils.rs:134-153 -> bool [N, offset+N]; built with torch index arithmetic on the device
    (mask construction is host-side plumbing in the reference too: arange/ge/logical_and)."""
    offset = offset or 0
    device = device or torch.device("cuda", torch.cuda.current_device())
    rinds = torch.arange(offset + N, device=device)[None, :]
    linds = torch.arange(offset, offset + N, device=device)[:, None]
    mask = linds >= rinds
    if window_size is not None:
        mask = mask & (linds <= rinds + window_size)
    return mask


def create_attention_mask(h, cache=None, return_array=None):
    """utils.rs:156-188: None for T == 1, else "causal" or a bool array."""
    return_array = bool(return_array)
    T = h.shape[1]
    if T <= 1:
        return None
    offset, window = 0, None
    c = cache[0] if cache else None
    if c is not None:
        offset = c.offset()
        ms = c.max_size()
        if ms is not None:
            window = ms
            offset = min(offset, ms)
            return_array = return_array or (offset + T) > ms
    if return_array:
        return create_causal_mask(T, offset, window, device=h.device)
    return AttentionMask.Causal


def scaled_dot_product_attention(queries, keys, values, cache, scale, mask=None, stream=None):
    """utils.rs:191-209 (the `_cache` argument is unused there as well)."""
    if mask is SdpaMask.Causal or mask == "causal":
        m = fast.ScaledDotProductAttentionMask.Causal
    else:
        m = mask
    return fast.scaled_dot_product_attention(queries, keys, values, scale, m, stream)
