"""DiT joint text+image attention of the image crates (manual matmul/softmax chain in the
reference: flux-klein-mlx/src/klein_model.rs:124-162,460-483,651-659;
zimage-mlx/src/zimage_model.rs:208-235,355-384)."""
import torch

from . import _lib
from .array import desc, ref, stream_ptr


def apply_rope(x, cos, sin, stream=None, out=None):
    """apply_rope / apply_rope_3axis: x [B,S,H,D], cos/sin [B,S,D/2] (or [B,S,1,D/2]) in x's dtype;
    adjacent pairs, out0 = x0*c - x1*s, out1 = x1*c + x0*s."""
    if out is None:
        out = torch.empty(x.shape, dtype=x.dtype, device=x.device)
    xd, od, cd, sd = desc(x), desc(out), desc(cos), desc(sin)
    _lib.check(_lib.lib().omx_dit_rope(ref(od), ref(xd), ref(cd), ref(sd), stream_ptr(stream)))
    return out


def joint_attention(q, k, v, scale, add_mask=None, out_dtype=None, stream=None):
    """softmax(scale * Q K^T [+ mask]) V over the concatenated [txt; img] sequence.
    q/k/v: [B,S,H,D] storage (the crates' layout before their transpose) -> out [B,S,H,D].
    out_dtype=torch.float32 reproduces the reference's promotion for 16-bit inputs."""
    B, S, H, D = q.shape
    out = torch.empty((B, q.shape[1], H, v.shape[3]), dtype=out_dtype or q.dtype, device=q.device)
    t = lambda a: a.transpose(1, 2)  # noqa: E731  [B,S,H,D] -> [B,H,S,D] view
    qd, kd, vd, od, md = desc(t(q)), desc(t(k)), desc(t(v)), desc(t(out)), desc(add_mask)
    _lib.check(_lib.lib().omx_dit_joint_attention(ref(od), ref(qd), ref(kd), ref(vd), float(scale), ref(md),
                                                  stream_ptr(stream)))
    return out
