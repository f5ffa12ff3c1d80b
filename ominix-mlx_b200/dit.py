"""DiT joint text+image attention of the image crates (manual matmul/softmax chain in the
reference: flux-klein-mlx/src/klein_model.rs:124-162,460-483,651-659;
zimage-mlx/src/zimage_model.rs:208-235,355-384)."""
import ctypes

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


def _ptr_list(descs):
    """list of OmxArray-or-None -> (const omx_array* const*) argument; None when every entry is None."""
    if all(d is None for d in descs):
        return None
    arr = (ctypes.POINTER(_lib.OmxArray) * len(descs))()
    for i, d in enumerate(descs):
        arr[i] = ctypes.pointer(d) if d is not None else None
    return arr


def attn_fused(q, k, v, scale, cos=None, sin=None, q_norm=None, k_norm=None, add_mask=None, out_dtype=None,
               stream=None):
    """The attention of one DiT block with its prologue in ONE launch + one attention kernel
    (klein_model.rs:443-489 / :641-663, zimage_model.rs:345-388).
    q, k, v: a [B,S,H,D] tensor (single stream) or a list of per-stream tensors in K/V order
    ([txt, img] for FLUX.2-klein); q_norm / k_norm: an nn.RmsNorm (or a list, one per stream, entries may
    be None) applied per head before the rotation; cos / sin: [B,S_total,D/2] per-token tables in q's dtype.
    Returns out [B,S_total,H,Dv]; out[:, :S_0] / out[:, S_0:] are the per-stream results."""
    qs, ks, vs = ([t] if isinstance(t, torch.Tensor) else list(t) for t in (q, k, v))
    n = len(qs)

    def norms(x):
        xs = list(x) if isinstance(x, (list, tuple)) else [x] * n
        if len(xs) != n:
            raise _lib.Exception_("one norm per stream expected")
        return xs
    qn, kn = norms(q_norm), norms(k_norm)
    eps_all = {m.eps for m in qn + kn if m is not None}
    if len(eps_all) > 1:
        raise _lib.Exception_("all norms of a fused DiT block must share one eps")
    eps = eps_all.pop() if eps_all else 0.0
    B, H = qs[0].shape[0], qs[0].shape[2]
    S = sum(t.shape[1] for t in qs)
    out = torch.empty((B, S, H, vs[0].shape[3]), dtype=out_dtype or qs[0].dtype, device=qs[0].device)
    qd, kd, vd = [desc(t) for t in qs], [desc(t) for t in ks], [desc(t) for t in vs]
    qw = [desc(m.weight) if m is not None else None for m in qn]
    kw = [desc(m.weight) if m is not None else None for m in kn]
    od, cd, sd, md = desc(out), desc(cos), desc(sin), desc(add_mask)
    _lib.check(_lib.lib().omx_dit_attn_fused(ref(od), n, _ptr_list(qd), _ptr_list(kd), _ptr_list(vd), _ptr_list(qw),
                                             _ptr_list(kw), float(eps), ref(cd), ref(sd), float(scale), ref(md),
                                             stream_ptr(stream)))
    return out
