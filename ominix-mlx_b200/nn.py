"""nn::Rope (mlx-rs/src/nn/positional_encoding.rs:17-137): parameter holder over fast::rope."""
from . import fast


class Rope:
    """RotaryPositionalEncoding {dimensions, traditional=false, base=10000, scale=1}
    (defaults: positional_encoding.rs:57-66)."""

    DEFAULT_TRADITIONAL = False
    DEFAULT_BASE = 10000.0
    DEFAULT_SCALE = 1.0

    def __init__(self, dimensions, traditional=DEFAULT_TRADITIONAL, base=DEFAULT_BASE, scale=DEFAULT_SCALE):
        self.dimensions = int(dimensions)
        self.traditional = bool(traditional)
        self.base = float(base)
        self.scale = float(scale)

    def forward(self, x, offset=0, stream=None):
        """forward(RopeInput{x, offset}) (positional_encoding.rs:120-134).  As in the reference the
        batch and head axes are NOT flattened: every batch row gets position offset + t."""
        return fast.rope(x, self.dimensions, self.traditional, self.base, self.scale, offset, None, stream)

    __call__ = forward


class RopeBuilder:
    """nn::RopeBuilder (builder pattern of the reference)."""

    def __init__(self, dimensions):
        self._d = dict(dimensions=dimensions)

    def traditional(self, v):
        self._d["traditional"] = v
        return self

    def base(self, v):
        self._d["base"] = v
        return self

    def scale(self, v):
        self._d["scale"] = v
        return self

    def build(self):
        return Rope(**self._d)


class RmsNorm:
    """nn::RmsNorm {weight [dims], eps = 1e-5} (mlx-rs/src/nn/normalization.rs:209-270): forward = fast::rms_norm."""

    DEFAULT_EPS = 1e-5

    def __init__(self, weight, eps=DEFAULT_EPS):
        self.weight = weight
        self.eps = float(eps)

    def forward(self, x, stream=None):
        return fast.rms_norm(x, self.weight, self.eps, stream)

    __call__ = forward
