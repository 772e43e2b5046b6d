"""Host-side mirror of the reference's model constructors (models.lua) for the apply_r path.

MODELS.create_G(dimensions, noiseDim, cuda)                       models.lua:201-203 -> create_G3 :104-143
MODELS.create_R(dimensions, noiseDim, noiseMethod, fixer, cuda)   models.lua:385-464

The returned objects implement the nn.Module protocol apply_r.lua uses (:forward, :evaluate,
:training, :float) but own no arithmetic: forward() calls libganrev_cuda.so.  Weights are either
a caller-supplied blob (include/ganrev.h "Weight blob") or a fresh "heuristic" initialisation
(weight-init.lua) drawn from a seeded numpy generator.
"""
import os

import numpy as np

from . import _lib, weights

_default_ctx = None


def default_context():
    """One context per process on cuda:LOCAL_RANK (replaces cutorch.setDevice, apply_r.lua:52-56)."""
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = _lib.Context(int(os.environ.get("LOCAL_RANK", "0")))
    return _default_ctx


def set_default_context(ctx):
    global _default_ctx
    _default_ctx = ctx


class _Module:
    def evaluate(self):      # nn.Module:evaluate() -- the library always runs eval-mode BN / Dropout
        return self

    def training(self):
        raise NotImplementedError("libganrev_cuda implements inference only (apply_r path)")

    def float(self):         # MODEL:float() (apply_r.lua:106-109): there is no CPU path
        raise NotImplementedError("no CPU fallback: the apply_r path runs on the B200 only")

    def cuda(self):
        return self


class Generator(_Module):
    """G3 (models.lua:104-143)."""

    def __init__(self, ctx, dimensions, noiseDim, blob):
        self.ctx, self.dimensions, self.noiseDim, self.blob = ctx, tuple(dimensions), int(noiseDim), blob
        C, H, W = self.dimensions
        ctx.load_G(C, H, W, self.noiseDim, blob)

    def forward(self, noise):
        noise = np.ascontiguousarray(noise, dtype=np.float32)
        if noise.ndim == 1:
            noise = noise[None, :]
        return self.ctx.forward_G(noise)


class Reverser(_Module):
    """R_default (models.lua:389-464).  The fixer variant's input Dropout(0.5) is always active
    (models.lua:399-406); its Bernoulli mask is drawn here on the host, or passed explicitly."""

    def __init__(self, ctx, slot, dimensions, noiseDim, noiseMethod, fixer, blob, mask_seed=5):
        assert noiseMethod in ("normal", "uniform")            # models.lua:390
        self.ctx, self.slot, self.dimensions, self.noiseDim = ctx, slot, tuple(dimensions), int(noiseDim)
        self.noiseMethod, self.fixer, self.blob = noiseMethod, bool(fixer), blob
        self._mask_rng = np.random.default_rng(mask_seed)
        C, H, W = self.dimensions
        ctx.load_R(slot, C, H, W, self.noiseDim, blob, tanh_out=(noiseMethod != "normal"))

    def draw_mask(self, shape):
        return (self._mask_rng.random(shape) >= 0.5).astype(np.uint8)   # keep with p = 1 - 0.5

    def forward(self, images, mask=None):
        images = np.ascontiguousarray(images, dtype=np.float32)
        if images.ndim == 3:
            images = images[None]
        if self.fixer and mask is None:
            mask = self.draw_mask(images.shape)
        return self.ctx.forward_R(self.slot, images, mask)


def create_G(dimensions, noiseDim, cuda=True, blob=None, seed=1, stress=False, ctx=None):
    assert cuda, "no CPU fallback"
    C, H, W = dimensions
    if blob is None:
        blob = weights.init_G(C, H, W, noiseDim, seed=seed, stress=stress)
    return Generator(ctx or default_context(), dimensions, noiseDim, blob)


def create_R(dimensions, noiseDim, noiseMethod="normal", fixer=False, cuda=True, blob=None, seed=2, stress=False,
             ctx=None, slot=None):
    assert cuda, "no CPU fallback"
    C, H, W = dimensions
    if blob is None:
        blob = weights.init_R(C, H, W, noiseDim, seed=seed, stress=stress)
    if slot is None:
        slot = 1 if fixer else 0
    return Reverser(ctx or default_context(), slot, dimensions, noiseDim, noiseMethod, fixer, blob)


def load_G(path, ctx=None):
    """`torch.load(OPT.G)` of apply_r.lua:62-69 without Torch7: parse the `.net` file (t7.py), take the
    geometry from its `opt` table and load `ckpt.G`.  Returns (Generator, opt)."""
    from . import t7
    ckpt = t7.load(path)
    C, H, W, nd, blob = t7.g_blob(ckpt)
    return Generator(ctx or default_context(), (C, H, W), nd, blob), ckpt["opt"]


def load_R(path, dimensions, noiseDim, noiseMethod="normal", fixer=False, ctx=None, slot=None):
    """`torch.load(OPT.R).R` / `torch.load(OPT.R_fixer).R` of apply_r.lua:92-103 without Torch7."""
    from . import t7
    C, H, W = dimensions
    blob = t7.r_blob(t7.load(path), C, H, W, noiseDim)
    return create_R(dimensions, noiseDim, noiseMethod, fixer, blob=blob, ctx=ctx, slot=slot)
