"""Weight-blob layout and random initialisation for G3 / R_default.

The blob is the flat float32 concatenation of every parameter in models.lua module
order (include/ganrev.h "Weight blob"): conv weight [Cout][Cin][3][3] then bias, Linear
weight [out][in] then bias, each BatchNorm as gamma, beta, running_mean, running_var.

`init_*` reproduces the *distribution* of the reference's "heuristic" initialisation
(weight-init.lua:14-16, 52-73: U(+-1/sqrt(fan_in)) weights, every bias zeroed, BN gamma
~U(0,1), running mean 0 / var 1 [upstream nn defaults]); Torch's MT19937 stream itself
cannot be reproduced without Torch7.  `stress=True` gives O(1) activations and non-zero
bias/beta/mean/var so that parity tests exercise every term.
"""
import numpy as np


def g_layout(C, H, W, nd):
    """models.lua:115-132 (create_G3)."""
    F = 512 * (H // 4) * (W // 4)
    return [
        ("lin.w", (F, nd)), ("lin.b", (F,)),
        ("bn0.g", (F,)), ("bn0.b", (F,)), ("bn0.m", (F,)), ("bn0.v", (F,)),
        ("c1.w", (256, 512, 3, 3)), ("c1.b", (256,)),
        ("bn1.g", (256,)), ("bn1.b", (256,)), ("bn1.m", (256,)), ("bn1.v", (256,)),
        ("c2.w", (128, 256, 3, 3)), ("c2.b", (128,)),
        ("bn2.g", (128,)), ("bn2.b", (128,)), ("bn2.m", (128,)), ("bn2.v", (128,)),
        ("c3.w", (C, 128, 3, 3)), ("c3.b", (C,)),
    ]


def r_layout(C, H, W, nd):
    """models.lua:409-451 (create_R_default)."""
    F = 128 * (H // 4) * (W // 4)
    out = []
    chans = [(C, 64), (64, 64), (64, 64), (64, 128), (128, 128), (128, 128)]
    for i, (ci, co) in enumerate(chans, start=1):
        out += [(f"c{i}.w", (co, ci, 3, 3)), (f"c{i}.b", (co,)),
                (f"bn{i}.g", (co,)), (f"bn{i}.b", (co,)), (f"bn{i}.m", (co,)), (f"bn{i}.v", (co,))]
    out += [("l1.w", (512, F)), ("l1.b", (512,)),
            ("bn7.g", (512,)), ("bn7.b", (512,)), ("bn7.m", (512,)), ("bn7.v", (512,)),
            ("l2.w", (nd, 512)), ("l2.b", (nd,))]
    return out


def blob_floats(layout):
    return int(sum(int(np.prod(s)) for _, s in layout))


def pack(params, layout):
    parts = []
    for name, shape in layout:
        a = np.asarray(params[name], dtype=np.float32)
        assert a.shape == tuple(shape), (name, a.shape, shape)
        parts.append(a.ravel())
    return np.ascontiguousarray(np.concatenate(parts))


def unpack(blob, layout):
    blob = np.asarray(blob, dtype=np.float32).ravel()
    assert blob.size == blob_floats(layout), (blob.size, blob_floats(layout))
    out, o = {}, 0
    for name, shape in layout:
        n = int(np.prod(shape))
        out[name] = blob[o:o + n].reshape(shape)
        o += n
    return out


def _init(layout, seed, stress):
    rng = np.random.default_rng(seed)
    p = {}
    for name, shape in layout:
        kind = name.split(".")[1]
        if kind == "w":
            fan_in = int(np.prod(shape[1:]))
            # weight-init.lua:14-16 + Torch's reset(): U(+-1/sqrt(fan_in)); stress: He-uniform
            bound = np.sqrt(6.0 / fan_in) if stress else 1.0 / np.sqrt(fan_in)
            p[name] = rng.uniform(-bound, bound, size=shape).astype(np.float32)
            p["_fan_" + name.split(".")[0]] = fan_in
        elif name.startswith("bn"):
            if kind == "g":
                p[name] = (rng.uniform(0.5, 1.5, size=shape) if stress else rng.uniform(0.0, 1.0, size=shape)).astype(np.float32)
            elif kind == "b":   # weight-init.lua:70-72 zeroes every .bias, including BN beta
                p[name] = (rng.normal(0, 0.1, size=shape) if stress else np.zeros(shape)).astype(np.float32)
            elif kind == "m":
                p[name] = (rng.normal(0, 0.1, size=shape) if stress else np.zeros(shape)).astype(np.float32)
            elif kind == "v":
                p[name] = (rng.uniform(0.5, 1.5, size=shape) if stress else np.ones(shape)).astype(np.float32)
        else:                   # conv / linear bias: zeroed by weight-init.lua:70-72
            p[name] = (rng.uniform(-0.1, 0.1, size=shape) if stress else np.zeros(shape)).astype(np.float32)
    return {k: v for k, v in p.items() if not k.startswith("_")}


def init_G(C, H, W, nd, seed=1, stress=False):
    lay = g_layout(C, H, W, nd)
    return pack(_init(lay, seed, stress), lay)


def init_R(C, H, W, nd, seed=2, stress=False):
    lay = r_layout(C, H, W, nd)
    return pack(_init(lay, seed, stress), lay)
