"""B-torchcpu: G3 / R_default in PyTorch-CPU -- TEST / BASELINE INFRASTRUCTURE ONLY.

The living descendant of Torch7's TH/THNN running the graphs of models.lua:104-143 and :389-464, used by bench.py's
cpu_baseline leg as a labelled stand-in for "the reference's Torch7 CPU path" (which cannot run here: no lua/luajit/th).
Like everything under oracle/ it may be imported only by tests/, smoke() and bench.py's CPU legs.  PARITY UNPINNED.
"""
import numpy as np
import torch
import torch.nn.functional as F

EPS = 1e-5


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))


class _Net:
    def __init__(self, p):
        self.p = {k: _t(v) for k, v in p.items()}

    def bn(self, x, name):
        p = self.p
        return F.batch_norm(x, p[name + ".m"], p[name + ".v"], p[name + ".g"], p[name + ".b"], training=False, eps=EPS)


class TorchG(_Net):
    """models.lua:115-133"""

    def __init__(self, p, C, H, W, nd):
        super().__init__(p)
        self.H, self.W = H, W

    def __call__(self, noise):
        p = self.p
        with torch.no_grad():
            x = F.relu(self.bn(F.linear(noise, p["lin.w"], p["lin.b"]), "bn0")).view(-1, 512, self.H // 4, self.W // 4)
            x = F.interpolate(x, scale_factor=2, mode="nearest")
            x = F.relu(self.bn(F.conv2d(x, p["c1.w"], p["c1.b"], padding=1), "bn1"))
            x = F.interpolate(x, scale_factor=2, mode="nearest")
            x = F.relu(self.bn(F.conv2d(x, p["c2.w"], p["c2.b"], padding=1), "bn2"))
            return torch.sigmoid(F.conv2d(x, p["c3.w"], p["c3.b"], padding=1))


class TorchR(_Net):
    """models.lua:409-454 (eval mode: Dropout identity, SpatialDropout(0.25) -> x0.75)"""

    def __init__(self, p, C, H, W, nd):
        super().__init__(p)

    def __call__(self, x):
        p = self.p
        with torch.no_grad():
            for i in (1, 2, 3):
                x = F.elu(self.bn(F.conv2d(x, p[f"c{i}.w"], p[f"c{i}.b"], padding=1), f"bn{i}"))
            x = F.max_pool2d(x, 2, 2)
            for i in (4, 5, 6):
                x = F.elu(self.bn(F.conv2d(x, p[f"c{i}.w"], p[f"c{i}.b"], padding=1), f"bn{i}"))
            x = F.max_pool2d(x * 0.75, 2, 2)
            x = x.reshape(x.shape[0], -1)
            x = F.elu(self.bn(F.linear(x, p["l1.w"], p["l1.b"]), "bn7"))
            return F.linear(x, p["l2.w"], p["l2.b"])
