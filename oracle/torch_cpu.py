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


# ------------------------------------------------------------------ R training step (train_r.lua:138-170), SURVEY 8f rank 4
def train_masks(rng, B, C, H, W, fixer):
    shapes = ([(B, C, H, W)] if fixer else []) + [(B, 64, H, W), (B, 64, H, W), (B, 64, H // 2, W // 2), (B, 128, H // 2, W // 2), (B, 128, H // 2, W // 2)]
    out = [(rng.random(s) >= 0.5).astype(np.uint8) for s in shapes]
    out.append((rng.random((B, 128)) >= 0.25).astype(np.uint8))          # nn.SpatialDropout(0.25): keep with probability 0.75
    out.append((rng.random((B, 512)) >= 0.5).astype(np.uint8))
    return out


def train_R_step(pkg, blob, C, H, W, nd, images, noise, masks, fixer, tanh_out, l1, l2, clamp):
    """R training step of train_r.lua:138-170 (R_default in training mode, nn.MSECriterion, L1/L2 penalties, clamp) by PyTorch-CPU
    autograd; `pkg` is the gan-reverser_b200 package (weight layouts only).  Returns (loss, f, grads as a blob-shaped vector incl. penalties + clamp, new running statistics dict)."""
    lay = pkg.weights.r_layout(C, H, W, nd)
    p = {k: torch.tensor(v.copy(), dtype=torch.float32, requires_grad=not (k.endswith(".m") or k.endswith(".v"))) for k, v in pkg.weights.unpack(blob, lay).items()}
    x = torch.tensor(images)
    mk = [torch.tensor(m.astype(np.float32)) for m in masks]
    if fixer:
        x = x * mk.pop(0)                                                 # nn.Dropout(0.5, true): v1, no rescale
    run = {}

    def bn(z, i):
        rm, rv = p[f"bn{i}.m"].detach().clone(), p[f"bn{i}.v"].detach().clone()
        y = F.batch_norm(z, rm, rv, p[f"bn{i}.g"], p[f"bn{i}.b"], training=True, momentum=0.1, eps=1e-5)
        run[f"bn{i}.m"], run[f"bn{i}.v"] = rm.numpy(), rv.numpy()
        return y

    for i in range(1, 7):
        x = F.elu(bn(F.conv2d(x, p[f"c{i}.w"], p[f"c{i}.b"], padding=1), i))
        if i == 3:
            x = F.max_pool2d(x, 2) * mk[2] * 2.0
        elif i == 6:
            x = F.max_pool2d(x * mk[5][:, :, None, None], 2)
        else:
            x = x * mk[i - 1] * 2.0
    x = x.reshape(x.shape[0], -1)
    x = F.elu(bn(F.linear(x, p["l1.w"], p["l1.b"]), 7)) * mk[6] * 2.0
    pred = F.linear(x, p["l2.w"], p["l2.b"])
    if tanh_out:
        pred = torch.tanh(pred)
    loss = F.mse_loss(pred, torch.tensor(noise))
    loss.backward()
    f = float(loss.detach())
    grads = {}
    for k, v in p.items():
        if v.requires_grad:
            g = v.grad + l1 * torch.sign(v.detach()) + l2 * v.detach()
            f += l1 * float(v.detach().abs().sum()) + l2 * float((v.detach() ** 2).sum()) / 2.0
            grads[k] = (g.clamp(-clamp, clamp) if clamp else g).numpy()
        else:
            grads[k] = np.zeros(v.shape, np.float32)
    return float(loss.detach()), f, pkg.weights.pack(grads, lay), run
