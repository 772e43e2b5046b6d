"""Independent second statement of G3 / R_default in PyTorch-CPU (torch.nn.functional).

Used only by tests and tests/golden/make_golden.py to cross-check the C oracle
(SURVEY.md section 8c).  Layer order follows models.lua:104-143 and :389-464.
"""
import numpy as np
import torch
import torch.nn.functional as F

EPS = 1e-5


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))


def _bn(x, p, name):
    return F.batch_norm(x, _t(p[name + ".m"]), _t(p[name + ".v"]), _t(p[name + ".g"]), _t(p[name + ".b"]),
                        training=False, eps=EPS)


def forward_G(p, C, H, W, nd, noise):
    with torch.no_grad():
        x = _t(noise)
        x = F.linear(x, _t(p["lin.w"]), _t(p["lin.b"]))
        x = F.relu(_bn(x, p, "bn0"))
        x = x.view(-1, 512, H // 4, W // 4)
        x = F.interpolate(x, scale_factor=2, mode="nearest")
        x = F.relu(_bn(F.conv2d(x, _t(p["c1.w"]), _t(p["c1.b"]), padding=1), p, "bn1"))
        x = F.interpolate(x, scale_factor=2, mode="nearest")
        x = F.relu(_bn(F.conv2d(x, _t(p["c2.w"]), _t(p["c2.b"]), padding=1), p, "bn2"))
        x = torch.sigmoid(F.conv2d(x, _t(p["c3.w"]), _t(p["c3.b"]), padding=1))
        return x.numpy()


def forward_R(p, C, H, W, nd, images, mask=None, tanh_out=False):
    with torch.no_grad():
        x = _t(images)
        if mask is not None:
            x = x * _t(np.asarray(mask, dtype=np.float32))       # v1 dropout: no rescale
        for i in (1, 2, 3):
            x = F.elu(_bn(F.conv2d(x, _t(p[f"c{i}.w"]), _t(p[f"c{i}.b"]), padding=1), p, f"bn{i}"))
        x = F.max_pool2d(x, 2, 2)
        for i in (4, 5, 6):
            x = F.elu(_bn(F.conv2d(x, _t(p[f"c{i}.w"]), _t(p[f"c{i}.b"]), padding=1), p, f"bn{i}"))
        x = F.max_pool2d(x * 0.75, 2, 2)                          # SpatialDropout(0.25) in eval
        x = x.reshape(x.shape[0], -1)
        x = F.elu(_bn(F.linear(x, _t(p["l1.w"]), _t(p["l1.b"])), p, "bn7"))
        x = F.linear(x, _t(p["l2.w"]), _t(p["l2.b"]))
        if tanh_out:
            x = torch.tanh(x)
        return x.numpy()
