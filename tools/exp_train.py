"""R training step (train_r.lua:138-170) at the reference's batch size: time per step on the GPU (CUDA events, whole step incl. the
G forward of the batch) beside the same step in PyTorch-CPU autograd (the stand-in for Torch7's nn on the host cores).
python tools/exp_train.py [batch]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch
from __graft_entry__ import load_package
pkg = load_package()
from test_gpu_train import _masks, _torch_step
C, H, W, nd = 1, 32, 32, 100
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
rng = np.random.default_rng(0)
ctx = pkg.Context(0)
ctx.load_G(C, H, W, nd, pkg.weights.init_G(C, H, W, nd))
rb = pkg.weights.init_R(C, H, W, nd)
ctx.train_R_init(C, H, W, nd, rb)
noise = rng.standard_normal(size=(B, nd)).astype(np.float32)
masks = _masks(rng, B, C, H, W, False)
for _ in range(3):
    ctx.train_R_step(noise, masks)
ctx.profile_reset(); ctx.profile_enable(True)
t0 = time.time()
n = 20
for _ in range(n):
    ctx.train_R_step(noise, masks)
wall = (time.time() - t0) / n * 1e3
ctx.profile_enable(False)
pr = ctx.profile()
ms = pr["train_R_step"]["ms"] / n
g_ms = sum(v["ms"] for k, v in pr.items() if k.startswith("g_")) / n
print(f"GPU: batch {B}: {ms:.3f} ms per step on the device (R forward + backward + Adam; the batch's G forward adds {g_ms:.3f} ms), {wall:.3f} ms wall per call, "
      f"{B / (ms + g_ms) * 1e3:.0f} faces/s; {6 * B * 174.7e6 / (ms * 1e-3) * 1e-12:.2f} TFLOP/s fp32 on R's 3 x 174.7 MMAC per face")
torch.set_num_threads(os.cpu_count() or 1)
images = ctx.forward_G(noise)
t0 = time.time()
for _ in range(3):
    _torch_step(pkg, rb, C, H, W, nd, images, noise, masks, False, False, 0.0, 1e-4, 1.0)
print(f"PyTorch-CPU autograd, {torch.get_num_threads()} threads: {(time.time() - t0) / 3 * 1e3:.1f} ms per step (forward + backward, without the optimiser)")
ctx.close()
