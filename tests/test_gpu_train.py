"""GPU parity of the R training step (csrc/train.cuh; train_r.lua:138-170) against a PyTorch-CPU autograd
restatement of the same graph: R_default (models.lua:389-464) in training mode with the SAME weights, batch and dropout masks,
nn.MSECriterion, the L2 penalty and gradient clamp of train_r.lua:150-163 and optim.adam's update.  fp32 on both sides: the
tolerances below cover the different summation orders only."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
F = torch.nn.functional


from oracle.torch_cpu import train_masks as _masks, train_R_step as _torch_step   # the PyTorch-CPU restatement (also bench.py's train_leg baseline)


@pytest.mark.parametrize("C,H,W,nd,B,fixer,tanh_out", [(1, 32, 32, 32, 8, False, False), (3, 16, 16, 20, 6, True, True), (1, 32, 32, 100, 32, False, False),
                                                       (3, 64, 64, 24, 3, True, False), (1, 16, 16, 12, 37, False, True)])
def test_train_step_matches_torch_autograd(pkg, C, H, W, nd, B, fixer, tanh_out):
    rng = np.random.default_rng(B + nd)
    ctx = pkg.Context(0)
    try:
        gb = pkg.weights.init_G(C, H, W, nd, seed=1, stress=True)
        rb = pkg.weights.init_R(C, H, W, nd, seed=2, stress=True)
        ctx.load_G(C, H, W, nd, gb)
        ctx.train_R_init(C, H, W, nd, rb, tanh_out=tanh_out, fixer=fixer)
        lay = pkg.weights.r_layout(C, H, W, nd)
        noise = (rng.uniform(-1, 1, size=(B, nd)) if tanh_out else rng.standard_normal(size=(B, nd))).astype(np.float32)
        blob = rb.copy()
        hyper = dict(lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8, l1=0.0, l2=1e-4, clamp=1.0)
        m_state = np.zeros_like(blob); v_state = np.zeros_like(blob)
        is_param = pkg.weights.pack({k: np.full(s, 0.0 if k.endswith((".m", ".v")) else 1.0, np.float32) for k, s in lay}, lay) > 0
        for step in range(1, 4):
            masks = _masks(rng, B, C, H, W, fixer)
            images = ctx.forward_G(noise)                                  # the batch the library trains on (same kernels)
            want_loss, want_f, want_g, run = _torch_step(pkg, blob, C, H, W, nd, images, noise, masks, fixer, tanh_out, hyper["l1"], hyper["l2"], hyper["clamp"])
            loss, f = ctx.train_R_step(noise, masks, **hyper)
            assert abs(loss - want_loss) <= 2e-4 * max(1.0, abs(want_loss)), (step, loss, want_loss)
            assert abs(f - want_f) <= 2e-4 * max(1.0, abs(want_f)), (step, f, want_f)
            g = ctx.train_R_state(1)
            # gradients: per tensor, errors relative to the tensor's largest gradient (biases in front of a batch norm have a
            # mathematically zero gradient: both sides hold rounding noise there)
            got, want = pkg.weights.unpack(g, lay), pkg.weights.unpack(want_g, lay)
            for k, _ in lay:
                if k.endswith((".m", ".v")):
                    continue
                scale = max(float(np.abs(want[k]).max()), 1e-6)
                err = float(np.abs(got[k] - want[k]).max())
                if k.endswith(".b") and k[0] in "cl" and k != "l2.b":      # conv / linear bias followed by batch norm: exactly zero in
                    assert err <= 2e-5 and scale <= 2e-5, (step, k, err, scale)   # real arithmetic, rounding noise on both sides
                else:
                    # fp32 sums in another order, amplified by batch norms over few samples: tight in the L2 sense, looser per element
                    rel = float(np.linalg.norm(got[k] - want[k]) / max(np.linalg.norm(want[k]), 1e-12))
                    assert rel <= 3e-3 and err <= 3e-2 * scale + 1e-7, (step, k, rel, err, scale)
            # Adam on the library's own gradients, restated in numpy (optim.adam): parameters must match tightly
            m_state = np.where(is_param, hyper["beta1"] * m_state + (1 - hyper["beta1"]) * g, 0).astype(np.float32)
            v_state = np.where(is_param, hyper["beta2"] * v_state + (1 - hyper["beta2"]) * g * g, 0).astype(np.float32)
            step_size = hyper["lr"] * np.sqrt(1 - hyper["beta2"] ** step) / (1 - hyper["beta1"] ** step)
            new = pkg.weights.unpack(blob, lay)
            new = pkg.weights.pack({**new, **run}, lay)                    # running statistics from the torch side
            new = np.where(is_param, new - np.float32(step_size) * m_state / (np.sqrt(v_state) + np.float32(hyper["eps"])), new).astype(np.float32)
            got_blob = ctx.train_R_state(0)
            np.testing.assert_allclose(got_blob[is_param], new[is_param], rtol=2e-5, atol=2e-7)
            np.testing.assert_allclose(got_blob[~is_param], new[~is_param], rtol=2e-4, atol=2e-6)   # running mean / var
            blob = got_blob                                                # next step starts from the library's parameters
        # the trained blob drives the inference kernels
        ctx.load_R(0, C, H, W, nd, blob, tanh_out=tanh_out)
        att = ctx.forward_R(0, ctx.forward_G(noise))
        assert np.isfinite(att).all()
    finally:
        ctx.close()


def test_train_reduces_loss(pkg):
    """A few dozen steps on fresh batches (reference defaults: batch 32, Adam 1e-3, L2 1e-4, clamp 1): the criterion goes down."""
    C, H, W, nd, B = 1, 32, 32, 32, 32
    rng = np.random.default_rng(3)
    ctx = pkg.Context(0)
    try:
        ctx.load_G(C, H, W, nd, pkg.weights.init_G(C, H, W, nd, seed=1, stress=True))
        ctx.train_R_init(C, H, W, nd, pkg.weights.init_R(C, H, W, nd, seed=2))
        losses = []
        for _ in range(40):
            noise = rng.standard_normal(size=(B, nd)).astype(np.float32)
            losses.append(ctx.train_R_step(noise, _masks(rng, B, C, H, W, False))[0])
        assert np.isfinite(losses).all()
        assert np.mean(losses[-5:]) < 0.9 * np.mean(losses[:5]), losses
    finally:
        ctx.close()


def test_train_error_paths(pkg):
    ctx = pkg.Context(0)
    try:
        with pytest.raises(pkg.GanrevError):
            ctx._train_floats = 10
            ctx.train_R_state(0)                                            # not initialised
        rb = pkg.weights.init_R(1, 32, 32, 32)
        ctx.train_R_init(1, 32, 32, 32, rb)
        with pytest.raises(pkg.GanrevError):                                # G not loaded
            ctx.train_R_step(np.zeros((4, 32), np.float32), _masks(np.random.default_rng(0), 4, 1, 32, 32, False))
        ctx.load_G(1, 32, 32, 32, pkg.weights.init_G(1, 32, 32, 32))
        with pytest.raises(pkg.GanrevError):                                # wrong mask size
            ctx.train_R_step(np.zeros((4, 32), np.float32), [np.ones((3,), np.uint8)])
        with pytest.raises(pkg.GanrevError):                                # wrong blob size
            ctx.train_R_init(1, 32, 32, 32, rb[:-1])
    finally:
        ctx.close()
