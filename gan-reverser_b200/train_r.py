"""Mirror of train_r.lua's training loop (train_r.lua:129-170) on top of the library's R training step.

    local noise = NN_UTILS.createNoiseInputs(OPT.batchSize)          -> nn_utils.createNoiseInputs
    local images = MODEL_G:forward(noise):clone()                    -> inside ganrev_train_R_step (G in eval mode)
    optim.adam(fevalR, PARAMETERS_R, OPTSTATE.adam.R)                -> ganrev_train_R_step (forward, MSE, backward, penalties, clamp, Adam)

Dropout is never drawn inside the library: `draw_masks` draws the keep-masks of one batch with a seeded numpy generator
(Torch's own MT19937 stream cannot be reproduced without Torch7) in the module order include/ganrev.h documents.
"""
import numpy as np

from . import nn_utils, weights


def draw_masks(rng, B, C, H, W, fixer=False):
    """uint8 keep-masks of one batch: [fixer input, p = 0.5] five nn.Dropout() (p = 0.5), nn.SpatialDropout(0.25) per (sample,
    channel), nn.Dropout(0.5) behind the first Linear (models.lua:399-449)."""
    shapes = ([(B, C, H, W)] if fixer else []) + [(B, 64, H, W), (B, 64, H, W), (B, 64, H // 2, W // 2), (B, 128, H // 2, W // 2), (B, 128, H // 2, W // 2)]
    masks = [(rng.random(s) >= 0.5).astype(np.uint8) for s in shapes]
    masks.append((rng.random((B, 128)) >= 0.25).astype(np.uint8))
    masks.append((rng.random((B, 512)) >= 0.5).astype(np.uint8))
    return masks


def train(ctx, dimensions, noiseDim, nbBatches, batchSize=32, noiseMethod="normal", fixer=False, r_blob=None, seed=1,
          R_L1=0.0, R_L2=1e-4, R_clamp=1.0, learningRate=1e-3, log=None):
    """train_r.lua main(): G must already be loaded into `ctx` (train_r.lua:96-98 loads it from --G).  Returns (blob, losses);
    the blob is what train_r.lua:228-235 saves and what models.create_R(..., blob=blob) / ganrev_load_R take."""
    C, H, W = dimensions
    rng = np.random.default_rng(seed)
    if r_blob is None:
        r_blob = weights.init_R(C, H, W, noiseDim, seed=seed + 1)         # MODELS.create_R(...) with the heuristic init (train_r.lua:106)
    ctx.train_R_init(C, H, W, noiseDim, r_blob, tanh_out=(noiseMethod != "normal"), fixer=fixer)
    losses = []
    for batchIdx in range(1, nbBatches + 1):
        noise = nn_utils.createNoiseInputs(batchSize, noiseDim, noiseMethod, rng=rng)
        loss, _ = ctx.train_R_step(noise, draw_masks(rng, batchSize, C, H, W, fixer), lr=learningRate, l1=R_L1, l2=R_L2, clamp=R_clamp)
        losses.append(loss)
        if log:
            log("[batch %d of %d (%.2f%%)] loss R=%.4f" % (batchIdx, nbBatches, 100.0 * batchIdx / nbBatches, loss))   # train_r.lua:173
    return ctx.train_R_state(0), losses
