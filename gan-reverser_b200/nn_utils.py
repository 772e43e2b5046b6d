"""Host-side mirror of utils/nn_utils.lua for the apply_r path."""
import numpy as np


def forwardBatched(model, input, batchSize=32, **kw):
    """NN_UTILS.forwardBatched (utils/nn_utils.lua:5-33).  The reference slices the input into
    batchSize rows because Torch7 graphs hold one batch of activations; the library chunks
    internally (ganrev_set_option "chunk"), so the whole input goes down in one call and
    batchSize is accepted only for signature compatibility."""
    return model.forward(input, **kw)


def createNoiseInputs(N, noiseDim, method="normal", rng=None):
    """NN_UTILS.createNoiseInputs (utils/nn_utils.lua:39-51): N x noiseDim, N(0,1) or U(-1,1)."""
    rng = rng if rng is not None else np.random.default_rng(1)
    if method == "uniform":
        return rng.uniform(-1.0, 1.0, size=(N, noiseDim)).astype(np.float32)
    if method == "normal":
        return rng.normal(0.0, 1.0, size=(N, noiseDim)).astype(np.float32)
    raise ValueError(f"Unknown noise method '{method}'")


def toBatch(x):
    """NN_UTILS.toBatch (utils/nn_utils.lua:248-263): add a leading batch dimension of 1."""
    return np.asarray(x)[None, ...]
