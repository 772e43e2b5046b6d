"""gan-reverser_b200: B200-native apply_r hot path (G/R inference, cosine search, kmeans,
fix/anomaly) behind a C-ABI shared library, with a host-side mirror of the reference's
Lua entry points.  The CUDA library is loaded lazily and there is no CPU fallback."""
from . import weights, _lib, models, nn_utils, apply_r, sample, dist, t7, present, train_r  # noqa: F401
from ._lib import Context, GanrevError  # noqa: F401

MODELS = models
NN_UTILS = nn_utils
