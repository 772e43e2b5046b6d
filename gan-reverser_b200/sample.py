"""Host-side mirror of the one arithmetic function of sample.lua that sits next to the apply_r path
(SURVEY.md section 8f, rank 3).  Same name and argument meaning as the Lua function; the work is
done by libganrev_cuda.so through the C ABI in include/ganrev.h."""
import numpy as np

from . import models


def findClosestNeighboursOf(images, trainingSet, ctx=None):
    """sample.lua:128-148.  `images`: list/array of image tensors; `trainingSet`: [N x C x H x W] (the Lua
    code loads it with DATASET.loadImages).  Returns a list of (image, closest training image, distance)
    like the Lua table {img, closestImg, closestDist}; with an empty training set the last two are None."""
    ctx = ctx or models.default_context()
    imgs = np.ascontiguousarray(np.stack([np.asarray(i, np.float32) for i in images]) if len(images) else np.zeros((0, 1), np.float32))
    ts = np.ascontiguousarray(trainingSet, np.float32)
    if imgs.shape[0] == 0:
        return []
    ids, dist = ctx.nearest_l2(imgs, ts)
    out = []
    for i in range(imgs.shape[0]):
        if ids[i] < 0:
            out.append((imgs[i], None, None))
        else:
            out.append((imgs[i], ts[ids[i]].copy(), float(dist[i])))
    return out
