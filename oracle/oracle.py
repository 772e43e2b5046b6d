"""ctypes wrapper around oracle/libganrev_oracle.so -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module (see oracle/ganrev_oracle.h).  PARITY UNPINNED.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libganrev_oracle.so")

f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("ganrev_oracle.c", "ganrev_oracle.h", "Makefile")]
    stale = force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src)
    if stale:
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.orc_blob_floats_G.restype = C.c_size_t
        L.orc_blob_floats_G.argtypes = [C.c_int] * 4
        L.orc_blob_floats_R.restype = C.c_size_t
        L.orc_blob_floats_R.argtypes = [C.c_int] * 4
        L.orc_forward_G.argtypes = [f32p, C.c_int, C.c_int, C.c_int, C.c_int, f32p, C.c_int64, f32p]
        L.orc_forward_R.argtypes = [f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, f32p, C.c_void_p,
                                    C.c_int64, f32p]
        L.orc_cosine.restype = C.c_float
        L.orc_cosine.argtypes = [f32p, f32p, C.c_int]
        L.orc_search_cosine.argtypes = [f32p, C.c_int64, C.c_int, f32p, C.c_int, C.c_int, i64p, f32p]
        L.orc_kmeans_shift.argtypes = [f32p, C.c_int64, C.c_int, C.c_int64]
        L.orc_kmeans.argtypes = [f32p, C.c_int64, C.c_int, C.c_int, C.c_int, f32p, C.c_int, f32p, f32p, i32p]
        L.orc_assign_cosine_min.argtypes = [f32p, C.c_int64, C.c_int, f32p, C.c_int, i32p, f32p]
        L.orc_cluster_members.argtypes = [i32p, f32p, C.c_int64, C.c_int, C.c_int, f32p, C.c_int,
                                          i64p, i32p, f32p]
        L.orc_l2.argtypes = [f32p, f32p, C.c_int64, C.c_int, f64p]
        L.orc_nearest_l2.argtypes = [f32p, C.c_int, f32p, C.c_int64, C.c_int, i64p, f64p]
        L.orc_l2_sequential.argtypes = [f32p, f32p, C.c_int64, C.c_int, f64p]
        L.orc_anomaly_flags.argtypes = [f64p, C.c_int64, C.c_int64, C.c_double, u8p, C.POINTER(C.c_double)]
        L.orc_num_threads.restype = C.c_int
        L.orc_set_num_threads.argtypes = [C.c_int]
        _lib = L
    return _lib


def _chk(rc, what):
    if rc != 0:
        raise RuntimeError(f"oracle {what} failed rc={rc}")


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def blob_floats_G(Cc, H, W, nd):
    return lib().orc_blob_floats_G(Cc, H, W, nd)


def blob_floats_R(Cc, H, W, nd):
    return lib().orc_blob_floats_R(Cc, H, W, nd)


def forward_G(blob, Cc, H, W, nd, noise):
    blob, noise = _f32(blob), _f32(noise)
    assert blob.size == blob_floats_G(Cc, H, W, nd), (blob.size, blob_floats_G(Cc, H, W, nd))
    N = noise.shape[0]
    out = np.empty((N, Cc, H, W), np.float32)
    _chk(lib().orc_forward_G(blob, Cc, H, W, nd, noise, N, out), "forward_G")
    return out


def forward_R(blob, Cc, H, W, nd, images, mask=None, tanh_out=False):
    blob, images = _f32(blob), _f32(images)
    assert blob.size == blob_floats_R(Cc, H, W, nd)
    N = images.shape[0]
    out = np.empty((N, nd), np.float32)
    mp = None
    if mask is not None:
        mask = np.ascontiguousarray(mask, dtype=np.uint8)
        assert mask.size == images.size
        mp = mask.ctypes.data_as(C.c_void_p)
    _chk(lib().orc_forward_R(blob, Cc, H, W, nd, int(tanh_out), images, mp, N, out), "forward_R")
    return out


def cosine(a, b):
    a, b = _f32(a).ravel(), _f32(b).ravel()
    return float(lib().orc_cosine(a, b, a.size))


def search_cosine(db, queries, k):
    db, queries = _f32(db), _f32(queries)
    N, d = db.shape
    Q = queries.shape[0]
    ids = np.empty((Q, k), np.int64)
    sc = np.empty((Q, k), np.float32)
    _chk(lib().orc_search_cosine(db, N, d, queries, Q, k, ids, sc), "search_cosine")
    return ids, sc


def kmeans_shift(x, n_total=None):
    x = _f32(x)
    N, d = x.shape
    return lib().orc_kmeans_shift(x, N, d, N if n_total is None else n_total)


def kmeans(x, k, niter, init, shift=-1):
    x, init = _f32(x), _f32(init)
    N, d = x.shape
    cen = np.empty((k, d), np.float32)
    tot = np.empty((k,), np.float32)
    lab = np.empty((N,), np.int32)
    _chk(lib().orc_kmeans(x, N, d, k, niter, init, shift, cen, tot, lab), "kmeans")
    return cen, tot, lab


def assign_cosine_min(x, centroids):
    x, centroids = _f32(x), _f32(centroids)
    N, d = x.shape
    cl = np.empty((N,), np.int32)
    cv = np.empty((N,), np.float32)
    _chk(lib().orc_assign_cosine_min(x, N, d, centroids, centroids.shape[0], cl, cv), "assign")
    return cl, cv


def cluster_members(cluster, cosv, k, m, images):
    cluster = np.ascontiguousarray(cluster, np.int32)
    cosv = _f32(cosv)
    N = cluster.shape[0]
    images = _f32(images).reshape(N, -1)
    px = images.shape[1]
    ids = np.empty((k, m), np.int64)
    cnt = np.empty((k,), np.int32)
    mean = np.empty((k, px), np.float32)
    _chk(lib().orc_cluster_members(cluster, cosv, N, k, m, images, px, ids, cnt, mean), "cluster_members")
    return ids, cnt, mean


def l2(a, b, sequential=False):
    a, b = _f32(a), _f32(b)
    N = a.shape[0]
    a2, b2 = a.reshape(N, -1), b.reshape(N, -1)
    out = np.empty((N,), np.float64)
    fn = lib().orc_l2_sequential if sequential else lib().orc_l2
    _chk(fn(a2, b2, N, a2.shape[1], out), "l2")
    return out


def nearest_l2(queries, images):
    """sample.lua:128-148: per query the set image with the first strictly smallest torch.dist."""
    q, x = _f32(queries), _f32(images)
    Q, N = q.shape[0], x.shape[0]
    q2 = q.reshape(Q, -1)
    x2 = x.reshape(N, -1) if N else np.zeros((0, q2.shape[1]), np.float32)
    ids = np.empty((Q,), np.int64)
    dist = np.empty((Q,), np.float64)
    _chk(lib().orc_nearest_l2(q2, Q, x2, N, q2.shape[1], ids, dist), "nearest_l2")
    return ids, dist


def anomaly_flags(l2v, n_calc, n_show, quantile):
    l2v = np.ascontiguousarray(l2v, np.float64)
    flags = np.zeros((n_show,), np.uint8)
    thr = C.c_double(0.0)
    _chk(lib().orc_anomaly_flags(l2v, n_calc, n_show, quantile, flags, C.byref(thr)), "anomaly_flags")
    return flags, thr.value


def num_threads():
    return lib().orc_num_threads()


def set_num_threads(n):
    lib().orc_set_num_threads(n)
