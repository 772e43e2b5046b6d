"""Torch7 binary serialisation reader: `.net` checkpoints -> the weight blobs of include/ganrev.h
(SURVEY.md section 8f, rank 1).

apply_r.lua loads `{G=<nn.Sequential>, opt=<table>}` (apply_r.lua:62-69) and `{R=..., opt=...}`
(apply_r.lua:92-103) with torch.load; train_r.lua:228-235 / train.lua:241-257 write them with
torch.save after NN_UTILS.prepareNetworkForSave.  No Torch7 runs in this image and the reference
ships no sample file, so this is a restatement of the published format of torch7/File.lua
[upstream torch/torch7, not vendored] -- PARITY UNPINNED against real files; it is pinned only
by round trips through the writer in tests/t7_writer.py, which follows the same description.

Format (binary mode, little endian):
  object   := int32 type, payload
  type     := 0 nil | 1 number (float64) | 2 string (int32 n, n bytes) | 3 table | 4 torch object
              | 5 boolean (int32) | 6, 7, 8 function (index, dumped chunk, upvalues)
  table    := int32 index, [if index not seen before:] int32 n, n x (key object, value object)
  torch    := int32 index, [if new:] version string "V <n>" (absent in legacy files), class name
              string, class payload
  Tensor   := int32 ndim, ndim x long size, ndim x long stride, long storage offset (1-based),
              storage object (torch.*Storage or nil)
  Storage  := long n, n raw elements
  other    := the object's field table, itself an `object` (nn modules have no write method)
`long` is 8 bytes unless the file was written on a 32-bit build (`long_size=4`).
Objects are shared by index: a second occurrence of an index is a reference to the first.
"""
import struct

import numpy as np

from . import weights

TYPE_NIL, TYPE_NUMBER, TYPE_STRING, TYPE_TABLE, TYPE_TORCH, TYPE_BOOLEAN = 0, 1, 2, 3, 4, 5
TYPE_FUNCTION, TYPE_LEGACY_RECUR_FUNCTION, TYPE_RECUR_FUNCTION = 6, 7, 8

_STORAGE_DTYPES = {
    "Double": np.float64, "Float": np.float32, "Half": np.float16, "Long": np.int64, "Int": np.int32,
    "Short": np.int16, "Char": np.int8, "Byte": np.uint8,
}


class T7Error(ValueError):
    pass


class TorchObject:
    """A non-tensor torch class instance (nn / cudnn modules): class name + field table."""

    def __init__(self, torch_typename, fields):
        self.torch_typename = torch_typename
        self.fields = fields if isinstance(fields, dict) else {}

    def __getitem__(self, k):
        return self.fields[k]

    def get(self, k, default=None):
        return self.fields.get(k, default)

    def __repr__(self):
        return f"<{self.torch_typename} {sorted(map(str, self.fields))}>"


class LuaFunction:
    def __init__(self, dumped, upvalues):
        self.dumped, self.upvalues = dumped, upvalues


def _storage_dtype(classname):
    # torch.FloatStorage, torch.CudaStorage (floats), torch.CudaDoubleStorage, torch.LongTensor, ...
    base = classname.split(".", 1)[1]
    for suffix in ("Storage", "Tensor"):
        if base.endswith(suffix):
            base = base[: -len(suffix)]
    if base.startswith("Cuda"):
        base = base[4:] or "Float"
    if base not in _STORAGE_DTYPES:
        raise T7Error(f"unsupported storage class {classname}")
    return np.dtype(_STORAGE_DTYPES[base]).newbyteorder("<")


class Reader:
    def __init__(self, data, long_size=8):
        self.b = memoryview(data)
        self.o = 0
        self.long_size = long_size
        self.memo = {}

    def _take(self, n):
        if self.o + n > len(self.b):
            raise T7Error(f"truncated file: need {n} bytes at offset {self.o}")
        v = self.b[self.o:self.o + n]
        self.o += n
        return v

    def int32(self):
        return struct.unpack("<i", self._take(4))[0]

    def long(self):
        return struct.unpack("<q" if self.long_size == 8 else "<i", self._take(self.long_size))[0]

    def string(self):
        n = self.int32()
        if n < 0:
            raise T7Error(f"negative string length at offset {self.o}")
        return bytes(self._take(n)).decode("latin-1")

    def obj(self):
        t = self.int32()
        if t == TYPE_NIL:
            return None
        if t == TYPE_NUMBER:
            v = struct.unpack("<d", self._take(8))[0]
            return int(v) if v == int(v) and abs(v) < 2 ** 53 else v
        if t == TYPE_BOOLEAN:
            return self.int32() == 1
        if t == TYPE_STRING:
            return self.string()
        if t == TYPE_FUNCTION:
            # legacy function record (torch7 File.lua readObject): NO object index and no memoisation --
            # int32 size, the dumped chunk, then the upvalues object
            dumped = bytes(self._take(self.int32()))
            return LuaFunction(dumped, self.obj())
        if t in (TYPE_TABLE, TYPE_TORCH, TYPE_LEGACY_RECUR_FUNCTION, TYPE_RECUR_FUNCTION):
            index = self.int32()
            if index in self.memo:
                return self.memo[index]
            if t == TYPE_TABLE:
                out = {}
                self.memo[index] = out                      # registered first: tables may contain themselves
                n = self.int32()
                for _ in range(n):
                    k = self.obj()
                    out[k] = self.obj()
                return out
            if t == TYPE_TORCH:
                return self._torch(index)
            dumped = bytes(self._take(self.int32()))
            fn = LuaFunction(dumped, None)
            self.memo[index] = fn
            fn.upvalues = self.obj()
            return fn
        raise T7Error(f"unknown type tag {t} at offset {self.o - 4}")

    def _torch(self, index):
        version = self.string()
        if version.startswith("V ") and version[2:].isdigit():
            classname = self.string()
        else:
            classname = version                               # legacy file: no version string
        if classname.startswith("torch.") and classname.endswith("Storage"):
            n = self.long()
            dt = _storage_dtype(classname)
            arr = np.frombuffer(self._take(n * dt.itemsize), dtype=dt, count=n)
            self.memo[index] = arr
            return arr
        if classname.startswith("torch.") and classname.endswith("Tensor"):
            ndim = self.int32()
            size = [self.long() for _ in range(ndim)]
            stride = [self.long() for _ in range(ndim)]
            offset = self.long() - 1
            placeholder = {}
            self.memo[index] = placeholder
            storage = self.obj()
            if storage is None or ndim == 0:
                t = np.zeros(size if ndim else (0,), _storage_dtype(classname).newbyteorder("="))
            else:
                need = offset + sum((s - 1) * st for s, st in zip(size, stride)) + 1 if all(s > 0 for s in size) else 0
                if offset < 0 or need > storage.size:
                    raise T7Error(f"tensor view [{size} / {stride} @ {offset}] exceeds its storage of {storage.size}")
                t = np.lib.stride_tricks.as_strided(storage[offset:], shape=size, strides=[st * storage.itemsize for st in stride])
                t = np.array(t, dtype=storage.dtype.newbyteorder("="))   # own, contiguous, native-endian copy
            self.memo[index] = t
            return t
        obj = TorchObject(classname, None)
        self.memo[index] = obj
        fields = self.obj()
        obj.fields = fields if isinstance(fields, dict) else {}
        return obj


def loads(data, long_size=8):
    r = Reader(data, long_size)
    out = r.obj()
    return out


def load(path, long_size=8):
    with open(path, "rb") as f:
        return loads(f.read(), long_size)


# ---------------------------------------------------------------------------------------------
# nn.Sequential -> weight blob
# ---------------------------------------------------------------------------------------------
def _lua_list(tbl):
    """A Lua array-table {1=..., 2=..., ...} as a Python list."""
    out, i = [], 1
    while i in tbl:
        out.append(tbl[i])
        i += 1
    return out


def flatten_modules(module):
    """Leaf modules of (nested) containers in forward order (nn.Sequential:add order, models.lua)."""
    if isinstance(module, TorchObject) and isinstance(module.get("modules"), dict):
        out = []
        for m in _lua_list(module["modules"]):
            out += flatten_modules(m)
        return out
    return [module]


def _kind(m):
    name = m.torch_typename.split(".", 1)[1] if isinstance(m, TorchObject) else ""
    if name == "Linear":
        return "linear"
    if name in ("SpatialConvolution", "SpatialConvolutionMM"):
        return "conv"
    if name in ("BatchNormalization", "SpatialBatchNormalization"):
        return "bn"
    return None


def _bn_params(m, n, where):
    g, b, mean = m.get("weight"), m.get("bias"), m.get("running_mean")
    if g is None or b is None:
        raise T7Error(f"{where}: BatchNormalization without affine parameters is not what models.lua builds")
    if m.get("running_var") is not None:
        var = np.asarray(m["running_var"], np.float64)
    elif m.get("running_std") is not None:                     # older nn / cudnn: running_std = 1/sqrt(var + eps)
        eps = float(m.get("eps", 1e-5))
        var = 1.0 / np.square(np.asarray(m["running_std"], np.float64)) - eps
    else:
        raise T7Error(f"{where}: no running_var / running_std")
    eps = float(m.get("eps", 1e-5))
    if abs(eps - 1e-5) > 1e-12:                                # the library folds BN with eps = 1e-5 (include/ganrev.h)
        var = var + (eps - 1e-5)
    out = [np.asarray(a, np.float32).reshape(-1) for a in (g, b, mean, var)]
    for a in out:
        if a.size != n:
            raise T7Error(f"{where}: BatchNormalization has {a.size} features, expected {n}")
    return out


def _blob_from(model, layout, where):
    """Walk the leaf modules; every Linear / SpatialConvolution / BatchNormalization must match the next
    entries of `layout` (weights.g_layout / r_layout) in order and shape."""
    mods = [m for m in flatten_modules(model) if _kind(m)]
    params, li = {}, 0
    names = [n for n, _ in layout]
    shapes = dict(layout)
    for m in mods:
        if li >= len(names):
            raise T7Error(f"{where}: more parameterised modules than the architecture has")
        base = names[li].split(".")[0]
        kind = _kind(m)
        expects_bn = names[li].endswith(".g")
        if expects_bn != (kind == "bn") or (not expects_bn and (kind == "conv") != base.startswith("c")):
            raise T7Error(f"{where}: found {m.torch_typename} where the architecture has {names[li]}")
        if kind in ("linear", "conv"):
            want_w, want_b = shapes[f"{base}.w"], shapes[f"{base}.b"]
            w = np.asarray(m["weight"], np.float32)
            if w.size != int(np.prod(want_w)):
                raise T7Error(f"{where}: {m.torch_typename} weight has {w.size} elements, {base} needs {want_w}")
            if kind == "conv" and (int(m.get("kW", 3)) != 3 or int(m.get("kH", 3)) != 3):
                raise T7Error(f"{where}: {base} is not a 3x3 convolution")
            params[f"{base}.w"] = w.reshape(want_w)               # SpatialConvolutionMM keeps a 2-D view of the same memory
            bias = m.get("bias")
            params[f"{base}.b"] = np.zeros(want_b, np.float32) if bias is None else np.asarray(bias, np.float32).reshape(want_b)
            li += 2
        else:
            n = shapes[f"{base}.g"][0]
            for suffix, a in zip("gbmv", _bn_params(m, n, f"{where}/{base}")):
                params[f"{base}.{suffix}"] = a
            li += 4
    if li != len(names):
        raise T7Error(f"{where}: checkpoint ends before {names[li]}")
    return weights.pack(params, layout)


def dims_from_opt(opt):
    """IMG_DIMENSIONS as apply_r.lua:62-79 derives them."""
    C = 1 if opt.get("colorSpace") == "y" else 3
    return C, int(opt["height"]), int(opt["width"])


def g_blob(ckpt):
    """`torch.load(OPT.G)` content -> (C, H, W, noiseDim, blob) for ganrev_load_G (apply_r.lua:62-69)."""
    opt = ckpt["opt"]
    C, H, W = dims_from_opt(opt)
    nd = int(opt["noiseDim"])
    return C, H, W, nd, _blob_from(ckpt["G"], weights.g_layout(C, H, W, nd), "G")


def r_blob(ckpt, C, H, W, nd):
    """`torch.load(OPT.R).R` -> blob for ganrev_load_R (apply_r.lua:92-103); geometry comes from G's opt."""
    return _blob_from(ckpt["R"], weights.r_layout(C, H, W, nd), "R")
