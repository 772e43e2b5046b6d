"""Test-only writer for the Torch7 binary format described in gan-reverser_b200/t7.py (torch7/File.lua
[upstream]): lets the reader be exercised without Torch7.  Not part of the product."""
import struct

import numpy as np

_CLASS = {np.dtype(np.float32): "Float", np.dtype(np.float64): "Double", np.dtype(np.int64): "Long",
          np.dtype(np.int32): "Int", np.dtype(np.uint8): "Byte", np.dtype(np.int16): "Short", np.dtype(np.int8): "Char"}


class Module:
    """A torch class instance that serialises as its field table (nn modules)."""

    def __init__(self, typename, **fields):
        self.typename, self.fields = typename, fields


class TensorView:
    """A tensor that is a strided view of a larger storage (offset 0-based here, 1-based on disk)."""

    def __init__(self, storage, size, stride, offset, cuda=False):
        self.storage, self.size, self.stride, self.offset, self.cuda = storage, size, stride, offset, cuda


class Writer:
    def __init__(self, long_size=8, legacy=False, cuda=False):
        self.out = bytearray()
        self.long_size, self.legacy, self.cuda = long_size, legacy, cuda
        self.ids = {}
        self.keep = []          # keeps every indexed object alive: id() values must not be reused

    def i32(self, v):
        self.out += struct.pack("<i", v)

    def long(self, v):
        self.out += struct.pack("<q" if self.long_size == 8 else "<i", v)

    def string(self, s):
        b = s.encode("latin-1")
        self.i32(len(b)); self.out += b

    def _index(self, obj):
        """returns True if the object was already written (a reference suffices)"""
        key = id(obj)
        self.keep.append(obj)
        if key in self.ids:
            self.i32(self.ids[key]); return True
        self.ids[key] = len(self.ids) + 1
        self.i32(self.ids[key]); return False

    def _header(self, classname):
        if not self.legacy:
            self.string("V 1")
        self.string(classname)

    def obj(self, o):
        if o is None:
            self.i32(0)
        elif isinstance(o, bool):
            self.i32(5); self.i32(1 if o else 0)
        elif isinstance(o, (int, float)):
            self.i32(1); self.out += struct.pack("<d", float(o))
        elif isinstance(o, str):
            self.i32(2); self.string(o)
        elif isinstance(o, (dict, list)):
            self.i32(3)
            if self._index(o):
                return
            items = list(o.items()) if isinstance(o, dict) else [(i + 1, v) for i, v in enumerate(o)]
            self.i32(len(items))
            for k, v in items:
                self.obj(k); self.obj(v)
        elif isinstance(o, np.ndarray):
            a = np.ascontiguousarray(o)
            st = [int(np.prod(a.shape[i + 1:])) for i in range(a.ndim)]
            self._tensor(o, a.reshape(-1), list(a.shape), st, 0, self.cuda)
        elif isinstance(o, TensorView):
            self._tensor(o, o.storage, o.size, o.stride, o.offset, o.cuda)
        elif isinstance(o, Module):
            self.i32(4)
            if self._index(o):
                return
            self._header(o.typename)
            self.obj(o.fields)
        else:
            raise TypeError(type(o))

    def _tensor(self, key, storage, size, stride, offset, cuda):
        self.i32(4)
        if self._index(key):
            return
        name = ("Cuda" if cuda and storage.dtype == np.float32 else ("Cuda" if cuda else "") + _CLASS[storage.dtype])
        self._header(f"torch.{name}Tensor")
        self.i32(len(size))
        for s in size: self.long(s)
        for s in stride: self.long(s)
        self.long(offset + 1)
        self.i32(4)
        if self._index(storage):
            return
        self._header(f"torch.{name}Storage")
        self.long(storage.size)
        self.out += storage.astype(storage.dtype.newbyteorder("<")).tobytes()


def dumps(o, **kw):
    w = Writer(**kw)
    w.obj(o)
    return bytes(w.out)
