"""Torch7 binary serialisation (`torch.save` / `torch.load`, torch7 File.lua + DiskFile in binary mode, little endian,
64-bit `long`) — the container of the reference's checkpoints (src/model/model.lua:45-80 load, :720-725 save).

Object stream: every object starts with an int32 type tag
    0 nil | 1 number (float64) | 2 string (int32 length + bytes) | 3 table | 4 torch object | 5 boolean (int32)
    6 function (legacy) | 7, 8 function with reference index
Tables and torch objects carry an int32 index; an index seen before is a back reference (shared storages, the shared
weights of cloned modules, cyclic graph nodes).  A table is `int32 n` followed by n (key, value) objects.  A torch object
is the version string "V 1", its class name, then the class's own `write`: tensors and storages as below, any other class
(nn modules, nngraph nodes) as a table of its fields.
    torch.*Tensor : int32 ndim, int64 size[ndim], int64 stride[ndim], int64 storage_offset (1-based), storage object
    torch.*Storage: int64 n, n raw elements

`load` returns plain Python values: tables become `dict` (or `list` when the keys are 1..n), tensors numpy arrays,
other torch objects `T7Object(cls, fields)`.  `save` writes the same shapes back (dict / list / numpy / T7Object /
scalars), sharing objects that are the same Python object.  Functions are read (and dropped) but never written.
"""
import struct

import numpy as np

TYPE_NIL, TYPE_NUMBER, TYPE_STRING, TYPE_TABLE, TYPE_TORCH, TYPE_BOOLEAN, TYPE_FUNCTION = 0, 1, 2, 3, 4, 5, 6
TYPE_RECUR_FUNCTION, LEGACY_TYPE_RECUR_FUNCTION = 8, 7

_DTYPES = {"Double": np.float64, "Float": np.float32, "Half": np.float16, "Long": np.int64, "Int": np.int32,
           "Short": np.int16, "Char": np.int8, "Byte": np.uint8}
_NAMES = {np.dtype(v): k for k, v in _DTYPES.items()}


class T7Object:
    """a torch class instance that is not a tensor / storage: class name + its fields (a dict)"""

    def __init__(self, cls, fields=None):
        self.cls = cls
        self.fields = fields if fields is not None else {}

    def __getitem__(self, k):
        return self.fields[k]

    def get(self, k, default=None):
        return self.fields.get(k, default)

    def __repr__(self):
        return f"T7Object({self.cls}, {list(self.fields)[:8]})"


class _TableKey:
    """identity-hashed wrapper for tables used as table KEYS (nngraph's mapindex maps node data tables to indices)"""

    def __init__(self, obj):
        self.obj = obj

    def __hash__(self):
        return id(self.obj)

    def __eq__(self, other):
        return isinstance(other, _TableKey) and other.obj is self.obj


def _base_class(cls):
    # cutorch tensors / storages ("torch.CudaTensor", "torch.CudaDoubleTensor") hold the same payload as the host classes
    name = cls[len("torch."):]
    kind = "Tensor" if name.endswith("Tensor") else ("Storage" if name.endswith("Storage") else None)
    if kind is None:
        return None, None
    t = name[:-len(kind)]
    if t.startswith("Cuda"):
        t = t[4:] or "Float"
    return kind, _DTYPES.get(t)


class _Reader:
    def __init__(self, f):
        self.f = f
        self.memo = {}

    def _rd(self, fmt):
        n = struct.calcsize(fmt)
        b = self.f.read(n)
        if len(b) != n:
            raise EOFError("truncated Torch7 file")
        return struct.unpack("<" + fmt, b)[0]

    def _string(self):
        n = self._rd("i")
        return self.f.read(n).decode("latin-1")

    def obj(self):
        t = self._rd("i")
        if t == TYPE_NIL:
            return None
        if t == TYPE_NUMBER:
            v = self._rd("d")
            return int(v) if (v == v and abs(v) < 2 ** 53 and v == int(v)) else v
        if t == TYPE_STRING:
            return self._string()
        if t == TYPE_BOOLEAN:
            return self._rd("i") == 1
        if t == TYPE_FUNCTION:
            n = self._rd("i")
            self.f.read(n)
            self.obj()                      # upvalues
            return None
        if t in (TYPE_RECUR_FUNCTION, LEGACY_TYPE_RECUR_FUNCTION):
            idx = self._rd("i")
            if idx in self.memo:
                return self.memo[idx]
            self.memo[idx] = None
            n = self._rd("i")
            self.f.read(n)
            self.obj()
            return None
        if t == TYPE_TABLE:
            idx = self._rd("i")
            if idx in self.memo:
                return self.memo[idx]
            d = {}
            self.memo[idx] = d
            n = self._rd("i")
            for _ in range(n):
                k = self.obj()
                v = self.obj()
                if isinstance(k, (dict, list, T7Object)):
                    k = _TableKey(k)
                d[k] = v
            return d
        if t == TYPE_TORCH:
            idx = self._rd("i")
            if idx in self.memo:
                return self.memo[idx]
            ver = self._string()
            cls = self._string() if ver.startswith("V ") else ver      # files older than "V 1" have no version string
            kind, dt = _base_class(cls) if cls.startswith("torch.") else (None, None)
            if kind == "Storage":
                n = self._rd("q")
                a = np.frombuffer(self.f.read(n * np.dtype(dt).itemsize), dtype=dt).copy()
                self.memo[idx] = a
                return a
            if kind == "Tensor":
                nd = self._rd("i")
                size = [self._rd("q") for _ in range(nd)]
                stride = [self._rd("q") for _ in range(nd)]
                off = self._rd("q") - 1
                holder = [None]
                self.memo[idx] = holder          # placeholder (a tensor cannot contain itself)
                st = self.obj()
                if st is None or nd == 0:
                    a = np.zeros([0] * max(nd, 1), dtype=dt)
                else:
                    a = np.lib.stride_tricks.as_strided(st[off:], shape=size, strides=[s * st.itemsize for s in stride]).copy()
                self.memo[idx] = a
                return a
            o = T7Object(cls)
            self.memo[idx] = o
            body = self.obj()
            o.fields = body if isinstance(body, dict) else {"_value": body}
            return o
        raise ValueError(f"unknown Torch7 type tag {t}")


def _listify(v, seen=None):
    """tables whose keys are exactly 1..n become lists (recursively, cycle safe)"""
    seen = {} if seen is None else seen
    if id(v) in seen:
        return seen[id(v)]
    if isinstance(v, T7Object):
        seen[id(v)] = v
        v.fields = {k: _listify(x, seen) for k, x in v.fields.items()}
        return v
    if isinstance(v, dict):
        n = len(v)
        if n > 0 and all(isinstance(k, int) for k in v) and set(v) == set(range(1, n + 1)):
            out = []
            seen[id(v)] = out
            out.extend(_listify(v[i], seen) for i in range(1, n + 1))
            return out
        seen[id(v)] = v
        for k in list(v):
            v[k] = _listify(v[k], seen)
        return v
    return v


def load(path, listify=True):
    with open(path, "rb") as f:
        v = _Reader(f).obj()
    return _listify(v) if listify else v


class _Writer:
    def __init__(self, f):
        self.f = f
        self.memo = {}
        self.keep = []          # keeps written objects alive so ids stay unique
        self.next = 1

    def _w(self, fmt, v):
        self.f.write(struct.pack("<" + fmt, v))

    def _string(self, s):
        b = s.encode("latin-1")
        self._w("i", len(b))
        self.f.write(b)

    def _index(self, o):
        """returns True when `o` was written before (a back reference was emitted)"""
        if id(o) in self.memo:
            self._w("i", self.memo[id(o)])
            return True
        self.memo[id(o)] = self.next
        self.keep.append(o)
        self._w("i", self.next)
        self.next += 1
        return False

    def _torch_header(self, cls):
        self._string("V 1")
        self._string(cls)

    def obj(self, v):
        if v is None:
            self._w("i", TYPE_NIL)
        elif isinstance(v, (bool, np.bool_)):
            self._w("i", TYPE_BOOLEAN)
            self._w("i", 1 if v else 0)
        elif isinstance(v, (int, float, np.integer, np.floating)):
            self._w("i", TYPE_NUMBER)
            self._w("d", float(v))
        elif isinstance(v, str):
            self._w("i", TYPE_STRING)
            self._string(v)
        elif isinstance(v, np.ndarray):
            self._w("i", TYPE_TORCH)
            if self._index(v):
                return
            name = _NAMES[v.dtype]
            a = np.ascontiguousarray(v)
            self._torch_header(f"torch.{name}Tensor")
            self._w("i", a.ndim)
            for s in a.shape:
                self._w("q", s)
            for s in a.strides:
                self._w("q", s // a.itemsize)
            self._w("q", 1)
            if a.size == 0:
                self._w("i", TYPE_NIL)
                return
            self._w("i", TYPE_TORCH)             # its storage: a fresh object
            self._w("i", self.next)
            self.next += 1
            self._torch_header(f"torch.{name}Storage")
            self._w("q", a.size)
            self.f.write(a.tobytes())
        elif isinstance(v, T7Object):
            self._w("i", TYPE_TORCH)
            if self._index(v):
                return
            self._torch_header(v.cls)
            self._table_body(v.fields, fresh=True)
        elif isinstance(v, (list, tuple)):
            self._table(v, {i + 1: x for i, x in enumerate(v)})
        elif isinstance(v, dict):
            self._table(v, v)
        else:
            raise TypeError(f"cannot serialise {type(v)} to Torch7")

    def _table(self, ident, d):
        self._w("i", TYPE_TABLE)
        if self._index(ident):
            return
        self._w("i", len(d))
        for k, x in d.items():
            self.obj(k.obj if isinstance(k, _TableKey) else k)
            self.obj(x)

    def _table_body(self, d, fresh):
        # the field table of a torch object is a table object of its own
        self._w("i", TYPE_TABLE)
        self._w("i", self.next)
        self.next += 1
        self._w("i", len(d))
        for k, x in d.items():
            self.obj(k.obj if isinstance(k, _TableKey) else k)
            self.obj(x)


def save(path, value):
    with open(path, "wb") as f:
        _Writer(f).obj(value)
