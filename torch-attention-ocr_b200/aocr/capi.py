"""ctypes binding of include/aocr.h (the Python twin of lua/aocr_ffi.lua)."""
import ctypes as C
import os

import numpy as np

GROUPS = ["cnn", "enc_fw", "enc_bw", "decoder", "proj"]   # src/model/model.lua:150
_HERE = os.path.dirname(os.path.abspath(__file__))


def lib_path():
    return os.environ.get("AOCR_LIB", os.path.join(os.path.dirname(_HERE), "lib", "libaocr.so"))


class AocrError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libaocr error {code}: {msg}")
        self.code = code
        self.msg = msg


class AocrConfig(C.Structure):
    _fields_ = [("batch_size", C.c_int32), ("max_encoder_l", C.c_int32), ("max_decoder_l", C.c_int32),
                ("encoder_num_hidden", C.c_int32), ("encoder_num_layers", C.c_int32),
                ("decoder_num_layers", C.c_int32), ("target_vocab_size", C.c_int32),
                ("target_embedding_size", C.c_int32), ("input_feed", C.c_int32), ("dropout", C.c_float),
                ("learning_rate", C.c_float), ("dp_rank", C.c_int32), ("dp_world", C.c_int32),
                ("global_batch", C.c_int32), ("gemm_mode", C.c_int32)]


_SYMBOLS = {
    "aocr_create": (C.c_int, [C.POINTER(AocrConfig), C.c_int, C.POINTER(C.c_void_p)]),
    "aocr_destroy": (None, [C.c_void_p]),
    "aocr_last_error": (C.c_char_p, [C.c_void_p]),
    "aocr_param_groups": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int64)]),
    "aocr_init_params": (C.c_int, [C.c_void_p, C.c_uint64]),
    "aocr_set_params": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int64]),
    "aocr_get_params": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int64]),
    "aocr_get_grads": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int64]),
    "aocr_set_bn_stats": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64]),
    "aocr_get_bn_stats": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64]),
    "aocr_forward_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                        C.POINTER(C.c_double)]),
    "aocr_group_norms": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "aocr_sgd_update": (C.c_int, [C.c_void_p, C.c_double, C.c_double]),
    "aocr_grad_scale": (C.c_int, [C.c_void_p, C.c_int, C.c_double]),
    "aocr_param_axpy": (C.c_int, [C.c_void_p, C.c_int, C.c_double]),
    "aocr_train_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                  C.c_double, C.POINTER(C.c_double)]),
    "aocr_decode_greedy": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_double),
                                     C.POINTER(C.c_int32)]),
    "aocr_decode_beam": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                   C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_double),
                                   C.POINTER(C.c_int32)]),
    "aocr_trie_load": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(C.POINTER(C.c_int32)), C.POINTER(C.c_int32)]),
    "aocr_trie_from_words": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(C.POINTER(C.c_int32)), C.POINTER(C.c_int32)]),
    "aocr_trie_free": (None, [C.POINTER(C.c_int32)]),
    "aocr_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_int64]),
    "aocr_host_free": (None, [C.c_void_p]),
    "aocr_get_logprobs": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int64]),
    "aocr_debug_read": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64]),
    "aocr_stage_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]),
    "aocr_train_step_staged": (C.c_int, [C.c_void_p, C.c_double, C.c_int, C.POINTER(C.c_double)]),
    "aocr_decode_greedy_staged": (C.c_int, [C.c_void_p, C.c_int]),
    "aocr_grad_buffer": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]),
    "aocr_group_extent": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "aocr_forward_backward_staged": (C.c_int, [C.c_void_p]),
    "aocr_sgd_update_async": (C.c_int, [C.c_void_p, C.c_double, C.c_double]),
    "aocr_read_loss": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    "aocr_stream": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "aocr_set_allreduce": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "aocr_synchronize": (C.c_int, [C.c_void_p]),
    "aocr_launch_count": (C.c_int64, [C.c_void_p]),
    "aocr_dp_unique_id": (C.c_int, [C.c_void_p]),
    "aocr_dp_init": (C.c_int, [C.c_void_p, C.c_void_p]),
    "aocr_last_global_error": (C.c_char_p, []),
    "aocr_set_global_batch": (C.c_int, [C.c_void_p, C.c_int32]),
    "aocr_prof_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "aocr_prof_read": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int64),
                                 C.POINTER(C.c_double)]),
}


_SELFTEST = ("aocr_selftest_gemm", (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                              C.c_void_p, C.c_void_p, C.c_char_p, C.c_int, C.c_int]))


def selftest_gemm(A, B, ta=False, tb=False, mode=0, swap=False, splits=0):
    """C = A @ B through the library's GEMM back ends (test hook; not part of include/aocr.h)."""
    lib = Lib.get()
    fn = getattr(lib.dll, _SELFTEST[0])
    fn.restype, fn.argtypes = _SELFTEST[1]
    M, K = A.shape
    K2, N = B.shape
    assert K == K2
    a = np.ascontiguousarray(A.T if ta else A, dtype=np.float32)
    b = np.ascontiguousarray(B.T if tb else B, dtype=np.float32)
    c = np.empty((M, N), np.float32)
    err = C.create_string_buffer(512)
    rc = fn(M, N, K, int(ta), int(tb), mode, int(swap), _ptr(a), _ptr(b), _ptr(c), err, 512, splits)
    if rc != 0:
        raise AocrError(rc, err.value.decode())
    return c


def exported_symbols():
    return sorted(_SYMBOLS)


class Lib:
    """Loaded libaocr.so with typed prototypes.  Raises if the library is missing (no fallback)."""
    _inst = None

    def __init__(self, path=None):
        path = path or lib_path()
        if not os.path.exists(path):
            raise FileNotFoundError(
                f"{path} not found: build it with `python __graft_entry__.py` / `make -C torch-attention-ocr_b200` "
                "(the hot path has no CPU or PyTorch fallback)")
        self.path = path
        self.dll = C.CDLL(path)
        for name, (res, args) in _SYMBOLS.items():
            fn = getattr(self.dll, name)
            fn.restype = res
            fn.argtypes = args

    @classmethod
    def get(cls):
        if cls._inst is None:
            cls._inst = Lib()
        return cls._inst

    def dp_unique_id(self) -> bytes:
        """128-byte NCCL unique id (call on rank 0, ship to every rank, then Handle.dp_init on all of them)"""
        buf = C.create_string_buffer(128)
        rc = self.dll.aocr_dp_unique_id(buf)
        if rc != 0:
            raise RuntimeError("aocr_dp_unique_id failed: " + self.dll.aocr_last_global_error().decode())
        return buf.raw


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class _PinnedBlock:
    """owner of one aocr_host_alloc block.  Arrays are made over `buf`; `buf` holds a reference to its owner, so the
    block is returned (aocr_host_free) only when the last array over it is gone."""

    def __init__(self, nbytes):
        import weakref
        self.lib = Lib.get()
        p = C.c_void_p()
        rc = self.lib.dll.aocr_host_alloc(C.byref(p), int(nbytes))
        if rc != 0:
            raise AocrError(rc, self.lib.dll.aocr_last_global_error().decode())
        self.ptr, self.nbytes = p.value, int(nbytes)
        self.buf = (C.c_ubyte * self.nbytes).from_address(self.ptr)
        self.buf._owner = self
        _PINNED_OWNERS[self.ptr] = weakref.ref(self, lambda _r, k=self.ptr: _PINNED_OWNERS.pop(k, None))

    def array(self, shape, dtype):
        dtype = np.dtype(dtype)
        n = int(np.prod(shape, dtype=np.int64))
        assert n * dtype.itemsize <= self.nbytes
        return np.frombuffer(self.buf, dtype=dtype, count=n).reshape(shape)

    def __del__(self):
        try:
            if self.ptr:
                self.lib.dll.aocr_host_free(C.c_void_p(self.ptr))
                self.ptr = None
        except Exception:
            pass


_PINNED_OWNERS = {}     # address of a live block -> weak reference to its owner (introspection: is_pinned)


def _nbytes(shape, dtype):
    return max(int(np.prod(shape, dtype=np.int64)) * np.dtype(dtype).itemsize, 1)


def host_empty(shape, dtype=np.float32):
    """np.empty in page-locked host memory (aocr_host_alloc): what the data layer allocates batch tensors from so that
    the library's host-to-device copies are asynchronous.  Raises AocrError when no CUDA device is present."""
    return _PinnedBlock(_nbytes(shape, dtype)).array(shape, dtype)


def is_pinned(a):
    """True when the array's memory lies inside a live aocr_host_alloc block"""
    addr = a.ctypes.data
    for ptr, ref in list(_PINNED_OWNERS.items()):
        blk = ref()
        if blk is not None and blk.ptr is not None and ptr <= addr < ptr + blk.nbytes:
            return True
    return False


class PinnedRing:
    """Allocator of batch tensors for the data layer (`DataGen(..., alloc=PinnedRing(depth))`): `depth` generations of
    page-locked blocks, recycled round-robin — page-locking is expensive (and releasing it synchronises the device), so
    blocks are reused instead of allocated per batch.  `next_batch()` opens the next generation; each call then hands
    out that generation's next block (grown when a batch needs more).  A batch's arrays are overwritten `depth` batches
    later: depth must exceed the number of batches alive at once (prefetch queue + the one being built + the one the
    step reads: prefetch + 2)."""

    def __init__(self, depth=6):
        assert depth >= 2
        self.depth = depth
        self.generations = [[] for _ in range(depth)]
        self.cur, self.k = -1, 0
        self.allocations = 0          # page-locking calls made so far (tests: stays flat once warm)

    def next_batch(self):
        self.cur = (self.cur + 1) % self.depth
        self.k = 0

    def __call__(self, shape, dtype=np.float32):
        if self.cur < 0:
            self.next_batch()
        gen, i = self.generations[self.cur], self.k
        self.k += 1
        need = _nbytes(shape, dtype)
        if i >= len(gen):
            gen.append(None)
        if gen[i] is None or gen[i].nbytes < need:
            grown = need if gen[i] is None else max(need, gen[i].nbytes * 3 // 2)
            gen[i] = _PinnedBlock(grown)        # arrays still alive over the old block keep it until they die
            self.allocations += 1
        return gen[i].array(shape, dtype)


class Trie:
    """Dictionary trie of the constrained decode: loadDictionary (src/utils/utils.lua:177-218) as the flat child table the
    library consumes.  Trie(path=...) reads one word per line; Trie(words=[...]) takes them from memory."""

    def __init__(self, path=None, words=None, allow_digit_prefix=False):
        lib = Lib.get()
        self.lib = lib
        self.table = C.POINTER(C.c_int32)()
        n = C.c_int32()
        if path is not None:
            rc = lib.dll.aocr_trie_load(os.fsencode(path), int(bool(allow_digit_prefix)), C.byref(self.table), C.byref(n))
        else:
            rc = lib.dll.aocr_trie_from_words("\n".join(words).encode(), int(bool(allow_digit_prefix)), C.byref(self.table),
                                              C.byref(n))
        if rc != 0:
            raise AocrError(rc, "dictionary could not be read" if path else "bad dictionary word")
        self.num_nodes = n.value

    def numpy(self, V=39):
        return np.ctypeslib.as_array(self.table, shape=(self.num_nodes, V + 1)).copy()

    def __del__(self):
        try:
            if self.table:
                self.lib.dll.aocr_trie_free(self.table)
                self.table = C.POINTER(C.c_int32)()
        except Exception:
            pass


class Handle:
    """RAII wrapper of one aocr_handle; methods map 1:1 onto the C ABI."""

    def __init__(self, cfg: AocrConfig, device=0):
        self.lib = Lib.get()
        self.h = C.c_void_p()
        rc = self.lib.dll.aocr_create(C.byref(cfg), device, C.byref(self.h))
        if rc != 0:
            raise AocrError(rc, self.lib.dll.aocr_last_error(None).decode())
        self.cfg = cfg
        n = C.c_int32()
        sizes = (C.c_int64 * 5)()
        self._ck(self.lib.dll.aocr_param_groups(self.h, C.byref(n), sizes))
        self.group_sizes = [int(s) for s in sizes]

    def _ck(self, rc):
        if rc != 0:
            raise AocrError(rc, self.lib.dll.aocr_last_error(self.h).decode())

    def close(self):
        if self.h:
            self.lib.dll.aocr_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- parameters
    def init_params(self, seed=910820):
        """fresh parameters with Torch7's reset() distributions (what Model:create's module constructors draw)"""
        self._ck(self.lib.dll.aocr_init_params(self.h, int(seed) & 0xFFFFFFFFFFFFFFFF))

    def set_global_batch(self, n):
        self._ck(self.lib.dll.aocr_set_global_batch(self.h, int(n)))

    def set_params(self, group, arr):
        a = np.ascontiguousarray(arr, dtype=np.float32)
        self._ck(self.lib.dll.aocr_set_params(self.h, group, _ptr(a), a.size))

    def get_params(self, group):
        a = np.empty(self.group_sizes[group], np.float32)
        self._ck(self.lib.dll.aocr_get_params(self.h, group, _ptr(a), a.size))
        return a

    def get_grads(self, group):
        a = np.empty(self.group_sizes[group], np.float32)
        self._ck(self.lib.dll.aocr_get_grads(self.h, group, _ptr(a), a.size))
        return a

    def set_bn_stats(self, layer, mean, var):
        m = np.ascontiguousarray(mean, dtype=np.float32)
        v = np.ascontiguousarray(var, dtype=np.float32)
        self._ck(self.lib.dll.aocr_set_bn_stats(self.h, layer, _ptr(m), _ptr(v), m.size))

    def get_bn_stats(self, layer):
        n = 256 if layer == 0 else 512
        m, v = np.empty(n, np.float32), np.empty(n, np.float32)
        self._ck(self.lib.dll.aocr_get_bn_stats(self.h, layer, _ptr(m), _ptr(v), n))
        return m, v

    # ---- steps
    @staticmethod
    def _batch(images, targets, targets_eval):
        img = np.ascontiguousarray(images, dtype=np.float32)
        tg = np.ascontiguousarray(targets, dtype=np.int32)
        te = np.ascontiguousarray(targets_eval, dtype=np.int32)
        assert img.ndim == 4 and img.shape[1] == 1 and img.shape[2] == 32, "images must be (b,1,32,W)"
        assert tg.shape == te.shape and tg.shape[0] == img.shape[0]
        return img, tg, te, img.shape[0], img.shape[3], tg.shape[1]

    def forward_backward(self, images, targets, targets_eval):
        img, tg, te, b, W, T = self._batch(images, targets, targets_eval)
        loss = C.c_double()
        self._ck(self.lib.dll.aocr_forward_backward(self.h, _ptr(img), b, W, _ptr(tg), _ptr(te), T, C.byref(loss)))
        return loss.value

    def train_step(self, images, targets, targets_eval, lr):
        img, tg, te, b, W, T = self._batch(images, targets, targets_eval)
        loss = C.c_double()
        self._ck(self.lib.dll.aocr_train_step(self.h, _ptr(img), b, W, _ptr(tg), _ptr(te), T, lr, C.byref(loss)))
        return loss.value

    def group_norms(self):
        pn, gn = (C.c_double * 5)(), (C.c_double * 5)()
        self._ck(self.lib.dll.aocr_group_norms(self.h, pn, gn))
        return list(pn), list(gn)

    def sgd_update(self, lr, clip=5.0):
        self._ck(self.lib.dll.aocr_sgd_update(self.h, lr, clip))

    def grad_scale(self, group, s):
        self._ck(self.lib.dll.aocr_grad_scale(self.h, group, s))

    def param_axpy(self, group, a):
        self._ck(self.lib.dll.aocr_param_axpy(self.h, group, a))

    def decode_greedy(self, images, targets, targets_eval):
        img, tg, te, b, W, T = self._batch(images, targets, targets_eval)
        Ld = self.cfg.max_decoder_l
        labels = np.empty((b, Ld), np.int32)
        pred, gold = np.empty(b, np.float64), np.empty(b, np.float64)
        loss, nc = C.c_double(), C.c_int32()
        self._ck(self.lib.dll.aocr_decode_greedy(self.h, _ptr(img), b, W, _ptr(tg), _ptr(te), T, _ptr(labels),
                                                 _ptr(pred), _ptr(gold), C.byref(loss), C.byref(nc)))
        return {"labels": labels, "pred_scores": pred, "gold_scores": gold, "loss_sum": loss.value,
                "num_correct": nc.value}

    def decode_beam(self, images, targets, targets_eval, beam_size, trie=None):
        """beam search, optionally constrained to a dictionary (`trie`: a Trie)"""
        img, tg, te, b, W, T = self._batch(images, targets, targets_eval)
        Ld = self.cfg.max_decoder_l
        labels = np.empty((b, Ld), np.int32)
        pred, gold = np.empty(b, np.float64), np.empty(b, np.float64)
        loss, nc = C.c_double(), C.c_int32()
        self._ck(self.lib.dll.aocr_decode_beam(self.h, _ptr(img), b, W, _ptr(tg), _ptr(te), T, int(beam_size),
                                               C.cast(trie.table, C.c_void_p) if trie is not None else None,
                                               trie.num_nodes if trie is not None else 0, _ptr(labels), _ptr(pred),
                                               _ptr(gold), C.byref(loss), C.byref(nc)))
        return {"labels": labels, "pred_scores": pred, "gold_scores": gold, "loss_sum": loss.value,
                "num_correct": nc.value}

    def get_logprobs(self, which, rows):
        a = np.empty((rows, self.cfg.target_vocab_size), np.float32)
        self._ck(self.lib.dll.aocr_get_logprobs(self.h, which, _ptr(a), a.size))
        return a

    def debug_read(self, name, shape):
        a = np.empty(shape, np.float32)
        self._ck(self.lib.dll.aocr_debug_read(self.h, name.encode(), _ptr(a), a.size))
        return a

    # ---- device-resident path
    def stage_batch(self, images, targets, targets_eval):
        img, tg, te, b, W, T = self._batch(images, targets, targets_eval)
        self._ck(self.lib.dll.aocr_stage_batch(self.h, _ptr(img), b, W, _ptr(tg), _ptr(te), T))

    def train_step_staged(self, lr, sync=True):
        loss = C.c_double()
        self._ck(self.lib.dll.aocr_train_step_staged(self.h, lr, 1 if sync else 0, C.byref(loss)))
        return loss.value if sync else None

    def decode_greedy_staged(self, sync=True):
        self._ck(self.lib.dll.aocr_decode_greedy_staged(self.h, 1 if sync else 0))

    def forward_backward_staged(self):
        self._ck(self.lib.dll.aocr_forward_backward_staged(self.h))

    def sgd_update_async(self, lr, clip=5.0):
        self._ck(self.lib.dll.aocr_sgd_update_async(self.h, lr, clip))

    def read_loss(self):
        loss = C.c_double()
        self._ck(self.lib.dll.aocr_read_loss(self.h, C.byref(loss)))
        return loss.value

    def grad_buffer(self):
        p, n = C.c_void_p(), C.c_int64()
        self._ck(self.lib.dll.aocr_grad_buffer(self.h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def group_extent(self, group):
        o, n = C.c_int64(), C.c_int64()
        self._ck(self.lib.dll.aocr_group_extent(self.h, group, C.byref(o), C.byref(n)))
        return o.value, n.value

    def stream(self):
        p = C.c_void_p()
        self._ck(self.lib.dll.aocr_stream(self.h, C.byref(p)))
        return p.value or 0

    def set_allreduce(self, cb):
        self._ar_cb = cb     # keep the ctypes callback alive
        self._ck(self.lib.dll.aocr_set_allreduce(self.h, C.cast(cb, C.c_void_p), None))

    def synchronize(self):
        self._ck(self.lib.dll.aocr_synchronize(self.h))

    def launch_count(self):
        return int(self.lib.dll.aocr_launch_count(self.h))

    def dp_init(self, unique_id: bytes):
        """native NCCL exchange (collective over the handle's dp_world); `unique_id` from Lib.dp_unique_id() of rank 0"""
        assert len(unique_id) == 128
        buf = C.create_string_buffer(unique_id, 128)
        self._ck(self.lib.dll.aocr_dp_init(self.h, buf))

    def prof_enable(self, on=True):
        self._ck(self.lib.dll.aocr_prof_enable(self.h, 1 if on else 0))

    def prof_read(self, cls):
        ms, n, w = C.c_double(), C.c_int64(), C.c_double()
        self._ck(self.lib.dll.aocr_prof_read(self.h, cls, C.byref(ms), C.byref(n), C.byref(w)))
        return ms.value, n.value, w.value
