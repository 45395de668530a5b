"""The reference's data layer (src/data/data_gen.lua): `DataGen` reads "image_path label" lines, decodes each image,
converts it to luma, rescales it to height 32 and buckets by width; `nextBatch` returns the batch tuple
`{images, targets, targets_eval, num_nonzeros, img_paths}` that `Model.step` consumes.  `SyntheticDataGen` emits the
same tuple from seeded random images / labels (benchmarks, tests).
"""
import os
import queue
import random
import threading

import numpy as np

ALPHABET = "0123456789abcdefghijklmnopqrstuvwxyz"


def str2numlist(s):
    """src/utils/utils.lua:104-118"""
    return [2] + [(ord(c) - 97 + 14) if ord(c) > 96 else (ord(c) - 48 + 4) for c in s] + [3]


def numlist2str(ids):
    """src/utils/utils.lua:120-134"""
    return "".join(chr(v - 14 + 97) if v > 13 else chr(v - 4 + 48) for v in ids)


def make_batch_from_labels(images, labels, img_paths=None):
    """images (b,1,32,W); labels: list of strings -> the reference batch tuple (as a list, 1:1 with Lua)."""
    lists = [str2numlist(s) for s in labels]
    b = len(lists)
    T = max(len(l) for l in lists) - 1
    targets = np.ones((b, T), np.int32)          # data_gen.lua:107
    targets_eval = np.ones((b, T), np.int32)     # data_gen.lua:109
    nnz = 0
    for i, l in enumerate(lists):
        nnz += len(l) - 1                        # data_gen.lua:112
        targets[i, :len(l) - 1] = l[:-1]
        targets_eval[i, :len(l) - 1] = l[1:]
    return [np.ascontiguousarray(images, dtype=np.float32), targets, targets_eval, nnz,
            img_paths or [f"synthetic/{i}.png" for i in range(b)]]


def synthetic_batch(b, W, max_label_len, seed=910820, force_T=None):
    """One seeded synthetic batch of the benchmark workloads: i.i.d. uniform gray levels 0..255, labels of
    1..max_label_len characters over [0-9a-z]; `force_T` pins the target length (the longest label gets force_T - 1
    characters).  Returns a dict with the arrays of the reference batch tuple."""
    rng = np.random.Generator(np.random.PCG64(seed))
    images = rng.integers(0, 256, size=(b, 1, 32, W)).astype(np.float32)
    lens = rng.integers(1, max_label_len + 1, size=b)
    if force_T is not None:
        lens[0] = force_T - 1
        lens = np.minimum(lens, force_T - 1)
    labels = ["".join(ALPHABET[i] for i in rng.integers(0, 36, size=int(n))) for n in lens]
    t = make_batch_from_labels(images, labels)
    return {"images": t[0], "targets": t[1], "targets_eval": t[2], "num_nonzeros": int(t[3]), "labels": labels}


class SyntheticDataGen:
    """Width-bucketed synthetic stand-in for DataGen (same nextBatch contract, incl. the final partial flush)."""

    def __init__(self, num_samples, widths=(100,), max_label_len=19, seed=910820, imgH=32):
        self.rng = np.random.Generator(np.random.PCG64(seed))
        self.num_samples, self.widths, self.max_label_len, self.imgH = num_samples, list(widths), max_label_len, imgH
        self.cursor = 0
        self.buffer = {}

    def size(self):
        return self.num_samples

    def shuffle(self):
        pass

    def _sample(self):
        W = int(self.widths[self.rng.integers(0, len(self.widths))])
        img = self.rng.integers(0, 256, size=(1, self.imgH, W)).astype(np.float32)
        n = int(self.rng.integers(1, self.max_label_len + 1))
        return W, img, "".join(ALPHABET[i] for i in self.rng.integers(0, 36, size=n))

    def nextBatch(self, batch_size):
        while self.cursor < self.num_samples:
            W, img, label = self._sample()
            self.cursor += 1
            self.buffer.setdefault(W, []).append((img, label))
            if len(self.buffer[W]) == batch_size:
                items = self.buffer.pop(W)
                return make_batch_from_labels(np.stack([i for i, _ in items]), [l for _, l in items])
        if not self.buffer:
            self.cursor = 0
            return None
        W = next(iter(self.buffer))
        items = self.buffer.pop(W)
        return make_batch_from_labels(np.stack([i for i, _ in items]), [l for _, l in items])


# ---- real data path (src/data/data_gen.lua:14-154) --------------------------------------------------------------
def rgb2y(img):
    """image.rgb2y [T7 image package]: luma 0.299 R + 0.587 G + 0.114 B of a (3,H,W) image in [0,1]; 1 channel: as is"""
    if img.shape[0] == 1:
        return img
    return (0.299 * img[0] + 0.587 * img[1] + 0.114 * img[2])[None]


def _scale_linear_1d(src, dst_len):
    """one axis (the last) of image.scale(..., 'bilinear') [T7 image package, image.c: scaleLinear_rowcol]: stretching
    interpolates between the two neighbours with end points aligned; shrinking averages the source interval of each
    destination pixel (fractional weights at its two ends)."""
    src_len = src.shape[-1]
    if dst_len == src_len:
        return src.copy()
    out = np.empty(src.shape[:-1] + (dst_len,), src.dtype)
    if dst_len > src_len:
        if src_len == 1:
            out[...] = src
            return out
        pos = np.arange(dst_len - 1, dtype=np.float32) * np.float32((src_len - 1) / (dst_len - 1))
        i0 = pos.astype(np.int64)
        f = (pos - i0).astype(src.dtype)
        out[..., :-1] = (1 - f) * src[..., i0] + f * src[..., i0 + 1]
        out[..., -1] = src[..., -1]
        return out
    scale = np.float32(src_len / dst_len)
    s0_i, s0_f = 0, 0.0
    for di in range(dst_len):
        s1 = np.float32(di + 1) * scale
        s1_i = int(s1)
        s1_f = float(s1 - s1_i)
        acc = (1 - s0_f) * src[..., s0_i]
        n = 1 - s0_f
        if s1_i > s0_i + 1:
            acc = acc + src[..., s0_i + 1:s1_i].sum(axis=-1)
            n += s1_i - s0_i - 1
        if s1_i < src_len:
            acc = acc + s1_f * src[..., s1_i]
            n += s1_f
        out[..., di] = acc / n
        s0_i, s0_f = s1_i, s1_f
    return out


def scale_bilinear(img, width, height):
    """image.scale(img, width, height) in its default 'bilinear' mode: rows first, then columns"""
    tmp = _scale_linear_1d(np.asarray(img, np.float32), width)                    # (C,H,W) -> (C,H,width)
    return np.ascontiguousarray(_scale_linear_1d(tmp.transpose(0, 2, 1), height).transpose(0, 2, 1))


def _read_pnm(path):
    """binary / ASCII PGM and PPM (P2, P3, P5, P6; 8 or 16 bit): (C,H,W) float in [0,1]"""
    with open(path, "rb") as f:
        data = f.read()
    pos, fields = 0, []
    while len(fields) < 4:                       # magic, width, height, maxval — whitespace separated, '#' comments
        while pos < len(data) and data[pos:pos + 1].isspace():
            pos += 1
        if data[pos:pos + 1] == b"#":
            while pos < len(data) and data[pos:pos + 1] != b"\n":
                pos += 1
            continue
        start = pos
        while pos < len(data) and not data[pos:pos + 1].isspace():
            pos += 1
        if start == pos:
            raise ValueError("truncated PNM header: %s" % path)
        fields.append(data[start:pos])
    magic, w, h, maxval = fields[0], int(fields[1]), int(fields[2]), int(fields[3])
    if magic not in (b"P2", b"P3", b"P5", b"P6") or not (0 < maxval < 65536):
        raise ValueError("unsupported PNM file: %s" % path)
    c = 3 if magic in (b"P3", b"P6") else 1
    if magic in (b"P5", b"P6"):
        pos += 1                                 # exactly one whitespace byte after maxval
        dt = np.dtype(">u2") if maxval > 255 else np.dtype("u1")
        a = np.frombuffer(data, dtype=dt, count=w * h * c, offset=pos)
    else:
        a = np.array(data[pos:].split()[:w * h * c], dtype=np.int64)
    if a.size != w * h * c:
        raise ValueError("truncated PNM file: %s" % path)
    a = a.reshape(h, w, c).astype(np.float32) / np.float32(maxval)
    return np.ascontiguousarray(a.transpose(2, 0, 1))


def _from_array(a):
    """a decoded array -> (C,H,W) float in [0,1]: (H,W), (H,W,3|4) or (1|3,H,W); integer types are scaled by their range"""
    a = np.asarray(a)
    scale = np.float32(1.0 / np.iinfo(a.dtype).max) if np.issubdtype(a.dtype, np.integer) else np.float32(1.0)
    a = a.astype(np.float32) * scale
    if a.ndim == 2:
        return a[None]
    if a.ndim == 3 and a.shape[0] in (1, 3) and a.shape[2] not in (1, 3, 4):
        return np.ascontiguousarray(a)
    if a.ndim == 3 and a.shape[2] in (1, 3, 4):
        return np.ascontiguousarray(a[:, :, :3].transpose(2, 0, 1)) if a.shape[2] >= 3 else np.ascontiguousarray(a.transpose(2, 0, 1))
    raise ValueError("cannot interpret an array of shape %s as an image" % (a.shape,))


def load_image(path):
    """image.load: (C,H,W) float in [0,1]; raises when the file cannot be decoded (data_gen.lua:65 uses pcall).
    `.npy` arrays and PGM / PPM files are read here; every other format (JPEG, PNG, ...) goes through PIL."""
    ext = os.path.splitext(path)[1].lower()
    if ext == ".npy":
        return _from_array(np.load(path, allow_pickle=False))
    if ext in (".pgm", ".ppm", ".pnm"):
        return _read_pnm(path)
    try:
        from PIL import Image
    except ImportError as e:
        raise RuntimeError("decoding %s needs PIL (only .npy / .pgm / .ppm are read without it)" % path) from e
    with Image.open(path) as im:
        if im.mode not in ("L", "RGB"):
            im = im.convert("RGB")
        a = np.asarray(im, dtype=np.float32) / 255.0
    return a[None] if a.ndim == 2 else np.ascontiguousarray(a.transpose(2, 0, 1))


class DataGen:
    """DataGen(data_base_dir, data_path, max_aspect_ratio) — src/data/data_gen.lua:14-154.

    `fixed_width`: the reference overrides every computed width with 100 (data_gen.lua:78), which
    makes its width bucketing dormant; fixed_width=100 reproduces that, fixed_width=None (default here) keeps the width
    ceil(aspect_ratio * 32) the line above computes, i.e. true bucketing.  `prefetch` > 0 decodes ahead on a worker
    thread so the device step does not wait for image I/O; `alloc` (e.g. `capi.PinnedRing`) supplies the batch tensors,
    page-locked so the library's input copies are asynchronous transfers."""

    def __init__(self, data_base_dir, data_path, max_aspect_ratio, fixed_width=None, log=print, seed=None, prefetch=0,
                 alloc=None):
        self.imgH = 32
        self.alloc = alloc           # (shape, dtype) -> ndarray for the batch tensors; capi.PinnedRing: page-locked memory
        self.data_base_dir, self.data_path = data_base_dir, data_path
        self.max_aspect_ratio, self.min_aspect_ratio = max_aspect_ratio, 0.5
        self.fixed_width = fixed_width
        self.log = log
        path = data_path if os.path.isfile(data_path) else os.path.join(data_base_dir, data_path)
        if not os.path.isfile(path):                                             # data_gen.lua:29-36
            raise FileNotFoundError("Error: Data file %s not found " % data_path)
        self.lines = []
        with open(path) as f:
            for idx, line in enumerate(f, 1):
                if idx % 1000000 == 0:
                    log("%d lines read" % idx)
                parts = line.split()
                if len(parts) >= 2:
                    self.lines.append([parts[0], parts[1], None, None])          # filename, label, image, id list
        self.cursor = 0
        self.buffer = {}
        self.rng = random.Random(seed)
        self.prefetch = prefetch
        self._q, self._worker, self._worker_bs = None, None, None

    def shuffle(self):                                                           # utils.lua:12-19 (Fisher-Yates)
        self._stop_worker()
        a = self.lines
        for counter in range(len(a), 1, -1):
            index = self.rng.randrange(counter)
            a[index], a[counter - 1] = a[counter - 1], a[index]

    def size(self):
        return len(self.lines)

    def _decode(self, rec):
        try:
            img = load_image(os.path.join(self.data_base_dir, rec[0]))            # data_gen.lua:65
        except Exception:
            return False
        img = 255.0 * rgb2y(img)                                                  # :69
        origH, origW = img.shape[1], img.shape[2]
        ar = min(max(origW / origH, self.min_aspect_ratio), self.max_aspect_ratio)   # :72-75
        imgW = int(np.ceil(ar * self.imgH))                                       # :76
        if self.fixed_width is not None:
            imgW = self.fixed_width                                               # :78
        rec[2] = scale_bilinear(img, imgW, self.imgH)                             # :79
        rec[3] = str2numlist(rec[1])                                              # :68
        return True

    def _batch(self, items, imgW):                                                # data_gen.lua:97-120 / 132-153
        b = len(items)
        alloc = self.alloc or np.empty
        if hasattr(alloc, "next_batch"):
            alloc.next_batch()
        images = alloc((b, 1, self.imgH, imgW), np.float32)
        for i, it in enumerate(items):
            images[i] = it[0]
        T = max(len(it[1]) for it in items) - 1
        targets = alloc((b, T), np.int32)
        targets_eval = alloc((b, T), np.int32)
        targets[...] = 1
        targets_eval[...] = 1
        nnz = 0
        for i, it in enumerate(items):
            l = it[1]
            nnz += len(l) - 1
            targets[i, :len(l) - 1] = l[:-1]
            targets_eval[i, :len(l) - 1] = l[1:]
        return [images, targets, targets_eval, nnz, [it[2] for it in items]]

    def _next(self, batch_size):
        while self.cursor < len(self.lines):
            rec = self.lines[self.cursor]
            if rec[2] is None:
                self._decode(rec)
            self.cursor += 1
            if rec[2] is None:
                continue                                                          # undecodable image: skipped (:83-84)
            imgW = rec[2].shape[2]
            self.buffer.setdefault(imgW, []).append((rec[2], rec[3], rec[0]))
            if len(self.buffer[imgW]) == batch_size:
                return self._batch(self.buffer.pop(imgW), imgW)
        if not self.buffer:                                                       # data_gen.lua:125-129
            self.cursor = 0
            return None
        imgW = next(iter(self.buffer))                                            # final flush: one partial bucket per call
        return self._batch(self.buffer.pop(imgW), imgW)

    # -- optional read-ahead: the same sequence of batches, produced by a worker thread
    def _stop_worker(self):
        if self._worker is not None:
            self._stop = True
            while self._worker.is_alive():
                try:
                    self._q.get(timeout=0.05)
                except queue.Empty:
                    pass
            self._worker, self._q = None, None

    def _run(self, batch_size):
        while not self._stop:
            b = self._next(batch_size)
            self._q.put(b)
            if b is None:
                return

    def nextBatch(self, batch_size):
        if self.prefetch <= 0:
            return self._next(batch_size)
        if self._worker is None or self._worker_bs != batch_size:
            self._stop_worker()
            self._q, self._stop, self._worker_bs = queue.Queue(maxsize=self.prefetch), False, batch_size
            self._worker = threading.Thread(target=self._run, args=(batch_size,), daemon=True)
            self._worker.start()
        b = self._q.get()
        if b is None:
            self._worker.join()
            self._worker = None
        return b
