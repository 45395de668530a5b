"""Batch tuple of the reference's data layer (src/data/data_gen.lua:97-120) from synthetic inputs.

`DataGen:nextBatch` returns `{images, targets, targets_eval, num_nonzeros, img_paths}`; real image decode
and disk I/O are out of scope, so this generator emits the same tuple from seeded random
images/labels, bucketed by width exactly like the reference (`self.buffer[imgW]`, data_gen.lua:92-96).
"""
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
