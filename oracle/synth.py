"""Synthetic batches in the reference's data-layer format (test infrastructure).

Restates the batch tuple `DataGen:nextBatch` emits (src/data/data_gen.lua:97-120):
  images       (b,1,32,W)  raw gray 0..255
  targets      (b,T) int32 = [GO, ch_1..ch_n]   padded with 1 (PAD)
  targets_eval (b,T) int32 = [ch_1..ch_n, EOS]  padded with 1
  num_nonzeros = sum(len+1),   T = max(len)+1
and the vocabulary map of src/utils/utils.lua:104-134 (1=PAD 2=GO 3=EOS, 4..13='0'..'9', 14..39='a'..'z').
"""
import numpy as np

ALPHABET = "0123456789abcdefghijklmnopqrstuvwxyz"


def str2numlist(s):
    """src/utils/utils.lua:104-118"""
    out = [2]
    for ch in s:
        l = ord(ch)
        out.append(l - 97 + 13 + 1 if l > 96 else l - 48 + 3 + 1)
    out.append(3)
    return out


def numlist2str(ids):
    """src/utils/utils.lua:120-134 (ids 1,2 map to '-' and '.', as in the reference)."""
    return "".join(chr(v - 1 - 13 + 97) if v > 13 else chr(v - 1 - 3 + 48) for v in ids)


def _stroke_images(rng, B, W):
    """Text-like synthetic images: per-image background level, a handful of random dark/bright strokes and
    blobs, mild pixel noise.  Unlike i.i.d. noise these give feature maps that differ between columns and
    images, so batch-norm gradients are not dominated by catastrophic cancellation in fp32."""
    img = np.empty((B, 1, 32, W), np.float64)
    for b in range(B):
        bg = rng.uniform(120, 250)
        im = np.full((32, W), bg)
        for _ in range(int(rng.integers(W // 12 + 2, W // 5 + 4))):
            h0, w0 = int(rng.integers(0, 30)), int(rng.integers(0, max(W - 2, 1)))
            dh, dw = int(rng.integers(2, 22)), int(rng.integers(1, 7))
            if rng.uniform() < 0.3:
                dh, dw = dw, int(rng.integers(3, 14))
            im[h0:h0 + dh, w0:w0 + dw] = rng.uniform(0, 110) if rng.uniform() < 0.8 else rng.uniform(200, 255)
        im += rng.normal(0, 6.0, size=im.shape)
        img[b, 0] = np.clip(im, 0, 255)
    return np.round(img * 4) / 4      # bilinear-resized inputs are non-integers; keep exact in fp32


def make_batch(B, W, max_label_len, seed=910820, min_label_len=1, force_T=None, kind="strokes"):
    """Seeded synthetic batch.  kind="noise": i.i.d. uniform integers 0..255 (cost-equivalent,
    numerically degenerate); kind="strokes": text-like structure (default for parity tests)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    if kind == "noise":
        images = rng.integers(0, 256, size=(B, 1, 32, W)).astype(np.float32)
    else:
        images = _stroke_images(rng, B, W).astype(np.float32)
    lens = rng.integers(min_label_len, max_label_len + 1, size=B)
    if force_T is not None:
        lens[0] = force_T - 1
        lens = np.minimum(lens, force_T - 1)
    labels = ["".join(ALPHABET[i] for i in rng.integers(0, 36, size=int(n))) for n in lens]
    lists = [str2numlist(s) for s in labels]
    T = max(len(l) for l in lists) - 1
    targets = np.ones((B, T), np.int32)
    targets_eval = np.ones((B, T), np.int32)
    nnz = 0
    for i, l in enumerate(lists):
        nnz += len(l) - 1
        targets[i, :len(l) - 1] = l[:-1]
        targets_eval[i, :len(l) - 1] = l[1:]
    return {"images": images, "targets": targets, "targets_eval": targets_eval,
            "num_nonzeros": int(nnz), "labels": labels}
