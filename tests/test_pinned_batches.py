"""Page-locked batch tensors (pinned-memory prefetch): the ring allocator's recycling logic and the data
layer's use of it run on the CPU with a stand-in block type; the real aocr_host_alloc path is the GPU test below."""
import ctypes
import os

import numpy as np
import pytest

from aocr import capi
from aocr.data import DataGen


class _FakeBlock:
    """same interface as capi._PinnedBlock over ordinary memory"""
    made = 0

    def __init__(self, nbytes):
        _FakeBlock.made += 1
        self.nbytes = int(nbytes)
        self.buf = (ctypes.c_ubyte * self.nbytes)()
        self.ptr = ctypes.addressof(self.buf)

    array = capi._PinnedBlock.array


@pytest.fixture
def fake_blocks(monkeypatch):
    _FakeBlock.made = 0
    monkeypatch.setattr(capi, "_PinnedBlock", _FakeBlock)
    return _FakeBlock


def test_ring_recycles_blocks_and_grows(fake_blocks):
    ring = capi.PinnedRing(depth=3)
    addr = []
    for step in range(9):
        ring.next_batch()
        img = ring((4, 1, 32, 100), np.float32)
        tgt = ring((4, 7), np.int32)
        assert img.shape == (4, 1, 32, 100) and img.dtype == np.float32 and tgt.dtype == np.int32
        img[...] = step
        addr.append((img.ctypes.data, tgt.ctypes.data))
    assert ring.allocations == 6 == fake_blocks.made              # 3 generations x 2 tensors, then reuse only
    assert addr[0] == addr[3] == addr[6] and addr[1] == addr[4] and addr[0] != addr[1] != addr[2]
    # a larger batch grows the generation's block; an array handed out earlier stays valid and untouched
    ring.next_batch()
    old = ring((4, 1, 32, 100), np.float32)
    old[...] = 7.0
    ring.cur -= 1
    ring.next_batch()                                             # the same generation again
    big = ring((8, 1, 32, 400), np.float32)
    big[...] = 1.0
    assert ring.allocations == 7 and big.ctypes.data != old.ctypes.data and float(old.min()) == 7.0


def _write_pgms(tmp_path, n):
    rng = np.random.default_rng(1)
    lines = []
    for i in range(n):
        a = rng.integers(0, 256, size=(32, 64), dtype=np.uint8)
        (tmp_path / f"i{i}.pgm").write_bytes(b"P5 64 32 255\n" + a.tobytes())
        lines.append(f"i{i}.pgm w{i % 7}x")
    (tmp_path / "list.txt").write_text("\n".join(lines) + "\n")


def test_datagen_builds_batches_in_the_supplied_memory(tmp_path, fake_blocks):
    _write_pgms(tmp_path, 11)
    plain = DataGen(str(tmp_path), "list.txt", 10.0, log=lambda m: None)
    ring = capi.PinnedRing(depth=4)
    pinned = DataGen(str(tmp_path), "list.txt", 10.0, log=lambda m: None, alloc=ring, prefetch=2)
    n = 0
    while True:
        a, b = plain.nextBatch(4), pinned.nextBatch(4)
        if a is None:
            assert b is None
            break
        n += 1
        for x, y in zip(a[:3], b[:3]):
            assert x.dtype == y.dtype and np.array_equal(x, y)
        assert a[3] == b[3] and a[4] == b[4]
        gen = ring.generations[(n - 1) % 4]
        lo, hi = gen[0].ptr, gen[0].ptr + gen[0].nbytes
        assert lo <= b[0].ctypes.data < hi                        # the images live in the ring's block
    assert n == 3 and ring.allocations <= 3 * 3


def test_host_alloc_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present: covered by the GPU test")
    with pytest.raises(capi.AocrError) as e:
        capi.host_empty((4, 1, 32, 100))
    assert e.value.code == -2 and "aocr_host_alloc" in e.value.msg


@pytest.mark.gpu
def test_pinned_batches_step_like_pageable_ones():
    from aocr.data import synthetic_batch
    a = capi.host_empty((3, 5), np.int32)
    assert capi.is_pinned(a) and not capi.is_pinned(np.empty(4))
    ring = capi.PinnedRing(depth=2)
    batch = synthetic_batch(4, 100, 6, seed=11)
    ring.next_batch()
    pinned = {}
    for k in ("images", "targets", "targets_eval"):
        pinned[k] = ring(batch[k].shape, batch[k].dtype)
        pinned[k][...] = batch[k]
        assert capi.is_pinned(pinned[k])
    cfg = capi.AocrConfig(batch_size=4, max_encoder_l=30, max_decoder_l=10, encoder_num_hidden=512, encoder_num_layers=1,
                          decoder_num_layers=2, target_vocab_size=39, target_embedding_size=20, input_feed=1, dropout=0.0,
                          learning_rate=0.1, dp_rank=0, dp_world=1, global_batch=0, gemm_mode=0)
    losses = []
    for src in (batch, pinned):
        h = capi.Handle(cfg, 0)
        h.init_params(910820)
        losses.append(h.forward_backward(src["images"], src["targets"], src["targets_eval"]))
        h.close()
    assert losses[0] == losses[1] and np.isfinite(losses[0])
    del a, pinned, ring, src
    import gc
    gc.collect()
    assert len(capi._PINNED_OWNERS) == 0                          # every block went back through aocr_host_free
