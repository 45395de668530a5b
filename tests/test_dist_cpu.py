"""world_size-2 gloo tests (CPU) of the data-parallel host logic: the exchange-hook protocol of aocr/dist.py
(kind 0 statistics, kind 1 gradient buckets, kind 2 join), batch sharding, and the algebra the engine relies on
(sum of per-rank gradients at 1/global_batch == single-device gradient; clip after the all-reduce), checked with
the float64 oracle standing in for the engine."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "torch-attention-ocr_b200")]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import ctypes as C
    from aocr.dist import GradSync, shard
    from oracle import Config, GROUPS, Oracle, init_params, init_bn_stats, make_batch

    torch.set_num_threads(2)
    cfg = Config(batch_size=4, max_encoder_l=12, max_decoder_l=7, encoder_num_hidden=8)
    full = make_batch(4, 40, 5, seed=21)
    batch = [full["images"], full["targets"], full["targets_eval"], full["num_nonzeros"], None]
    mine = shard(batch, rank, world)
    assert mine[0].shape[0] == 2
    orc = Oracle(cfg, init_params(cfg, 3), init_bn_stats(cfg))
    # per-rank gradients at the GLOBAL batch scale (BN statistics are per-rank here: the oracle has no SyncBN,
    # so use eval-mode-free quantities only: decoder/proj/encoder grads depend on BN through the CNN output, which
    # differs per rank -- therefore compare against a reference computed with the same per-rank CNN statistics)
    loss, grads, _ = orc.forward_backward(mine[0], mine[1], mine[2], global_batch=4)
    # a mock engine: flat buffer in the physical order [proj|decoder|enc_fw|enc_bw|cnn], drives the hook protocol
    order = ["proj", "decoder", "enc_fw", "enc_bw", "cnn"]
    flat = np.concatenate([grads[g] for g in order]).astype(np.float32)
    offs = np.cumsum([0] + [grads[g].size for g in order])
    bufs = {}

    def wrap(ptr, n):
        return torch.from_numpy(bufs[int(ptr)])[:n]

    gs = GradSync(wrap)
    hook = gs.cb

    def call(arr, kind):
        if arr is None:
            hook(None, None, 0, kind)
            return
        key = arr.ctypes.data
        bufs[key] = arr
        hook(None, C.c_void_p(key), arr.size, kind)

    stats = np.full(8, float(rank + 1), np.float32)
    call(stats, 0)                                        # BN statistics: summed immediately
    assert np.allclose(stats, 3.0)
    for a, b in ((0, 2), (2, 4), (4, 5)):                 # the three gradient buckets, then the join
        call(flat[offs[a]:offs[b]], 1)
    call(None, 2)
    assert [k for k, _ in gs.log.calls] == [0, 1, 1, 1, 2]
    # every rank now holds the same summed gradient
    chk = torch.from_numpy(flat.copy())
    dist.all_reduce(chk, op=dist.ReduceOp.MAX)
    assert np.array_equal(chk.numpy(), flat)
    total = torch.tensor([loss])
    dist.all_reduce(total)
    if rank == 0:
        out.put((flat.copy(), float(total[0]), offs.tolist()))
    dist.destroy_process_group()


def test_exchange_protocol_and_gradient_sum_two_ranks():
    world, port = 2, 29500 + (os.getpid() % 1000)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    flat, loss_sum, offs = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process reference: the two shards evaluated separately at global_batch=4 and summed
    sys.path[:0] = [ROOT]
    from oracle import Config, Oracle, init_params, init_bn_stats, make_batch
    cfg = Config(batch_size=4, max_encoder_l=12, max_decoder_l=7, encoder_num_hidden=8)
    full = make_batch(4, 40, 5, seed=21)
    ref, ref_loss = None, 0.0
    for r in range(2):
        orc = Oracle(cfg, init_params(cfg, 3), init_bn_stats(cfg))
        sl = slice(2 * r, 2 * r + 2)
        l, g, _ = orc.forward_backward(full["images"][sl], full["targets"][sl], full["targets_eval"][sl], global_batch=4)
        v = np.concatenate([g[k] for k in ["proj", "decoder", "enc_fw", "enc_bw", "cnn"]])
        ref = v if ref is None else ref + v
        ref_loss += l
    assert abs(loss_sum - ref_loss) < 1e-9 * abs(ref_loss)
    np.testing.assert_allclose(flat, ref.astype(np.float32), rtol=2e-5, atol=1e-7)


def test_shard_keeps_reference_batch_format():
    sys.path[:0] = [os.path.join(ROOT, "torch-attention-ocr_b200")]
    from aocr.dist import shard
    from aocr.data import make_batch_from_labels
    b = make_batch_from_labels(np.zeros((4, 1, 32, 100), np.float32), ["ab", "c", "defg", "h"])
    parts = [shard(b, r, 2) for r in range(2)]
    assert parts[0][0].shape == (2, 1, 32, 100) and parts[1][1].shape == b[1][2:].shape
    assert parts[0][3] + parts[1][3] == b[3]                      # num_nonzeros adds up
    with pytest.raises(AssertionError):
        shard(b, 0, 3)
