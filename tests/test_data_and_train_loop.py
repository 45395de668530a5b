"""CPU tests of the callers on either side of the hot path: the data layer
(src/data/data_gen.lua) and the train / test driver (src/train.lua).  No device work: the driver is exercised with a
stand-in model that records the calls it receives."""
import math
import os

import numpy as np
import pytest

from aocr.data import DataGen, load_image, scale_bilinear, rgb2y, str2numlist, numlist2str


def _write_images(tmp_path, specs):
    from PIL import Image
    rng = np.random.default_rng(0)
    lines = []
    for i, (w, h, label, mode) in enumerate(specs):
        a = rng.integers(0, 256, size=(h, w, 3) if mode == "RGB" else (h, w), dtype=np.uint8)
        name = f"img{i}.png"
        Image.fromarray(a, mode=mode).save(tmp_path / name)
        lines.append(f"{name} {label}")
    (tmp_path / "list.txt").write_text("\n".join(lines) + "\n")
    return lines


def test_scale_bilinear_follows_torch_image_semantics():
    x = np.arange(12, dtype=np.float32).reshape(1, 3, 4)
    up = scale_bilinear(x, 7, 3)                       # stretching: end points aligned, linear in between
    assert up.shape == (1, 3, 7) and up[0, 0, 0] == 0 and up[0, 0, -1] == 3
    np.testing.assert_allclose(up[0, 0], np.linspace(0, 3, 7), atol=1e-6)
    down = scale_bilinear(x, 2, 3)                     # shrinking: mean over each destination pixel's source interval
    np.testing.assert_allclose(down[0], [[0.5, 2.5], [4.5, 6.5], [8.5, 10.5]], atol=1e-6)
    frac = scale_bilinear(np.array([[[0., 1., 2.]]], np.float32), 2, 1)      # 3 -> 2: intervals [0,1.5) and [1.5,3)
    np.testing.assert_allclose(frac[0, 0], [(0 + 0.5 * 1) / 1.5, (0.5 * 1 + 2) / 1.5], atol=1e-6)
    assert scale_bilinear(x, 4, 3) is not x and np.array_equal(scale_bilinear(x, 4, 3), x)
    y = rgb2y(np.stack([np.full((2, 2), 1.0), np.zeros((2, 2)), np.zeros((2, 2))]).astype(np.float32))
    assert y.shape == (1, 2, 2) and abs(y[0, 0, 0] - 0.299) < 1e-6


def test_image_decoders_agree(tmp_path):
    """the same picture as PNG (through PIL), .npy, binary and ASCII PGM / PPM decodes to the same (C,H,W) array in
    [0,1] (image.load's contract, data_gen.lua:65); an undecodable file raises (the caller's pcall skips it)"""
    from PIL import Image
    rng = np.random.default_rng(3)
    gray = rng.integers(0, 256, size=(7, 11), dtype=np.uint8)
    rgb = rng.integers(0, 256, size=(7, 11, 3), dtype=np.uint8)
    Image.fromarray(gray, mode="L").save(tmp_path / "g.png")
    Image.fromarray(rgb, mode="RGB").save(tmp_path / "c.png")
    np.save(tmp_path / "g.npy", gray)
    np.save(tmp_path / "c.npy", rgb)
    np.save(tmp_path / "c_chw.npy", (rgb.transpose(2, 0, 1) / 255.0).astype(np.float32))
    (tmp_path / "g.pgm").write_bytes(b"P5\n# a comment\n11 7\n255\n" + gray.tobytes())
    (tmp_path / "c.ppm").write_bytes(b"P6 11 7 255\n" + rgb.tobytes())
    (tmp_path / "g_ascii.pgm").write_text("P2\n11 7\n255\n" + " ".join(str(v) for v in gray.ravel()) + "\n")
    (tmp_path / "g16.pgm").write_bytes(b"P5 11 7 65535\n" + (gray.astype(">u2") * 257).tobytes())
    g_ref, c_ref = load_image(str(tmp_path / "g.png")), load_image(str(tmp_path / "c.png"))
    assert g_ref.shape == (1, 7, 11) and c_ref.shape == (3, 7, 11) and g_ref.dtype == np.float32
    for name in ("g.npy", "g.pgm", "g_ascii.pgm", "g16.pgm"):
        np.testing.assert_allclose(load_image(str(tmp_path / name)), g_ref, atol=1e-6, err_msg=name)
    for name in ("c.npy", "c_chw.npy", "c.ppm"):
        np.testing.assert_allclose(load_image(str(tmp_path / name)), c_ref, atol=1e-6, err_msg=name)
    (tmp_path / "bad.pgm").write_bytes(b"P5 11 7 255\n" + gray.tobytes()[:10])
    (tmp_path / "bad.png").write_bytes(b"not an image")
    for name in ("bad.pgm", "bad.png", "absent.npy"):
        with pytest.raises(Exception):
            load_image(str(tmp_path / name))


def test_datagen_batches_follow_the_reference_format(tmp_path):
    specs = [(64, 32, "hello", "RGB"), (200, 50, "a1", "L"), (64, 32, "xyz9", "RGB"), (20, 40, "q", "L"), (64, 32, "ab", "RGB")]
    _write_images(tmp_path, specs)
    (tmp_path / "list.txt").write_text((tmp_path / "list.txt").read_text() + "missing.png zz\n")      # undecodable: skipped
    # the reference's behaviour: every image forced to width 100 (data_gen.lua:78) -> one bucket
    d = DataGen(str(tmp_path), "list.txt", 10.0, fixed_width=100, log=lambda m: None)
    assert d.size() == 6
    b = d.nextBatch(4)
    images, targets, targets_eval, nnz, paths = b
    assert images.shape == (4, 1, 32, 100) and images.dtype == np.float32 and 0 <= images.min() and images.max() <= 255
    assert paths == ["img0.png", "img1.png", "img2.png", "img3.png"]
    assert targets.shape == targets_eval.shape == (4, 6)              # longest label "hello": GO + 5 chars
    assert targets[0].tolist() == str2numlist("hello")[:-1] and targets_eval[0].tolist() == str2numlist("hello")[1:]
    assert targets[3].tolist() == [2, str2numlist("q")[1], 1, 1, 1, 1] and targets_eval[3].tolist()[:2] == [str2numlist("q")[1], 3]
    assert nnz == sum(len(l) + 1 for l in ("hello", "a1", "xyz9", "q"))
    assert numlist2str(targets_eval[2][:4]) == "xyz9"
    last = d.nextBatch(4)                                              # final flush: the partial bucket (data_gen.lua:130-153)
    assert last[0].shape[0] == 1 and last[4] == ["img4.png"]
    assert d.nextBatch(4) is None and d.cursor == 0                    # epoch end, cursor rewound (:125-129)
    # true bucketing (the width the line above :78 computes): ceil(clamp(W/H, 0.5, max_aspect) * 32)
    d2 = DataGen(str(tmp_path), "list.txt", 3.0, fixed_width=None, log=lambda m: None, prefetch=2)
    seen = {}
    while True:
        b = d2.nextBatch(2)
        if b is None:
            break
        seen.setdefault(b[0].shape[3], []).extend(b[4])
    assert seen == {64: ["img0.png", "img2.png", "img4.png"], 96: ["img1.png"], 16: ["img3.png"]}   # 200/50 -> clamp 3.0; 20/40 -> 0.5
    d.shuffle()
    assert sorted(r[0] for r in d.lines) == sorted(f"img{i}.png" for i in range(5)) + ["missing.png"]


class _FakeModel:
    def __init__(self, val_losses):
        self.global_step, self.optim_state, self.calls, self.saved = 0, {"learningRate": 0.1}, [], []
        self.val_losses = list(val_losses)

    def step(self, batch, forward_only, beam_size, trie):
        self.calls.append((forward_only, batch[0].shape[0], beam_size, self.optim_state["learningRate"]))
        if forward_only:
            return self.val_losses.pop(0) if self.val_losses else 1.0, [batch[3], float(batch[0].shape[0])]
        return 2.0 * batch[3], [batch[3], 0.0]

    def save(self, path):
        p = path + ".npz"
        open(p, "w").write("ckpt %d" % self.global_step)
        self.saved.append(p)
        return p

    def vis(self, d):
        self.vis_dir = d


class _Data:
    def __init__(self, n_batches, b=4):
        self.n, self.i, self.b, self.shuffles = n_batches, 0, b, 0

    def shuffle(self):
        self.shuffles += 1

    def nextBatch(self, bs):
        if self.i == self.n:
            self.i = 0
            return None
        self.i += 1
        return [np.zeros((self.b, 1, 32, 100), np.float32), None, None, 10, None]


def test_train_loop_checkpoints_validates_and_decays(tmp_path):
    from aocr.train import train, build_parser
    opt = build_parser().parse_args(["-phase", "train", "-learning_rate_min", "0.02", "-lr_decay", "0.5"])
    logs = []

    class L:
        def info(self, m):
            logs.append(m)
    model = _FakeModel(val_losses=[5.0, 6.0, 7.0, 3.0])           # val loss rises twice -> two decays, then falls
    tr, va = _Data(4), _Data(1)
    train(model, "train", 4, 2, tr, va, str(tmp_path), 2, math.inf, 1, False, str(tmp_path / "out"), None, opt, L())
    assert model.global_step == 8 and tr.shuffles == 2            # train.lua:94-96: shuffle at every epoch start
    # checkpoints at steps 2,4 / 6,8 plus one at each epoch end (train.lua:116-128,176-178); final-model follows the latest
    assert [os.path.basename(p) for p in model.saved] == ["model-2.npz", "model-4.npz", "model-4.npz", "model-6.npz",
                                                          "model-8.npz", "model-8.npz"]
    assert open(tmp_path / "final-model.npz").read() == "ckpt 8"
    assert not os.path.exists(tmp_path / ".final-model.tmp")
    # decay when the validation loss went up (train.lua:163-168,207-212): 5 -> 6 at the step-4 checkpoint (0.05), 6 -> 7 at
    # the end of epoch 1 (0.025); 7 -> 3 -> 1 -> 1 afterwards: no further decay, and never below learning_rate_min
    lrs = [c[3] for c in model.calls if not c[0]]
    assert lrs == [0.1] * 4 + [0.025] * 4
    assert any(m.startswith("Decay lr, current Lr: 0.050000") for m in logs)
    assert any(m.startswith("Decay lr, current Lr: 0.025000") for m in logs)
    assert logs[1] == "nan" and logs[2] == "%f" % math.exp(2.0)   # per-step perplexity of the totals BEFORE the step (Q8)
    assert any(m.startswith("Step 2 - training perplexity = %f" % math.exp(2.0)) for m in logs)
    assert any(m.startswith("Step 2 - Val Accuracy = 1.000000, loss = %f" % math.exp(5.0 / 10)) for m in logs)
    assert any(m.startswith("Epoch: 1, Step 4 - Val Accuracy") for m in logs)


def test_test_phase_is_one_forward_only_epoch(tmp_path):
    from aocr.train import train, build_parser
    opt = build_parser().parse_args(["-phase", "test", "-visualize"])
    logs = []

    class L:
        def info(self, m):
            logs.append(m)
    model = _FakeModel(val_losses=[])
    model.global_step = 77
    train(model, "test", 4, 50, _Data(3), None, str(tmp_path), 2, math.inf, 5, True, str(tmp_path / "o"), "TRIE", opt, L())
    assert model.global_step == 3 and all(c[0] and c[2] == 5 for c in model.calls) and len(model.calls) == 3
    assert model.vis_dir == str(tmp_path / "o") and not model.saved
    assert "Number of samples 8 - Accuracy = 1.000000" in logs and "Epoch: 1 Number of samples 12 - Accuracy = 1.000000" in logs


def test_cli_options_mirror_the_reference():
    from aocr.train import build_parser
    o = build_parser().parse_args([])
    assert (o.batch_size, o.max_decoder_l, o.max_encoder_l, o.learning_rate, o.lr_decay, o.seed) == (400, 50, 80, 0.1, 0.5, 910820)
    assert o.phase == "test" and o.beam_size == 1 and o.input_feed is False and o.target_vocab_size == 39
    ref = open("/root/reference/src/train.lua").read() if os.path.exists("/root/reference/src/train.lua") else None
    if ref:    # in the build container: every cmd:option of the reference is an option here
        import re
        for name in re.findall(r"cmd:option\('-([a-z_]+)'", ref):
            assert hasattr(o, name), name
