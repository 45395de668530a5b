"""CPU tests of the drop-in boundary: libaocr.so loads and exports every symbol include/aocr.h declares, the
ctypes prototypes cover the header, and the host-side mirrors (data layer, sgd_list) follow the reference.
No compute call is made here (there is no GPU in the build container)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "aocr.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(aocr_[a-z_0-9]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from aocr.capi import lib_path, exported_symbols
    path = lib_path()
    assert os.path.exists(path), "libaocr.so not built: run python -c 'import __graft_entry__ as g; g.build()'"
    dll = ctypes.CDLL(path)
    declared = header_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(dll, name), f"{name} declared in include/aocr.h but not exported"
    assert sorted(exported_symbols()) == declared, "ctypes prototypes out of sync with include/aocr.h"


def test_lua_ffi_cdef_declares_every_entry_point():
    """the LuaJIT twin cannot run here; at least its cdef must name every entry point of the header, with the same
    parameter count"""
    def protos(txt):
        txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
        out = {}
        for m in re.finditer(r"\b(aocr_[a-z_0-9]+)\s*\(([^;]*?)\)\s*;", txt, flags=re.S):
            if "(*" in m.group(0):
                continue                     # the callback typedef
            out[m.group(1)] = len([a for a in m.group(2).split(",") if a.strip()])
        return out
    hdr = protos(open(os.path.join(ROOT, "include", "aocr.h")).read())
    lua_txt = open(os.path.join(ROOT, "torch-attention-ocr_b200", "lua", "aocr_ffi.lua")).read()
    lua = protos(lua_txt.split("ffi.cdef[[", 1)[1].split("]]", 1)[0])
    assert set(hdr) == set(header_symbols())
    assert lua == hdr


def test_config_struct_matches_header():
    from aocr.capi import AocrConfig
    txt = open(os.path.join(ROOT, "include", "aocr.h")).read()
    body = re.search(r"typedef struct aocr_config \{(.*?)\} aocr_config;", txt, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        names = decl.split(None, 1)[1]
        fields += [n.strip() for n in names.split(",")]
    assert [f[0] for f in AocrConfig._fields_] == fields
    assert ctypes.sizeof(AocrConfig) == 4 * len(fields)


def test_create_fails_loudly_without_gpu_or_on_bad_config():
    import torch
    from aocr.capi import AocrConfig, Handle, AocrError
    cfg = AocrConfig(batch_size=4, max_encoder_l=30, max_decoder_l=10, encoder_num_hidden=512, encoder_num_layers=1,
                     decoder_num_layers=3, target_vocab_size=39, target_embedding_size=20, input_feed=1)
    with pytest.raises(AocrError) as e:
        Handle(cfg, 0)
    assert e.value.code in (-1, -2)
    if not torch.cuda.is_available():
        cfg.decoder_num_layers = 2
        with pytest.raises(AocrError) as e:      # no CPU fallback: creation must fail without a device
            Handle(cfg, 0)
        assert e.value.code == -2


def test_data_layer_format_matches_reference():
    from aocr.data import SyntheticDataGen, str2numlist, numlist2str, make_batch_from_labels
    assert str2numlist("a0z9") == [2, 14, 4, 39, 13, 3]            # utils.lua:104-118
    assert numlist2str([14, 4, 39, 13]) == "a0z9"
    imgs = np.zeros((2, 1, 32, 100), np.float32)
    images, tg, te, nnz, paths = make_batch_from_labels(imgs, ["ab", "c"])
    assert tg.tolist() == [[2, 14, 15], [2, 16, 1]] and te.tolist() == [[14, 15, 3], [16, 3, 1]]
    assert nnz == 5 and tg.dtype == np.int32 and images.shape == (2, 1, 32, 100)
    gen = SyntheticDataGen(10, widths=(100, 132), max_label_len=5, seed=1)
    seen, n = set(), 0
    while True:
        b = gen.nextBatch(4)
        if b is None:
            break
        assert len({b[0].shape[3]}) == 1 and b[0].shape[0] <= 4      # one width per batch, partial flushes allowed
        seen.add(b[0].shape[3])
        n += b[0].shape[0]
    assert n == 10 and seen <= {100, 132}


def test_sgd_list_mirror_clip_and_update():
    from aocr.optim import sgd_list

    class Vec:
        def __init__(self, v):
            self.v = np.asarray(v, np.float64)

        def norm(self):
            return float(np.linalg.norm(self.v))

        def mul(self, s):
            self.v *= s

        def add(self, a, other):
            self.v += a * other.v

    x = [Vec([1.0, 2.0]), Vec([0.5])]
    g = [Vec([30.0, 40.0]), Vec([0.1])]
    state = {"learningRate": 0.1}
    _, fx, stats = sgd_list(lambda _: (7.0, g, [3, 0]), x, state)
    np.testing.assert_allclose(x[0].v, [1 - 0.1 * 3.0, 2 - 0.1 * 4.0])     # clipped to norm 5
    np.testing.assert_allclose(x[1].v, [0.5 - 0.01])
    assert fx == [7.0] and state[1]["evalCounter"] == 1
