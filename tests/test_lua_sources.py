"""The Lua side of the boundary cannot be executed here (no LuaJIT / Torch7 in the image): lex every .lua file of the
shim and check block / bracket structure, that `lua/model.lua` defines every `model:` method the reference's train loop
calls, and that every `C.aocr_*` the shim calls is declared in its own cdef (and therefore in include/aocr.h, which
tests/test_boundary.py ties to the cdef)."""
import glob
import os
import re

import pytest

from lua_lint import LuaSyntaxError, check_structure, defined_methods, lint, tokenize

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LUA_DIR = os.path.join(ROOT, "torch-attention-ocr_b200", "lua")
LUA_FILES = sorted(glob.glob(os.path.join(LUA_DIR, "*.lua")))


def test_linter_accepts_and_rejects():
    ok = """
    local t = {1, 2, [3] = "x]]", s = [==[ long ]] string ]==]}   -- comment with end
    --[[ block comment
         function if end ]]
    local function f(a, ...)
      if a then return 1 elseif not a then return 2 else return 3 end
      for i = 1, 10 do while i < 3 do i = i + 1 end end
      repeat a = a - 1 until a == 0
      do local x = 0x10ULL end
      return function() end
    end
    """
    assert lint(ok)
    for bad in ("function f() if x then end", "local t = {1, 2", "x = 'abc", "for i = 1, 2 do end end",
                "if x return end", "repeat x = 1 end", "f(a]]"):
        with pytest.raises(LuaSyntaxError):
            lint(bad)


@pytest.mark.parametrize("path", LUA_FILES, ids=[os.path.basename(p) for p in LUA_FILES])
def test_lua_file_structure(path):
    assert len(LUA_FILES) >= 3
    toks = lint(open(path).read())
    assert len(toks) > 50


def test_model_lua_defines_what_train_lua_calls():
    # the methods src/train.lua invokes on the model object: create / load (:258-262), step (:102), save (:125,:177),
    # vis (:79), shutdown (:214), plus global_step / optim_state fields it reads
    toks = lint(open(os.path.join(LUA_DIR, "model.lua")).read())
    methods = defined_methods(toks, "model") | defined_methods(toks, "Model")
    for m in ("create", "load", "step", "save", "vis", "shutdown"):
        assert m in methods, f"lua/model.lua defines no method {m!r}: {sorted(methods)}"


def test_shim_calls_only_declared_entry_points():
    cdef = open(os.path.join(LUA_DIR, "aocr_ffi.lua")).read().split("ffi.cdef[[", 1)[1].split("]]", 1)[0]
    declared = set(re.findall(r"\b(aocr_[a-z0-9_]+)\s*\(", cdef))
    assert len(declared) > 20
    for path in LUA_FILES:
        toks = tokenize(open(path).read())
        called = {toks[i + 2][1] for i in range(len(toks) - 2)
                  if toks[i][1] in ("C", "lib") and toks[i + 1][1] == "." and toks[i + 2][1].startswith("aocr_")}
        if os.path.basename(path) == "model.lua":
            assert len(called) >= 10                      # the check sees the calls it is meant to check
        missing = called - declared
        assert not missing, f"{os.path.basename(path)} calls undeclared entry points {sorted(missing)}"


def test_lua_checkpoint_layout_matches_python_layout():
    """lua/aocr_ckpt.lua flattens named tensors in M.ORDER; aocr/layout.py param_specs is the order the library and the
    Python checkpoint reader use: the two lists must be the same, group by group"""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "torch-attention-ocr_b200"))
    from aocr.layout import GROUPS, param_specs
    src = open(os.path.join(LUA_DIR, "aocr_ckpt.lua")).read()
    lint(src)
    names = lambda body: re.findall(r"'([^']+)'", body)
    enc = names(re.search(r"local ENC = \{([^}]*)\}", src).group(1))
    order_body = src.split("M.ORDER = {", 1)[1]
    lua = {"enc_fw": enc, "enc_bw": enc}
    for g in ("cnn", "decoder", "proj"):
        lua[g] = names(re.search(g + r" = \{([^}]*)\}", order_body).group(1))
    assert re.search(r"enc_fw = ENC, enc_bw = ENC", order_body)
    spec = param_specs({})
    for g in GROUPS:
        assert lua[g] == [n for n, _ in spec[g]], g
    assert names(re.search(r"M.GROUPS = \{([^}]*)\}", src).group(1)) == list(GROUPS)


def test_model_lua_loads_every_checkpoint_kind():
    src = open(os.path.join(LUA_DIR, "model.lua")).read()
    assert "K.normalise(torch.load(model_path))" in src
    ck = open(os.path.join(LUA_DIR, "aocr_ckpt.lua")).read()
    for kind in ("'reference'", "'named'", "'flat'"):
        assert kind in ck
