"""aocr — host-side mirror of torch-Attention-OCR's Lua `Model` interface over libaocr.so (sm_100a).

The reference is Lua/Torch7 (no LuaJIT in this image), so the host side above the C ABI is Python
(`ctypes`), mirroring `src/model/model.lua` (Model:create/load/step/save/vis/shutdown), the batch tuple
of `src/data/data_gen.lua` and `optim.sgd_list` of `src/optim/optim_sgd.lua`.  The Lua twin that a
reference maintainer would drop in lives in `../lua/`.

There is NO CPU fallback: importing this package without a built libaocr.so raises.
"""
from .capi import Lib, AocrError, AocrConfig, GROUPS, Trie, lib_path  # noqa: F401
from .model import Model  # noqa: F401
from .optim import sgd_list  # noqa: F401
from .data import SyntheticDataGen, str2numlist, numlist2str  # noqa: F401
from . import dist  # noqa: F401
