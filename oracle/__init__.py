"""CPU oracle for the torch-Attention-OCR recognition hot path.

TEST INFRASTRUCTURE ONLY.  This package is a float64 CPU restatement of the
reference's algorithm (da03/torch-Attention-OCR, Lua/Torch7).  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import it.  The product path (the `aocr` package over `libaocr.so`)
never imports, calls or links anything in here.

PARITY UNPINNED: the reference ships no golden vectors, known-answer tests or
fixtures and its Torch7/LuaJIT stack cannot run in this
image, so the oracle is pinned only by (i) autograd and finite-difference
checks of its own hand-written backward and (ii) the committed fixtures under
`tests/golden/`, which this oracle itself generated (`tests/golden/make_golden.py`).
"""
from .layout import Config, GROUPS, param_specs, group_sizes, init_params, init_bn_stats, unflatten, flatten  # noqa: F401
from .model import Oracle  # noqa: F401
from .synth import make_batch, str2numlist, numlist2str  # noqa: F401
from .trie import load_dictionary, flatten as flatten_trie  # noqa: F401
