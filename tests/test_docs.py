"""Documentation that can drift is checked against the sources: every environment switch the library, the Python mirror
or bench.py reads is listed in DESIGN.md's appendix, and the appendix lists nothing that no longer exists."""
import glob
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _read(*parts):
    return open(os.path.join(ROOT, *parts)).read()


def test_environment_switch_table_is_complete():
    used = set()
    pats = [r'getenv\("(AOCR_[A-Z0-9_]+)"\)', r'environ\.get\("(AOCR_[A-Z0-9_]+)"', r'environ\["(AOCR_[A-Z0-9_]+)"\]']
    files = glob.glob(os.path.join(ROOT, "torch-attention-ocr_b200", "csrc", "*")) + \
        glob.glob(os.path.join(ROOT, "torch-attention-ocr_b200", "aocr", "*.py")) + [os.path.join(ROOT, "bench.py")]
    for f in files:
        if os.path.isfile(f):
            txt = open(f, errors="replace").read()
            for p in pats:
                used.update(re.findall(p, txt))
    assert len(used) > 20
    appendix = _read("DESIGN.md").split("## Appendix — environment switches", 1)[1]
    listed = set(re.findall(r"`(AOCR_[A-Z0-9_]+)`", appendix))
    assert used - listed == set(), f"switches read by the code but missing from DESIGN.md: {sorted(used - listed)}"
    assert listed - used == set(), f"switches in DESIGN.md that nothing reads: {sorted(listed - used)}"


def test_design_names_every_section8_row():
    design = _read("DESIGN.md")
    for row in ("a1 CNN", "a2 encoder", "a3 decoder", "a6 train step", "a7 greedy decode", "a11 sgd_list", "b boundary",
                "c oracle", "d measurement", "e multi-GPU", "f1 beam", "f2 data path", "f3 checkpoints", "f4 train loop"):
        assert row in design, row
    assert "parity unpinned" in design and "PARITY UNPINNED" in _read("oracle", "__init__.py")
