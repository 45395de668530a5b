"""include/aocr.h from plain C: the header is C99 (and C++) clean, and a gcc-built consumer links against libaocr.so,
runs the host-side dictionary entry points and sees aocr_create fail loudly (tests/c/abi_consumer.c).  No GPU needed."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = os.path.join(ROOT, "include")
LIBDIR = os.path.join(ROOT, "torch-attention-ocr_b200", "lib")

pytestmark = pytest.mark.skipif(shutil.which("gcc") is None, reason="gcc not found")


def _run(cmd, **kw):
    return subprocess.run(cmd, capture_output=True, text=True, timeout=120, **kw)


def test_header_is_plain_c_and_cxx(tmp_path):
    src = tmp_path / "hdr.c"
    src.write_text('#include "aocr.h"\nint main(void) { return 0; }\n')
    r = _run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-I", INC, str(src)])
    assert r.returncode == 0, r.stderr
    if shutil.which("g++"):
        r = _run(["g++", "-std=c++11", "-Wall", "-Werror", "-fsyntax-only", "-x", "c++", "-I", INC, str(src)])
        assert r.returncode == 0, r.stderr


def test_c_consumer_links_and_runs(tmp_path):
    assert os.path.isfile(os.path.join(LIBDIR, "libaocr.so")), "build the library first (__graft_entry__.build())"
    exe = tmp_path / "abi_consumer"
    r = _run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", INC,
              os.path.join(ROOT, "tests", "c", "abi_consumer.c"), "-o", str(exe),
              "-L", LIBDIR, "-laocr", "-Wl,-rpath," + LIBDIR])
    assert r.returncode == 0, r.stderr
    r = _run([str(exe)])
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout + r.stderr
    assert "create refused" in r.stdout
