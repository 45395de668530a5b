"""results.txt -> website/index.html (aocr/report.py; the job of src/visualizer/generate_html.py)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "torch-attention-ocr_b200"))

from aocr import report  # noqa: E402


def _write_results(d, rows):
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, "results.txt"), "w") as f:
        for r in rows:
            f.write("\t".join(r) + "\n")
        f.write("a malformed line without tabs\n")          # skipped, as generate_html.py:52 skips it


def test_edit_distance():
    assert report.edit_distance("kitten", "sitting") == 3
    assert report.edit_distance("", "abc") == 3
    assert report.edit_distance("abc", "abc") == 0


def test_page_from_results(tmp_path):
    base = tmp_path / "data"
    (base / "1" / "2").mkdir(parents=True)
    (base / "1" / "2" / "w_cat.jpg").write_bytes(b"\xff\xd8fake")
    out = tmp_path / "results"
    _write_results(str(out), [
        ("./1/2/w_cat.jpg", "cat", "cat", "-0.010000", "-0.010000"),
        ("./1/2/w_gone.jpg", "dog", "d<g", "-1.500000", "-3.250000"),
    ])
    freq = tmp_path / "freq.txt"
    freq.write_text("cat 120\ndog 7\n")
    r = report.generate(str(out), str(base), str(freq))
    assert r["rows"] == 2 and r["correct"] == 1 and r["missing_images"] == 1 and r["edit_distance"] == 1
    page = open(r["html"]).read()
    assert os.path.isfile(out / "website" / "images" / "1_2_w_cat.jpg")
    assert page.count('class="f-correct f-all"') == 1 and page.count('class="f-incorrect f-all"') == 1
    assert "d&lt;g" in page and "d<g" not in page            # labels are escaped
    assert "gold frequency: 120" in page and "predicted frequency: 0" in page
    assert "word accuracy 0.5000" in page


def test_cli_and_missing_results(tmp_path):
    import pytest
    with pytest.raises(FileNotFoundError):
        report.generate(str(tmp_path / "nothing"))
    out = tmp_path / "r"
    _write_results(str(out), [("a.png", "x", "x", "0.000000", "0.000000")])
    assert report.main(["--output_dir", str(out), "--data_base_dir", str(tmp_path), "--no_copy"]) == 0
    assert os.path.isfile(out / "website" / "index.html")
