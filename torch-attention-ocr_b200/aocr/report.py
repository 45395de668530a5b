"""results.txt -> a browsable page (the job of src/visualizer/generate_html.py:6-75).

`Model.vis(output_dir)` makes every evaluation step append one line per image to `<output_dir>/results.txt`:
`image path \\t gold label \\t predicted label \\t predicted score \\t gold score` (src/model/model.lua:628-632).  This
turns that file into `<output_dir>/website/index.html` with the images copied to `website/images/`, one entry per line,
classed correct / incorrect, with a filter (all / correct / incorrect).  Same command line as the reference's script
(`--output_dir`, `--data_base_dir`); differences, all additive: the page is one self-contained file (no template pair,
no external script), a header reports word accuracy and the mean edit distance, a missing image is shown as its path
instead of aborting the run, and the word-frequency table is optional (`--freq`: a pickle or a `word<TAB>count` text
file — the reference ships a 7.2M-word lexicon count next to its script, this repo does not).

    python -m aocr.report --output_dir results --data_base_dir /data/90kDICT32px
"""
import argparse
import html
import os
import pickle
import shutil
import sys

_PAGE_HEAD = """<!doctype html>
<html><head><meta charset="utf-8"><title>Attention-OCR results</title>
<style>
body{font-family:sans-serif;margin:1.5em;background:#fafafa}
ol{list-style:none;padding:0;display:flex;flex-wrap:wrap;gap:.8em}
li{background:#fff;border:1px solid #ccc;border-top:4px solid #3a3;padding:.6em;min-width:14em;font-size:.9em}
li.f-incorrect{border-top-color:#c33}
li img{max-height:48px;image-rendering:pixelated}
nav button{margin-right:.5em;padding:.3em .9em}
.path{color:#777;font-size:.8em}
</style></head><body>
<h1>Attention-OCR results</h1>
"""

_PAGE_TAIL = """</ol>
<script>
function show(cls){for(const li of document.querySelectorAll('ol li')){li.style.display=li.classList.contains(cls)?'':'none';}}
</script>
</body></html>
"""


def edit_distance(a, b):
    prev = list(range(len(b) + 1))
    for i, ca in enumerate(a, 1):
        cur = [i]
        for j, cb in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (ca != cb)))
        prev = cur
    return prev[-1]


def read_results(path):
    """-> [(image path, gold, predicted, predicted score, gold score)]; lines that do not have five fields are skipped
    (generate_html.py:52 does the same)."""
    rows = []
    with open(path) as f:
        for line in f:
            items = line.rstrip("\n").split("\t")
            if len(items) == 5:
                rows.append(tuple(items))
    return rows


def load_freq(path):
    """word -> count.  A `.pkl` / `.pickle` file is unpickled (the reference's freq.pkl; only open files you trust —
    unpickling runs code), anything else is read as `word count` lines."""
    if not path:
        return None
    if path.endswith((".pkl", ".pickle")):
        with open(path, "rb") as f:
            return dict(pickle.load(f, encoding="latin-1"))
    freq = {}
    with open(path) as f:
        for line in f:
            parts = line.split()
            if len(parts) >= 2:
                freq[parts[0]] = int(parts[1])
    return freq


def image_file_name(rel_path):
    # one flat directory: path separators become underscores (generate_html.py:56 drops the leading "./" the same way)
    p = rel_path[2:] if rel_path.startswith("./") else rel_path.lstrip("/")
    return p.replace("/", "_")


def generate(output_dir, data_base_dir="/", freq_path=None, copy_images=True):
    result_path = os.path.join(output_dir, "results.txt")
    if not os.path.exists(result_path):
        raise FileNotFoundError("Result file %s not found" % result_path)
    rows = read_results(result_path)
    freq = load_freq(freq_path)
    site = os.path.join(output_dir, "website")
    img_dir = os.path.join(site, "images")
    os.makedirs(img_dir, exist_ok=True)
    n_ok = sum(1 for r in rows if r[1] == r[2])
    dist = sum(edit_distance(r[1], r[2]) for r in rows)
    out = [_PAGE_HEAD]
    out.append("<p>%d images, %d correct (word accuracy %.4f), mean edit distance %.3f</p>\n"
               % (len(rows), n_ok, n_ok / max(len(rows), 1), dist / max(len(rows), 1)))
    out.append('<nav><button onclick="show(\'f-all\')">All</button><button onclick="show(\'f-correct\')">Correct</button>'
               '<button onclick="show(\'f-incorrect\')">Incorrect</button></nav>\n<ol>\n')
    missing = 0
    for img_path, gold, pred, score_pred, score_gold in rows:
        src = os.path.join(data_base_dir, img_path)
        name = image_file_name(img_path)
        have = False
        if copy_images and os.path.isfile(src):
            shutil.copy(src, os.path.join(img_dir, name))
            have = True
        else:
            missing += 1
        out.append('<li class="%s f-all">\n' % ("f-correct" if gold == pred else "f-incorrect"))
        if have:
            out.append('<img src="%s" alt="%s"/><br/>\n' % (html.escape(os.path.join("images", name)), html.escape(img_path)))
        out.append('<span class="path">%s</span><br/>\n' % html.escape(img_path))
        out.append("gold: %s (%s)<br/>\n" % (html.escape(gold), html.escape(score_gold)))
        out.append("predicted: %s (%s)<br/>\n" % (html.escape(pred), html.escape(score_pred)))
        if freq is not None:
            out.append("gold frequency: %d<br/>\npredicted frequency: %d<br/>\n" % (freq.get(gold, 0), freq.get(pred, 0)))
        out.append("</li>\n")
    out.append(_PAGE_TAIL)
    html_path = os.path.join(site, "index.html")
    with open(html_path, "w") as f:
        f.write("".join(out))
    return {"html": html_path, "images": len(rows) - missing, "missing_images": missing, "rows": len(rows),
            "correct": n_ok, "edit_distance": dist}


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("--output_dir", default="results", help="directory containing results.txt")
    ap.add_argument("--data_base_dir", default="/", help="base directory of the image paths in results.txt")
    ap.add_argument("--freq", default=None, help="optional word-frequency table (pickle or 'word count' text)")
    ap.add_argument("--no_copy", action="store_true", help="do not copy the images into website/images")
    a = ap.parse_args(argv)
    r = generate(a.output_dir, a.data_base_dir, a.freq, copy_images=not a.no_copy)
    print("%(html)s: %(rows)d entries, %(correct)d correct, %(missing_images)d images not found" % r)
    return 0


if __name__ == "__main__":
    sys.exit(main())
