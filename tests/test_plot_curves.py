"""tools/plot_curves.py: the output_*.m reader (format of cuda/test_MMult.cpp:41,128 as cuda/plot.py:5-28 reads it)
and the dependency-free SVG overlay.  CPU only."""
import os
import sys
import xml.etree.ElementTree as ET

import _libs

sys.path.insert(0, os.path.join(_libs.ROOT, "tools"))
import plot_curves  # noqa: E402

SAMPLE = """version = 'MMult_demo';
GPU Device 0: "NVIDIA B200" with compute capability 10.0

MY_MMult = [

 error: i 0  j 0 diff 6.143436  got -19.135733  expect -25.279169 diff too big !
256 1973.23 0.000000e+00 
512 9310.85 3.814697e-05 
1024 40142.88 8.773804e-05 
];
"""


def test_reader_skips_noise_and_keeps_rows(tmp_path):
    p = tmp_path / "output_demo.m"
    p.write_text(SAMPLE)
    label, xs, ys, ds = plot_curves.read_curve(str(p))
    assert label == "MMult_demo" and xs == [256, 512, 1024]
    assert ys == [1973.23, 9310.85, 40142.88] and ds[1] == 3.814697e-05


def test_committed_curves_parse():
    d = os.path.join(_libs.ROOT, "profiles")
    files = sorted(f for f in os.listdir(d) if f.startswith("output_") and f.endswith(".m"))
    assert files
    for f in files:
        label, xs, ys, _ = plot_curves.read_curve(os.path.join(d, f))
        assert label and len(xs) == len(ys)
        if "MMult_cuda_12" not in f and "MMult_cuda_11" not in f:        # those fail the harness check on B200
            assert xs and xs[-1] == 4096 and all(y > 0 for y in ys), f


def test_svg_is_well_formed_and_has_one_polyline_per_curve(tmp_path):
    a = tmp_path / "a.m"
    b = tmp_path / "b.m"
    a.write_text(SAMPLE)
    b.write_text(SAMPLE.replace("MMult_demo", "other").replace("40142.88", "63075.21"))
    for log in (False, True):
        svg = plot_curves.render([plot_curves.read_curve(str(a)), plot_curves.read_curve(str(b))], log=log)
        root = ET.fromstring(svg)
        ns = "{http://www.w3.org/2000/svg}"
        lines = root.findall(f"{ns}polyline")
        assert len(lines) == 2 and all(len(pl.get("points").split()) == 3 for pl in lines)
        text = "".join(t.text or "" for t in root.iter(f"{ns}text"))
        assert "MMult_demo" in text and "other" in text and "GFLOP/s" in text
    out = tmp_path / "o.svg"
    assert plot_curves.main(["-o", str(out), "--log", str(a), str(b)]) == 0 and out.exists()
