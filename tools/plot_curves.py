"""GFLOP/s-vs-N curves from the reference's `output_*.m` files, overlaid — what `cuda/plot.py:30-40` does with
matplotlib (not in this image), written as a dependency-free SVG so the OLD/NEW comparison of the tutorial
(`cuda/makefile:41-46`: `plot.py output_old.m output_new.m`) works here too.

    python tools/plot_curves.py [-o out.svg] [--log] output_a.m output_b.m ...

A file is: `version = 'name';`, optional device line, `MY_MMult = [`, rows `N gflops diff`, `];`
(`cuda/test_MMult.cpp:41,128`).  Rows are found by pattern, so harness error text in between is skipped."""
import math
import re
import sys

ROW = re.compile(r"^\s*(\d+)\s+([0-9.eE+-]+)\s+([0-9.eE+-]+|nan|inf)\s*$")
VERSION = re.compile(r"version\s*=\s*'([^']*)'")
COLORS = ["#1f77b4", "#d62728", "#2ca02c", "#ff7f0e", "#9467bd", "#8c564b", "#e377c2", "#7f7f7f", "#17becf", "#bcbd22"]


def read_curve(path):
    """-> (label, [N...], [GFLOP/s...], [diff...])"""
    label, xs, ys, ds = None, [], [], []
    with open(path) as f:
        for line in f:
            m = VERSION.search(line)
            if m and label is None:
                label = m.group(1)
            r = ROW.match(line)
            if r:
                xs.append(int(r.group(1)))
                ys.append(float(r.group(2)))
                ds.append(float(r.group(3)))
    return label or path, xs, ys, ds


def _ticks(lo, hi, log):
    if log:
        return [10.0 ** e for e in range(int(math.floor(math.log10(lo))), int(math.ceil(math.log10(hi))) + 1)]
    step = 10 ** math.floor(math.log10(hi - lo or 1))
    for mul in (1, 2, 5, 10):
        if (hi - lo) / (step * mul) <= 8:
            step *= mul
            break
    t, out = math.floor(lo / step) * step, []
    while t <= hi + 1e-9:
        out.append(t)
        t += step
    return out


def render(curves, log=False, width=960, height=560):
    curves = [c for c in curves if c[1]]
    if not curves:
        raise ValueError("no data rows found")
    L, R, T, B = 80, 260, 30, 50
    xs = [x for c in curves for x in c[1]]
    ys = [y for c in curves for y in c[2] if y > 0 or not log]
    x0, x1 = min(xs), max(xs)
    y0, y1 = (min(ys), max(ys)) if log else (0.0, max(ys))
    if x1 == x0:
        x1 = x0 + 1
    if y1 <= y0:
        y1 = y0 + 1
    fy = (lambda v: math.log10(max(v, y0))) if log else (lambda v: v)
    px = lambda x: L + (x - x0) / (x1 - x0) * (width - L - R)
    py = lambda y: height - B - (fy(y) - fy(y0)) / (fy(y1) - fy(y0)) * (height - T - B)
    o = [f'<svg xmlns="http://www.w3.org/2000/svg" width="{width}" height="{height}" font-family="sans-serif" font-size="12">',
         f'<rect width="{width}" height="{height}" fill="white"/>']
    for t in _ticks(y0, y1, log):
        if y0 <= t <= y1:
            o.append(f'<line x1="{L}" x2="{width - R}" y1="{py(t):.1f}" y2="{py(t):.1f}" stroke="#ddd"/>')
            o.append(f'<text x="{L - 6}" y="{py(t) + 4:.1f}" text-anchor="end">{t:g}</text>')
    for t in sorted(set(xs)):
        if (t - x0) % max(1, (x1 - x0) // 8) == 0 or t in (x0, x1):
            o.append(f'<text x="{px(t):.1f}" y="{height - B + 16}" text-anchor="middle">{t}</text>')
    o.append(f'<rect x="{L}" y="{T}" width="{width - L - R}" height="{height - T - B}" fill="none" stroke="black"/>')
    o.append(f'<text x="{(L + width - R) / 2}" y="{height - 10}" text-anchor="middle">shape (M = N = K)</text>')
    o.append(f'<text x="16" y="{(T + height - B) / 2}" text-anchor="middle" transform="rotate(-90 16 {(T + height - B) / 2})">GFLOP/s{" (log)" if log else ""}</text>')
    for i, (label, cx, cy, _) in enumerate(curves):
        col = COLORS[i % len(COLORS)]
        pts = " ".join(f"{px(x):.1f},{py(y):.1f}" for x, y in zip(cx, cy) if y > 0 or not log)
        o.append(f'<polyline fill="none" stroke="{col}" stroke-width="2" points="{pts}"/>')
        for x, y in zip(cx, cy):
            if y > 0 or not log:
                o.append(f'<circle cx="{px(x):.1f}" cy="{py(y):.1f}" r="2.5" fill="{col}"/>')
        ly = T + 14 + 18 * i
        o.append(f'<line x1="{width - R + 10}" x2="{width - R + 34}" y1="{ly - 4}" y2="{ly - 4}" stroke="{col}" stroke-width="2"/>')
        o.append(f'<text x="{width - R + 40}" y="{ly}">{label[:34]} ({max(cy):,.0f})</text>')
    o.append("</svg>")
    return "\n".join(o)


def main(argv):
    out, log, files = None, False, []
    it = iter(argv)
    for a in it:
        if a == "-o":
            out = next(it)
        elif a == "--log":
            log = True
        else:
            files.append(a)
    if not files:
        print(__doc__)
        return 2
    svg = render([read_curve(f) for f in files], log=log)
    if out:
        with open(out, "w") as f:
            f.write(svg)
        print("wrote", out)
    else:
        print(svg)
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
