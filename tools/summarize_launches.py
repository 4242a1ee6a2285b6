"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and
shares, plus (for ea_gemm_kernel) a per-grid-shape breakdown.  Usage:
    python tools/summarize_launches.py gpurun_out/launches.csv [--top N]"""
import csv
import collections
import io
import sys


def load(path):
    txt = open(path, errors="replace").read()
    i = txt.find('"ID"')
    rows = list(csv.DictReader(io.StringIO(txt[i:])))
    out = []
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "nsecond": 1, "usecond": 1e3, "msecond": 1e6, "second": 1e9}.get(unit, 1)
        out.append((r["Kernel Name"], r.get("Grid Size", ""), r.get("Block Size", ""), ns))
    return out


def main():
    path = sys.argv[1]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 25
    rows = load(path)
    tot = sum(r[3] for r in rows)
    print(f"{len(rows)} launches, total {tot / 1e6:.3f} ms (serialised, cold-cache; compare SHARES)")
    by = collections.defaultdict(lambda: [0, 0.0])
    for k, g, b, ns in rows:
        name = k.split("(")[0]
        by[name][0] += 1
        by[name][1] += ns
    print(f"{'kernel':44s} {'n':>5s} {'ms':>9s} {'share':>7s} {'avg us':>8s}")
    for name, (n, ns) in sorted(by.items(), key=lambda kv: -kv[1][1]):
        print(f"{name[:44]:44s} {n:5d} {ns / 1e6:9.3f} {100 * ns / tot:6.1f}% {ns / n / 1e3:8.2f}")
    print()
    bg = collections.defaultdict(lambda: [0, 0.0])
    for k, g, b, ns in rows:
        if "gemm" in k or "attn" in k:
            bg[(k.split("(")[0], g)][0] += 1
            bg[(k.split("(")[0], g)][1] += ns
    print(f"{'kernel / grid':60s} {'n':>5s} {'ms':>9s} {'avg us':>8s}")
    for (name, g), (n, ns) in sorted(bg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{(name[:30] + ' ' + g)[:60]:60s} {n:5d} {ns / 1e6:9.3f} {ns / n / 1e3:8.2f}")


if __name__ == "__main__":
    main()
