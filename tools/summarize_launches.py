"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and
shares, plus (for ea_gemm_kernel) a per-grid-shape breakdown.  Usage:
    python tools/summarize_launches.py gpurun_out/launches.csv [--top N] [--traffic-json out.json]
With `--traffic-json` (launch list captured with dram__bytes_read.sum,dram__bytes_write.sum as well)
the per-kernel DRAM bytes per launch are written as JSON; bench.py reports them as `roofline.traffic`."""
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


def load_traffic(path):
    """{kernel name: [launches, dram bytes read, dram bytes written]} from the same csv."""
    txt = open(path, errors="replace").read()
    i = txt.find('"ID"')
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    acc = collections.defaultdict(lambda: [set(), 0.0, 0.0])
    for r in csv.DictReader(io.StringIO(txt[i:])):
        m = r.get("Metric Name")
        if m not in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            continue
        v = float(r["Metric Value"].replace(",", "")) * scale.get(r.get("Metric Unit", "byte"), 1)
        name = r["Kernel Name"].split("(")[0].replace("void ", "").strip()
        a = acc[name]
        a[0].add(r["ID"])
        a[1 if m.endswith("read.sum") else 2] += v
    return {k: [len(v[0]), v[1], v[2]] for k, v in acc.items()}


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


    if "--traffic-json" in sys.argv:
        import json
        out = sys.argv[sys.argv.index("--traffic-json") + 1]
        tr = load_traffic(path)
        doc = {"source": path, "note": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, "
               "one launch at a time (serialised, cold L2): per-launch averages",
               "kernels": {k: {"launches": n, "dram_read_bytes_per_launch": rd / n, "dram_write_bytes_per_launch": wr / n,
                               "dram_bytes_per_launch": (rd + wr) / n} for k, (n, rd, wr) in tr.items() if n}}
        json.dump(doc, open(out, "w"), indent=1)
        print()
        print(f"{'kernel':44s} {'n':>5s} {'DRAM MB/launch':>15s}")
        for k, v in sorted(doc["kernels"].items(), key=lambda kv: -kv[1]["dram_bytes_per_launch"] * kv[1]["launches"]):
            print(f"{k[:44]:44s} {v['launches']:5d} {v['dram_bytes_per_launch'] / 1e6:15.3f}")


if __name__ == "__main__":
    main()
