"""Where the GEMM time of one denoising step goes: every ea_gemm call of a configs[1] step is recorded,
grouped by shape/epilogue signature, and each group is replayed alone inside one CUDA graph
(CUDA events; members of a group use their own weights, so weights come from HBM, not L2).
Usage: python tools/gemm_breakdown.py [out.json]"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from editanything_b200 import _lib as L  # noqa: E402
from editanything_b200 import ops  # noqa: E402
from editanything_b200.denoise import DenoiseEngine, ddim_schedule  # noqa: E402
from editanything_b200.unet_spec import SD15, make_state_dict  # noqa: E402


def time_group(recs, min_launches=60):
    reps = max(1, (min_launches + len(recs) - 1) // len(recs))
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    def issue(r):
        if r[0] == "grouped":
            ops.gemm_grouped(r[1])
        else:
            ops.gemm(r[0], r[1], r[2], **r[3])
    with torch.cuda.stream(s):
        for r in recs:
            issue(r)
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            for r in recs:
                issue(r)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (2 * reps * len(recs))


def main():
    dev = torch.device("cuda:0")
    cfg = SD15
    usd = make_state_dict(cfg, "unet", 101, device=dev)
    csds = [make_state_dict(cfg, "controlnet", 102, device=dev), make_state_dict(cfg, "controlnet", 103, device=dev)]
    eng = DenoiseEngine(cfg, usd, csds, dev)
    del usd, csds
    x, ctx, hints = bench.make_inputs(cfg, 2, 64, 77, 11)
    ts, a, ap = ddim_schedule(50)
    eng.prepare(ctx, hints, [0.5, 1.0])
    eng.begin(x[:1], guidance=9.0, use_graph=False)
    probe = bench.GemmProbe(ops)
    eng.ops = eng.runner.ops = eng.unet.ops = probe
    for c in eng.cns:
        c.ops = probe
    eng.runner.ops = probe
    eng.step(int(ts[0]), float(a[0]), float(ap[0]))
    probe.records.clear()
    eng.step(int(ts[1]), float(a[1]), float(ap[1]))
    torch.cuda.synchronize()
    ops.set_lane(0, False)
    groups = {}
    for r in probe.records:
        G = 1
        if r[0] == "grouped":
            G = len(r[1])
            a_, w_, out, kw = r[1][0]
            fl = r[4]
        else:
            a_, w_, out, kw, fl = r
        conv = kw.get("conv")
        mode = kw.get("mode", L.EA_GEMM_LINEAR)
        M = conv[0] * conv[1] * conv[2] if conv else (kw.get("M") or a_.shape[0])
        key = (mode, M, w_.shape[0], w_.shape[1], kw.get("act", 0), kw.get("residual") is not None,
               bool(kw.get("accumulate")), kw.get("rowvec") is not None, kw.get("out2") is not None,
               kw.get("a_extra") is not None, kw.get("out_f32") is not None, G, kw.get("ln") is not None,
               kw.get("rowstats_out") is not None)
        groups.setdefault(key, []).append(r)
    rows = []
    lib = L.lib()
    n_sm = torch.cuda.get_device_properties(0).multi_processor_count
    for key, recs in groups.items():
        us = time_group(recs)
        mode, M, N, K = key[:4]
        kw0 = recs[0][1][0][3] if recs[0][0] == "grouped" else recs[0][3]
        conv = kw0.get("conv")
        if conv:
            cin = conv[3]
            ex = kw0.get("a_extra")
            kb = 9 * ((cin + 63) // 64) + (9 * ((ex.shape[-1] + 63) // 64) if ex is not None else 0)
            if mode == getattr(L, "EA_GEMM_CONV1X1", -99):
                kb = (cin + 63) // 64
        else:
            kb = (K + 63) // 64
        out5 = (C.c_int * 5)()
        lib.ea_gemm_plan((M + 127) // 128, N, kb, key[4], ops.GEMM_WS_BYTES - 65536, n_sm, out5)
        fl = recs[0][4]
        rows.append({"mode": mode, "M": M, "N": N, "K": K, "act": key[4], "res": key[5], "acc": key[6], "rowvec": key[7],
                     "out2": key[8], "extra": key[9], "f32": key[10], "groups": key[11], "ln": key[12], "stats": key[13],
                     "n": len(recs), "us": round(us, 2),
                     "tflops": round(fl / us / 1e6, 1), "total_ms": round(us * len(recs) / 1e3, 3), "kb": kb,
                     "plan": list(out5)})
    rows.sort(key=lambda r: -r["total_ms"])
    if len(sys.argv) > 1:       # first: a closed stdout pipe (| head) must not lose the data
        json.dump(rows, open(sys.argv[1], "w"), indent=0)
    tot = sum(r["total_ms"] for r in rows)
    print(f"{len(probe.records)} gemm launches, {len(rows)} signatures, sum of isolated times {tot:.3f} ms")
    print("mode      M     N      K act res acc  n      us  TFLOP/s  total_ms  cum%  kb plan[bn,stages,splits,occ,two] (xG = grouped launch of G networks; plan shown for one group)")
    cum = 0.0
    for r in rows:
        cum += r["total_ms"]
        print(f"{r['mode']:3d} {r['M']:6d} {r['N']:5d} {r['K']:6d} {r['act']:3d} {int(r['res']):3d} {int(r['acc']):3d} {r['n']:3d} "
              f"{r['us']:7.2f} {r['tflops']:8.1f} {r['total_ms']:9.3f} {100 * cum / tot:5.1f} {r['kb']:4d} {r['plan']}"
              f"{' x' if r['extra'] else ''}{' rv' if r['rowvec'] else ''}{' o2' if r['out2'] else ''}{' f32' if r['f32'] else ''}"
              f"{' x' + str(r['groups']) + 'G' if r['groups'] > 1 else ''}{' ln' if r['ln'] else ''}{' st' if r['stats'] else ''}")


if __name__ == "__main__":
    main()
