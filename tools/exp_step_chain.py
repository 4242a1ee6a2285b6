"""Experiment (build with EA_NVCC_EXTRA=-DEA_GEMM_TIMING): the 362 GEMM launches of one real step,
re-issued back to back in one CUDA graph (exactly bench.py's roofline measurement), with %globaltimer
stamps per launch: in-context interval (last exit -> last exit) next to the CTA-0 phases."""
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

lib = L.load()
lib.ea_gemm_chain_read.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
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
eng.step(int(ts[0]), float(a[0]), float(ap[0]))
probe.records.clear()
eng.step(int(ts[1]), float(a[1]), float(ap[1]))
torch.cuda.synchronize()
ops.set_lane(0, False)
recs = probe.records
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for a_, w_, out, kw, _ in recs:
        ops.gemm(a_, w_, out, **kw)
torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize()
assert lib.ea_gemm_chain_reset() == 0
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for a_, w_, out, kw, _ in recs:
        ops.gemm(a_, w_, out, **kw)
g.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
g.replay()
e1.record()
torch.cuda.synchronize()
n = len(recs)
buf = (C.c_ulonglong * (8 * n))()
assert lib.ea_gemm_chain_read(buf, n) == 0
v = [list(buf[i * 8:(i + 1) * 8]) for i in range(n)]
print(f"{n} launches, graph replay {e0.elapsed_time(e1):.3f} ms; first CTA0 entry -> last exit {(v[-1][7] - v[0][0]) / 1e6:.3f} ms")
groups = {}
rows = []
for i, (a_, w_, out, kw, fl) in enumerate(recs):
    conv = kw.get("conv")
    mode = kw.get("mode", 0)
    M = conv[0] * conv[1] * conv[2] if conv else (kw.get("M") or a_.shape[0])
    key = (mode, M, w_.shape[0], w_.shape[1], kw.get("act", 0), kw.get("residual") is not None, bool(kw.get("accumulate")))
    iv = (v[i][7] - v[i - 1][7]) / 1e3 if i else (v[i][7] - v[i][2]) / 1e3
    ph = [(v[i][b] - v[i][a]) / 1e3 for a, b in [(2, 3), (3, 4), (4, 5)]]
    rel = (v[i][2] - v[i - 1][7]) / 1e3 if i else 0.0
    tail = (v[i][7] - v[i][5]) / 1e3
    groups.setdefault(key, []).append((iv, rel, ph[0], ph[1], ph[2], tail))
tot = 0
out_rows = []
for key, L_ in groups.items():
    k = len(L_)
    avg = [sum(x[j] for x in L_) / k for j in range(6)]
    out_rows.append((avg[0] * k / 1e3, key, k, avg))
    tot += avg[0] * k / 1e3
out_rows.sort(key=lambda r: -r[0])
print(f"sum of intervals {tot:.3f} ms")
print("mode      M     N      K act res acc   n  interval  = release + first-op + main + epilogue + tail(CTA0 done -> last exit)   total_ms")
for t, key, k, avg in out_rows:
    print(f"{key[0]:3d} {key[1]:6d} {key[2]:5d} {key[3]:6d} {key[4]:3d} {int(key[5]):3d} {int(key[6]):3d} {k:3d} {avg[0]:8.2f}    {avg[1]:6.2f} {avg[2]:6.2f} {avg[3]:6.2f} {avg[4]:6.2f} {avg[5]:6.2f}   {t:7.3f}")
if len(sys.argv) > 1:
    json.dump([{"key": list(r[1]), "n": r[2], "avg": r[3], "total_ms": r[0]} for r in out_rows], open(sys.argv[1], "w"))
