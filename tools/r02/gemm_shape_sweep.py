"""Per-shape sweep of the step's small-K linears (3 networks grouped like the lockstep encoder): tile width, kernel
flavour (one-tile kernel / CTA pairs / 8-warp persistent) -> us and TFLOP/s.  Dev tool (GPU box)."""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))

import torch

from editanything_b200 import _lib as L, ops

dev = torch.device("cuda:0")
dt = ops.half_dtype()
SHAPES = [  # (M, N, K, act, residual)
    (8192, 960, 320, "none", False), (8192, 2560, 320, "geglu", False), (8192, 320, 320, "none", True),
    (8192, 320, 1280, "none", True), (2048, 1920, 640, "none", False), (2048, 5120, 640, "geglu", False),
    (2048, 640, 640, "none", True), (512, 10240, 1280, "geglu", False),
]
G = 3
rows = []
for (M, N, K, act, res) in SHAPES:
    A = [torch.randn(M, K, device=dev).to(dt) for _ in range(G)]
    W = [(torch.randn(N, K, device=dev) * 0.05).to(dt) for _ in range(G)]
    No = N // 2 if act == "geglu" else N
    O = [torch.empty(M, No, device=dev, dtype=dt) for _ in range(G)]
    R = [torch.randn(M, No, device=dev).to(dt) for _ in range(G)] if res else [None] * G
    bias = [torch.randn(N, device=dev) for _ in range(G)]
    bns = (128,) if act == "geglu" else (64, 96, 128, 160, 192, 256)
    for flavour, kw in (("tile", dict(force_persistent=-1)), ("pair", dict(force_persistent=-1, force_2cta=1)),
                        ("persist8", dict(force_persistent=2)), ("persist4", dict(force_persistent=1))):
        for bn in bns:
            if bn > N:
                continue
            def run():
                ops.gemm_grouped([(A[g], W[g], O[g], dict(bias=bias[g], residual=R[g], force_bn=bn,
                                                          act=L.EA_ACT_GEGLU if act == "geglu" else L.EA_ACT_NONE, **kw))
                                  for g in range(G)])
            try:
                for _ in range(5):
                    run()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(40):
                    run()
                e1.record()
                torch.cuda.synchronize()
                us = e0.elapsed_time(e1) * 1000 / 40
            except Exception as e:  # shape / flavour not supported
                print(M, N, K, act, flavour, bn, "n/a", str(e)[:60])
                continue
            tf = 2.0 * M * N * K * G / us / 1e6
            rows.append(dict(M=M, N=N, K=K, act=act, res=res, flavour=flavour, bn=bn, us=round(us, 2), tflops=round(tf, 1)))
            print(f"{M:6d} {N:6d} {K:5d} {act:6s} res={int(res)} {flavour:9s} bn={bn:3d} {us:8.2f} us {tf:7.1f} TFLOP/s", flush=True)
json.dump(rows, open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/gemm_shape_sweep.json", "w"))
