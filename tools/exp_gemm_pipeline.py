"""Experiment: how K-block rate depends on pipeline depth, tile width, TMA shape (2-D linear vs 4-D
conv boxes) and weight residency (L2-hot vs HBM-cold)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from editanything_b200 import _lib as L, ops  # noqa: E402
from tools.bench_ops import timeit  # noqa: E402

dt = ops.half_dtype()
res = []
for kind, M, H, Cin, N in [("linear", 8192, 0, 2880, 320), ("conv", 8192, 64, 320, 320), ("linear", 2048, 0, 5760, 640),
                           ("conv", 2048, 32, 640, 640), ("linear", 4096, 0, 5120, 1280)]:
    K = Cin if kind == "linear" else 9 * Cin
    for cold in (0, 1):
        n_w = (int(300e6 // (N * K * 2)) + 1) if cold else 1
        ws = [(torch.randn(N, K, device="cuda") / K ** 0.5).to(dt) for _ in range(n_w)]
        if kind == "linear":
            x = torch.randn(M, K, device="cuda").to(dt)
        else:
            x = torch.randn(2, H, H, Cin, device="cuda").to(dt)
        y = torch.empty(M, N, device="cuda", dtype=dt)
        for bn, st in [(128, 3), (128, 6), (160, 3), (160, 5), (256, 4), (64, 4), (64, 8)]:
            it = [0]

            def fn():
                if kind == "linear":
                    ops.gemm(x, ws[it[0] % n_w], y, force_bn=bn, force_stages=st)
                else:
                    ops.gemm(x, ws[it[0] % n_w], y, mode=L.EA_GEMM_CONV_S1, conv=(2, H, H, Cin), force_bn=bn, force_stages=st)
                it[0] += 1
            us = timeit(fn)
            nkb = (K + 63) // 64
            r = {"kind": kind, "M": M, "N": N, "K": K, "cold": cold, "BN": bn, "stages": st, "us": round(us, 1),
                 "tflops": round(2.0 * M * N * K / us / 1e6, 1)}
            res.append(r)
            print(json.dumps(r), flush=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "exp_gemm_pipeline.json"), "w"), indent=1)
