"""What does the LayerNorm-fold consumer side cost?  Same GEMM with and without `ln=` (3 networks grouped). Dev tool."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch

from editanything_b200 import _lib as L, ops

dev = torch.device("cuda:0")
dt = ops.half_dtype()
G = 3
for (M, N, K, act, bn) in [(8192, 960, 320, "none", 256), (8192, 2560, 320, "geglu", 128), (2048, 1920, 640, "none", 128),
                           (2048, 5120, 640, "geglu", 128), (512, 10240, 1280, "geglu", 128), (8192, 320, 320, "none", 96)]:
    A = [torch.randn(M, K, device=dev).to(dt) for _ in range(G)]
    W = [(torch.randn(N, K, device=dev) * 0.05).to(dt) for _ in range(G)]
    No = N // 2 if act == "geglu" else N
    O = [torch.empty(M, No, device=dev, dtype=dt) for _ in range(G)]
    bias = [torch.randn(N, device=dev) for _ in range(G)]
    stats = [torch.rand(K // 32, M, 2, device=dev) + 1.0 for _ in range(G)]
    gvec = [torch.randn(N, device=dev) for _ in range(G)]
    for name, extra in (("plain", lambda g: {}), ("ln", lambda g: dict(ln=(stats[g], gvec[g], 1e-5)))):
        def run():
            ops.gemm_grouped([(A[g], W[g], O[g], dict(bias=bias[g], force_bn=bn, force_persistent=2,
                                                      act=L.EA_ACT_GEGLU if act == "geglu" else L.EA_ACT_NONE, **extra(g)))
                              for g in range(G)])
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
        print(f"{M:6d} {N:6d} {K:5d} {act:6s} bn={bn:3d} {name:6s} {us:8.2f} us {2.0 * M * N * K * G / us / 1e6:7.1f} TFLOP/s", flush=True)
