"""Experiment: split-K configuration sweep (tile width, number of splits, stages) for the
weight-streaming layers (8x8 and 16x16 latents), CUDA-graph timed with HBM-cold weights."""
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
for name, B, H, Cin, Cout in [("conv8 1280", 2, 8, 1280, 1280), ("conv8 2560", 2, 8, 2560, 1280), ("conv16 1280", 2, 16, 1280, 1280),
                              ("conv16 2560", 2, 16, 2560, 1280)]:
    K = 9 * Cin
    n_w = int(300e6 // (Cout * K * 2)) + 1
    ws = [(torch.randn(Cout, K, device="cuda") / K ** 0.5).to(dt) for _ in range(n_w)]
    x = torch.randn(B, H, H, Cin, device="cuda").to(dt)
    y = torch.empty(B * H * H, Cout, device="cuda", dtype=dt)
    mt = B * H * H // 128
    cfgs = [(0, 0, 0)]
    for bn in (64, 128, 256):
        nt = Cout // bn
        for sp in sorted({max(1, 148 // (mt * nt)), max(1, 296 // (mt * nt)), max(1, 74 // (mt * nt))}):
            if sp < 2:
                continue
            for st in (3, 6, 8):
                sb = 16384 + bn * 128
                if st * sb > 224 * 1024:
                    continue
                if mt * nt * sp > (296 if st * sb <= 110 * 1024 else 148):
                    continue
                cfgs.append((bn, sp, st))
    for bn, sp, st in cfgs:
        it = [0]

        def fn():
            ops.gemm(x, ws[it[0] % n_w], y, mode=L.EA_GEMM_CONV_S1, conv=(B, H, H, Cin), force_bn=bn, force_splits=sp,
                     force_stages=st)
            it[0] += 1
        try:
            us = timeit(fn)
        except RuntimeError as e:
            print(name, bn, sp, st, "ERR", str(e)[:60])
            continue
        r = {"case": name, "BN": bn, "splits": sp, "stages": st, "us": round(us, 1), "w_gbs": round(Cout * K * 2 / us / 1e3)}
        res.append(r)
        print(json.dumps(r), flush=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "exp_splitk.json"), "w"), indent=1)
