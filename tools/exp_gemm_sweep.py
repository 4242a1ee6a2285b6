"""Experiment: tile-shape sweep (BN x CTA pairs x stages) for the GEMM shapes that dominate a step
(tools/gemm_breakdown.py), timed as 40 launches over 8 operand sets inside one CUDA graph.
Prints the planner's choice next to the best forced configurations; feeds plan_gemm's constants."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from editanything_b200 import _lib as L, ops  # noqa: E402

dt = ops.half_dtype()
NL, NBUF = 40, 8


def make(M, N, K, conv, residual, act):
    if conv:
        B, H, W, Cin = conv
        M = B * H * W
        xs = [torch.randn(M, Cin, device="cuda").to(dt) for _ in range(NBUF)]
        ws = [(torch.randn(N, 9 * Cin, device="cuda") / (9 * Cin) ** 0.5).to(dt) for _ in range(NBUF)]
    else:
        xs = [torch.randn(M, K, device="cuda").to(dt) for _ in range(NBUF)]
        ws = [(torch.randn(N, K, device="cuda") / K ** 0.5).to(dt) for _ in range(NBUF)]
    n_out = N // 2 if act == L.EA_ACT_GEGLU else N
    ys = [torch.empty(M, n_out, device="cuda", dtype=dt) for _ in range(NBUF)]
    rs = [torch.randn(M, n_out, device="cuda").to(dt) for _ in range(NBUF)] if residual else None
    bs = [torch.randn(N, device="cuda") for _ in range(NBUF)]
    return xs, ws, ys, rs, bs


def time_cfg(bufs, conv, act, **kw):
    xs, ws, ys, rs, bs = bufs

    def one(i):
        j = i % NBUF
        k2 = dict(kw)
        if conv:
            k2.update(mode=L.EA_GEMM_CONV_S1, conv=conv)
        if rs is not None:
            k2["residual"] = rs[j]
        ops.gemm(xs[j], ws[j], ys[j], act=act, bias=bs[j], **k2)

    try:
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            one(0)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(NL):
                one(i)
    except RuntimeError as e:
        return None
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (2 * NL)


if __name__ == "__main__":
    SHAPES = [
        ("lin 8192x320 K320 +res", 8192, 320, 320, None, True, 0),
        ("lin 8192x320 K320", 8192, 320, 320, None, False, 0),
        ("lin 8192x960 K320 (qkv64)", 8192, 960, 320, None, False, 0),
        ("lin 8192x320 K1280 +res (ff2_64)", 8192, 320, 1280, None, True, 0),
        ("geglu 8192x2560 K320", 8192, 2560, 320, None, False, L.EA_ACT_GEGLU),
        ("lin 2048x640 K640 +res", 2048, 640, 640, None, True, 0),
        ("lin 2048x1920 K640 (qkv32)", 2048, 1920, 640, None, False, 0),
        ("lin 2048x640 K2560 +res (ff2_32)", 2048, 640, 2560, None, True, 0),
        ("geglu 2048x5120 K640", 2048, 5120, 640, None, False, L.EA_ACT_GEGLU),
        ("lin 512x1280 K1280 +res", 512, 1280, 1280, None, True, 0),
        ("lin 512x3840 K1280 (qkv16)", 512, 3840, 1280, None, False, 0),
        ("geglu 512x10240 K1280", 512, 10240, 1280, None, False, L.EA_ACT_GEGLU),
        ("lin 128x1280 K1280 +res", 128, 1280, 1280, None, True, 0),
        ("conv64 320->320", 0, 320, 0, (2, 64, 64, 320), False, 0),
        ("conv32 640->640", 0, 640, 0, (2, 32, 32, 640), False, 0),
        ("conv16 1280->1280", 0, 1280, 0, (2, 16, 16, 1280), False, 0),
    ]

    out = []
    for name, M, N, K, conv, res, act in SHAPES:
        bufs = make(M, N, K, conv, res, act)
        auto = time_cfg(bufs, conv, act)
        rows = []
        bns = [128] if act == L.EA_ACT_GEGLU else [b for b in (32, 64, 96, 128, 160, 192, 256) if b < N + 32]
        for bn in bns:
            for two in (-1, 1):
                if two == 1 and bn < 64:
                    continue
                for st in (2, 3, 4, 6, 8):
                    t = time_cfg(bufs, conv, act, force_bn=bn, force_2cta=two, force_stages=st)
                    if t is not None:
                        rows.append((round(t, 2), bn, two, st))
        rows.sort()
        print(f"== {name}: auto {auto:.2f} us; best: " + "  ".join(f"{t}us(bn{bn},two{two},st{st})" for t, bn, two, st in rows[:6]), flush=True)
        out.append({"name": name, "auto_us": auto, "rows": rows})
        del bufs
        torch.cuda.empty_cache()
    if len(sys.argv) > 1:
        json.dump(out, open(sys.argv[1], "w"))
