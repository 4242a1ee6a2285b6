"""Experiment (build with EA_NVCC_EXTRA=-DEA_GEMM_TIMING): where the time of a short GEMM goes when
launches follow each other inside a CUDA graph.  %globaltimer stamps per launch (CTA (0,0,0) phases +
grid-wide first entry / last exit); 40 launches of one shape in one graph, rows 8..36 averaged."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from editanything_b200 import _lib as L, ops  # noqa: E402

dt = ops.half_dtype()
lib = L.load()
lib.ea_gemm_chain_read.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
NL = 40


def run(name, M, N, K, conv=None, residual=False, act=0, nbuf=8, bias=False, interleave=False, **kw):
    if conv:
        B, H, W, Cin = conv
        xs = [torch.randn(B * H * W, Cin, device="cuda").to(dt) for _ in range(nbuf)]
        ws = [(torch.randn(N, 9 * Cin, device="cuda") / (9 * Cin) ** 0.5).to(dt) for _ in range(nbuf)]
        M = B * H * W
    else:
        xs = [torch.randn(M, K, device="cuda").to(dt) for _ in range(nbuf)]
        ws = [(torch.randn(N, K, device="cuda") / K ** 0.5).to(dt) for _ in range(nbuf)]
    n_out = N // 2 if act == L.EA_ACT_GEGLU else N
    ys = [torch.empty(M, n_out, device="cuda", dtype=dt) for _ in range(nbuf)]
    rs = [torch.randn(M, n_out, device="cuda").to(dt) for _ in range(nbuf)] if residual else None
    bs = [torch.randn(N, device="cuda") for _ in range(nbuf)] if bias else None
    if interleave:   # a short 1-wave GEMM between the launches: the measured kernel's CTAs start in lockstep
        ix = torch.randn(8192, 320, device="cuda").to(dt)
        iw = (torch.randn(320, 320, device="cuda") / 18).to(dt)
        iy = torch.empty(8192, 320, device="cuda", dtype=dt)

    def one(i):
        j = i % nbuf
        k2 = dict(kw)
        if conv:
            k2.update(mode=L.EA_GEMM_CONV_S1, conv=conv)
        if residual:
            k2["residual"] = rs[j]
        if bias:
            k2["bias"] = bs[j]
        if interleave:
            ops.gemm(ix, iw, iy)
        ops.gemm(xs[j], ws[j], ys[j], act=act, **k2)

    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        one(0)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    assert lib.ea_gemm_chain_reset() == 0
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(NL):
            one(i)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    tot = NL * (2 if interleave else 1)
    buf = (C.c_ulonglong * (8 * tot))()
    assert lib.ea_gemm_chain_read(buf, tot) == 0
    v = [list(buf[i * 8:(i + 1) * 8]) for i in range(tot)]
    if interleave:
        prev_exit = [v[2 * i][7] for i in range(NL)]
        v = [v[2 * i + 1] for i in range(NL)]
        print(f"   (interleaved: measured kernel's own interval = its last exit - the short kernel's last exit: "
              f"{sum((v[i][7] - prev_exit[i]) / 1e3 for i in range(8, 36)) / 28:.2f} us)")
    lo, hi = 8, 36
    def avg(f):
        return sum(f(i) for i in range(lo, hi)) / (hi - lo)
    print(f"== {name}: {e0.elapsed_time(e1) * 1e3 / NL:.2f} us/launch (events)")
    print("   CTA0: entry->setup %.2f | setup->released %.2f | released->first operands %.2f | main loop %.2f | epilogue %.2f us"
          % tuple(avg(lambda i, a=a, b=b: (v[i][b] - v[i][a]) / 1e3) for a, b in [(0, 1), (1, 2), (2, 3), (3, 4), (4, 5)]))
    print("   grid: first entry -> last exit %.2f us; launch-to-launch (first entries) %.2f us; prev last exit -> this CTA0 released %.2f us; "
          "this first entry - prev last exit %.2f us" % (
              avg(lambda i: (v[i][7] - v[i][6]) / 1e3), avg(lambda i: (v[i][6] - v[i - 1][6]) / 1e3),
              avg(lambda i: (v[i][2] - v[i - 1][7]) / 1e3), avg(lambda i: (v[i][6] - v[i - 1][7]) / 1e3)))
    print("   raw launch 10:", [x - v[10][6] for x in v[10]])


G = L.EA_ACT_GEGLU
if "cold" in sys.argv:
    run("geglu 8192x2560 K320 nbuf8", 8192, 2560, 320, act=G)
    run("geglu 8192x2560 K320 nbuf8 +bias", 8192, 2560, 320, act=G, bias=True)
    run("geglu 8192x2560 K320 nbuf8 +bias interleaved", 8192, 2560, 320, act=G, bias=True, interleave=True)
    run("geglu 8192x2560 K320 nbuf24 +bias interleaved", 8192, 2560, 320, act=G, bias=True, interleave=True, nbuf=24)
    run("lin 512x1280 K1280 nbuf8", 512, 1280, 1280)
    run("lin 512x1280 K1280 nbuf8 +bias", 512, 1280, 1280, bias=True)
    run("lin 512x1280 K1280 nbuf8 +bias interleaved", 512, 1280, 1280, bias=True, interleave=True)
    run("lin 512x1280 K1280 nbuf40 +bias", 512, 1280, 1280, bias=True, nbuf=40)
    run("lin 512x1280 K1280 nbuf40 +bias interleaved", 512, 1280, 1280, bias=True, nbuf=40, interleave=True)
    run("lin 8192x320 K320 nbuf40 +bias +res", 8192, 320, 320, bias=True, residual=True, nbuf=40)
    sys.exit(0)
run("linear 8192x320 K=320", 8192, 320, 320)
run("linear 8192x320 K=320 +residual", 8192, 320, 320, residual=True)
run("linear 8192x320 K=320 no pairs", 8192, 320, 320, force_2cta=-1)
run("linear 8192x2560 K=320 geglu", 8192, 2560, 320, act=L.EA_ACT_GEGLU)
run("linear 512x1280 K=1280", 512, 1280, 1280)
run("linear 2048x640 K=640", 2048, 640, 640)
run("linear 128x1280 K=1280", 128, 1280, 1280)
run("conv 8x8 1280->1280 (split-K)", 0, 1280, 0, conv=(2, 8, 8, 1280))
run("conv 64x64 320->320", 0, 320, 0, conv=(2, 64, 64, 320))
