"""Experiment (build with EA_NVCC_EXTRA=-DEA_GEMM_TIMING): per-K-block clock stamps of CTA (0,0,0) of a
linear GEMM: when the producer finds a free stage / has issued TMA, when the MMA thread sees the
operands / has issued MMAs+commit."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from editanything_b200 import _lib as L, ops  # noqa: E402

dt = ops.half_dtype()
M, N, K = 8192, 320, 2880
x = torch.randn(M, K, device="cuda").to(dt)
w = (torch.randn(N, K, device="cuda") / K ** 0.5).to(dt)
y = torch.empty(M, N, device="cuda", dtype=dt)
lib = L.load()
lib.ea_gemm_debug_read.argtypes = [C.POINTER(C.c_longlong), C.c_int]
for bn, st, two in [(160, 3, 0), (160, 6, 0), (128, 3, 0), (128, 6, 0), (256, 4, 0), (160, 6, 1), (256, 4, 1)]:
    for _ in range(3):
        ops.gemm(x, w, y, force_bn=bn, force_stages=st, force_2cta=two)
    torch.cuda.synchronize()
    buf = (C.c_longlong * 264)()
    assert lib.ea_gemm_debug_read(buf, 264) == 0
    v = list(buf)
    t0 = v[256]
    print(f"== BN={bn} stages={st} two={two}: entry->setup {v[257] - t0}, accumulator done @{v[262] - t0}, epilogue done @{v[263] - t0}")
    rows = []
    for kb in range(45):
        r = v[kb * 4:kb * 4 + 4]
        rows.append([q - t0 for q in r])
    for kb in (0, 1, 2, 3, 20, 21, 44):
        r = rows[kb]
        print(f"   kb{kb:2d}: stage free@{r[0]:6d}  tma issued +{r[1] - r[0]:4d}  landed@{r[2]:6d} (+{r[2] - r[1]:5d} after issue)  mma issued +{r[3] - r[2]:4d}")
    per = (rows[44][2] - rows[8][2]) / 36.0
    print(f"   steady state: {per:.0f} cycles per K-block (MMA ideal {2 * bn})")
