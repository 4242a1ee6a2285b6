"""Experiment (needs a build with EA_NVCC_EXTRA=-DEA_ATTN_TIMING): per-phase cycle stamps of the
softmax warpgroups of one CTA of the 64x64 self-attention (N=4096, d=40)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from editanything_b200 import _lib as L, ops  # noqa: E402

dt = ops.half_dtype()
B, h, N, d = 2, 8, 4096, 40
C_ = h * d
q = torch.randn(B, N, C_, device="cuda").to(dt)
k = torch.randn(B, N, C_, device="cuda").to(dt)
v = torch.randn(B, N, C_, device="cuda").to(dt)
o = torch.empty(B, N, C_, device="cuda", dtype=dt)
for _ in range(3):
    ops.attention(q, k, v, o, B=B, heads=h, Nq=N, Nkv=N, d=d, q_strides=(N * C_, C_), k_strides=(N * C_, C_),
                  v_strides=(N * C_, C_), o_strides=(N * C_, C_), scale=d ** -0.5)
torch.cuda.synchronize()
lib = L.load()
buf = (C.c_longlong * 128)()
lib.ea_attn_debug_read.argtypes = [C.POINTER(C.c_longlong), C.c_int]
assert lib.ea_attn_debug_read(buf, 128) == 0
vals = list(buf)
base = min(x for x in vals if x > 0)
for t in range(2):
    print(f"-- tile {'AB'[t]} (cycles since first stamp)")
    for j in range(8):
        row = vals[(t * 8 + j) * 8:(t * 8 + j) * 8 + 6]
        ds = [row[i + 1] - row[i] for i in range(5)]
        print(f"  kv{j + 8}: top@{row[0] - base:7d}  wait_S {ds[0]:5d}  ld {ds[1]:5d}  max {ds[2]:5d}  exp {ds[3]:5d}  publish {ds[4]:5d}")

mb = (C.c_longlong * 64)()
lib.ea_attn_debug_read_mma.argtypes = [C.POINTER(C.c_longlong), C.c_int]
if lib.ea_attn_debug_read_mma(mb, 64) == 0:
    mv = list(mb)
    print("-- MMA thread (cycles since first softmax stamp): p_ready seen | PV issued (+d) | S issued (+d)")
    for t in range(2):
        for j in range(4):
            r = mv[(t * 8 + j) * 4:(t * 8 + j) * 4 + 3]
            pub = vals[(t * 8 + j) * 8 + 5]
            nxt = vals[(t * 8 + j + 1) * 8 + 1]
            print(f"  tile {'AB'[t]} kv{j + 8}: softmax published@{pub - base:6d}  mma saw@{r[0] - base:6d}  PV issue {r[1] - r[0]:5d}  S issue {r[2] - r[1]:5d}  S(j+1) visible to softmax@{nxt - base:6d} (+{nxt - r[2]} after issue)")
