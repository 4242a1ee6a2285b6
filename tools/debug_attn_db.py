import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from editanything_b200 import ops
dt = ops.half_dtype()
torch.manual_seed(0)
for (B, h, Nq, Nkv, d) in [(1, 1, 256, 384, 64), (1, 1, 256, 384, 40), (1, 2, 256, 480, 40), (1, 1, 256, 768, 64), (2, 8, 512, 1000, 40)]:
    C_ = h * d
    q = torch.randn(B, Nq, C_, device="cuda").to(dt)
    k = torch.randn(B, Nkv, C_, device="cuda").to(dt)
    v = torch.randn(B, Nkv, C_, device="cuda").to(dt)
    o = torch.zeros(B, Nq, C_, device="cuda", dtype=dt)
    try:
        ops.attention(q, k, v, o, B=B, heads=h, Nq=Nq, Nkv=Nkv, d=d, q_strides=(Nq * C_, C_), k_strides=(Nkv * C_, C_),
                      v_strides=(Nkv * C_, C_), o_strides=(Nq * C_, C_), scale=d ** -0.5)
        torch.cuda.synchronize()
    except Exception as e:
        print((B, h, Nq, Nkv, d), "EXC", str(e)[:80]); break
    qf = q.float().reshape(B, Nq, h, d).permute(0, 2, 1, 3); kf = k.float().reshape(B, Nkv, h, d).permute(0, 2, 1, 3)
    vf = v.float().reshape(B, Nkv, h, d).permute(0, 2, 1, 3)
    ref = ((qf @ kf.transpose(-1, -2) * d ** -0.5).softmax(-1) @ vf).permute(0, 2, 1, 3).reshape(B, Nq, C_)
    err = (o.float() - ref)
    print((B, h, Nq, Nkv, d), "nan", int(torch.isnan(o.float()).sum()), "max_err tile0", float(err[:, :128].abs().nan_to_num(9).max()),
          "tile1", float(err[:, 128:256].abs().nan_to_num(9).max()), "ref_max", float(ref.abs().max()))
    # partial sums: emulate using only first n key tiles to see how many tiles were accumulated
    for nt in (1, 2, 3, 4):
        kk = min(Nkv, nt * 96)
        s = (qf[..., :, :] @ kf[..., :kk, :].transpose(-1, -2)) * d ** -0.5
        pr = (s.softmax(-1) @ vf[..., :kk, :]).permute(0, 2, 1, 3).reshape(B, Nq, C_)
        print("    vs first", nt, "tiles only:", float((o.float() - pr).abs().nan_to_num(9).max()))
