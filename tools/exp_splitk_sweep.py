"""Experiment: (BN, splits) sweep for the split-K (small-M, weight-streaming) shapes of a step, 40 launches
over 8 operand sets inside one CUDA graph (weights of 8 sets exceed L2 for the conv shapes).  Prints the
planner's choice next to the best forced configurations."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from editanything_b200 import _lib as L, ops  # noqa: E402
from tools.exp_gemm_sweep import make, time_cfg  # noqa: E402

SHAPES = [
    ("conv8 1280->1280 (M=128, kb180)", 0, 1280, 0, (2, 8, 8, 1280), False, 0),
    ("conv8 2560->1280 (M=128, kb360)", 0, 1280, 0, (2, 8, 8, 2560), False, 0),
    ("conv16 1280->1280 (M=512, kb180)", 0, 1280, 0, (2, 16, 16, 1280), False, 0),
    ("conv16 2560->1280 (M=512, kb360)", 0, 1280, 0, (2, 16, 16, 2560), False, 0),
    ("lin 512x1280 K5120 +res (ff2_16)", 512, 1280, 5120, None, True, 0),
    ("lin 128x1280 K5120 +res (ff2_8)", 128, 1280, 5120, None, True, 0),
]
out = []
for name, M, N, K, conv, res, act in SHAPES:
    bufs = make(M, N, K, conv, res, act)
    auto = time_cfg(bufs, conv, act)
    rows = []
    Mrows = conv[0] * conv[1] * conv[2] if conv else M
    for bn in (32, 64, 96, 128, 160, 256):
        tiles = ((Mrows + 127) // 128) * ((N + bn - 1) // bn)
        for sp in (1, 2, 3, 4, 5, 7, 10, 14, 18, 24):
            if tiles * sp > 148:      # the spinning fix-up needs every split CTA resident (1 CTA/SM at these
                continue              # stage counts): a forced config beyond that deadlocks and traps
            for st in (6, 8):
                t = time_cfg(bufs, conv, act, force_bn=bn, force_splits=sp, force_stages=st)
                if t is not None:
                    rows.append((round(t, 2), bn, sp, st))
    rows.sort()
    print(f"== {name}: auto {auto:.2f} us; best: " + "  ".join(f"{t}us(bn{bn},sp{sp},st{st})" for t, bn, sp, st in rows[:8]), flush=True)
    out.append({"name": name, "auto_us": auto, "rows": rows})
    del bufs
    torch.cuda.empty_cache()
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"))
