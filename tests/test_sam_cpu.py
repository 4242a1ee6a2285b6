"""CPU suite for the SAM image encoder: the oracle restatement against the HF port (run live —
transformers is part of the image) and against the HF-generated golden vectors; the host-side
graph of editanything_b200.sam with the operators emulated on CPU (tests/cpu_ops.py)."""
import os

import torch

from editanything_b200.sam import SamEncoderEngine
from editanything_b200.sam_spec import SAM_TINY, SAM_VIT_H, make_sam_state_dict, sam_param_shapes
from oracle import sam_oracle as S
from oracle.make_golden_sam import make_image
from tests import cpu_ops

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_param_count_matches_vit_h():
    n = 0
    for shape, _ in sam_param_shapes(SAM_VIT_H).values():
        k = 1
        for s in shape:
            k *= s
        n += k
    assert abs(n - 637.0e6) < 0.1e6, n      # 637.0 M


def test_oracle_matches_live_hf_port():
    cfg = SAM_TINY
    sd = make_sam_state_dict(cfg, 9)
    img = make_image(cfg, 1, 3)
    with torch.no_grad():
        ours = S.image_encoder(sd, cfg, img)
        ref = S.hf_encoder(cfg, sd)(img).last_hidden_state
    assert (ours - ref).abs().max().item() < 5e-5


def test_oracle_reproduces_hf_golden():
    g = torch.load(os.path.join(GOLD, "sam_tiny.pt"))
    m = g["meta"]
    sd = make_sam_state_dict(SAM_TINY, m["weight_seed"])
    with torch.no_grad():
        out = S.image_encoder(sd, SAM_TINY, make_image(SAM_TINY, m["B"], m["image_seed"]))
    assert (out - g["embedding"]).abs().max().item() < 5e-5


def test_host_graph_matches_golden_with_emulated_ops():
    g = torch.load(os.path.join(GOLD, "sam_tiny.pt"))
    m = g["meta"]
    sd = make_sam_state_dict(SAM_TINY, m["weight_seed"])
    eng = SamEncoderEngine(SAM_TINY, {"image_encoder." + k: v for k, v in sd.items()}, torch.device("cpu"),
                           backend=cpu_ops)
    out = eng.encode(make_image(SAM_TINY, m["B"], m["image_seed"]))
    assert out.shape == g["embedding"].shape
    assert (out - g["embedding"]).abs().max().item() < 2e-4
