"""Host-side graph logic of editanything_b200.nets / denoise against the REFERENCE-generated
golden vectors, with the operators emulated in fp32 on CPU (tests/cpu_ops.py)."""
import os

import pytest
import torch

from editanything_b200.denoise import DenoiseEngine, ddim_schedule
from editanything_b200.unet_spec import TINY, TINY21, make_state_dict
from oracle.inputs import make_inputs
from tests import cpu_ops

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CFGS = {"tiny_sd15": TINY, "tiny_sd21": TINY21, "tiny_sd15_32": TINY}


def _engine(name):
    g = torch.load(os.path.join(GOLD, name + ".pt"))
    m = g["meta"]
    cfg = CFGS[name]
    usd = make_state_dict(cfg, "unet", m["unet_seed"])
    csds = [make_state_dict(cfg, "controlnet", s) for s in m["cn_seeds"]]
    eng = DenoiseEngine(cfg, usd, csds, torch.device("cpu"), backend=cpu_ops)
    x, ctx, hints = make_inputs(cfg, m["B"], m["lat"], m["L"], m["in_seed"], n_controlnets=len(csds))
    eng.prepare(ctx, hints, m["scales"])
    return g, m, eng, x


@pytest.mark.parametrize("name", ["tiny_sd15", "tiny_sd21"])
def test_eps_matches_reference_golden(name):
    g, m, eng, x = _engine(name)
    for t in m["timesteps"]:
        eps = eng.eps(x, t)
        ref = g[f"eps_t{t}"]
        assert (eps - ref).abs().max().item() < 2e-4 * max(1.0, ref.abs().max().item())


def test_fused_step_matches_ddim_formula():
    g, m, eng, x = _engine("tiny_sd15")
    t = m["timesteps"][0]
    ref = g[f"eps_t{t}"]
    ts, a, ap = ddim_schedule(50)
    i = list(ts).index(t)
    lat0 = x[:1].clone()
    xx = torch.cat([lat0, lat0])
    eps = eng.eps(xx, t)
    eng.begin(lat0, guidance=9.0, use_graph=False)
    eng.step(t, a[i], ap[i])
    e = eps[:1] + 9.0 * (eps[1:] - eps[:1])
    x0 = (lat0 - (1 - a[i]) ** 0.5 * e) / a[i] ** 0.5
    xp = ap[i] ** 0.5 * x0 + (1 - ap[i]) ** 0.5 * e
    assert (eng.latents() - xp).abs().max().item() < 1e-4


def test_guess_mode_and_spatial_scale_map_match_oracle():
    """ControlNetModel2.forward scaling variants (utils/stable_diffusion_controlnet.py:777-802): guess mode =
    logspace(-1, 0, 13) per-residual factors on top of the scale; a tensor scale = a spatial map resized bilinearly
    (align_corners=True) to every residual's resolution (ea_gemm_args.row_scale)."""
    import torch
    from editanything_b200.denoise import DenoiseEngine
    from editanything_b200.unet_spec import TINY, build_topology, make_state_dict
    from oracle import unet_oracle as O
    from oracle.inputs import make_inputs
    from tests import cpu_ops
    cfg = TINY
    usd = make_state_dict(cfg, "unet", 41)
    csds = [make_state_dict(cfg, "controlnet", 42), make_state_dict(cfg, "controlnet", 43)]
    x, ctx, hints = make_inputs(cfg, 2, 16, 11, 3)
    eng = DenoiseEngine(cfg, usd, csds, torch.device("cpu"), backend=cpu_ops)
    ut, ct = build_topology(cfg), build_topology(cfg, with_decoder=False)
    t = 601
    # guess mode
    eng.prepare(ctx, hints, [0.7, 1.0], guess_mode=True)
    with torch.no_grad():
        ref = O.apply_model(usd, ut, [(sd, ct) for sd in csds], x, torch.full((2,), t), ctx, hints, [0.7, 1.0], guess_mode=True)
        plain = O.apply_model(usd, ut, [(sd, ct) for sd in csds], x, torch.full((2,), t), ctx, hints, [0.7, 1.0])
    got = eng.eps(x, t)
    assert (got - ref).abs().max().item() < 2e-4 and (got - plain).abs().max().item() > 1e-3
    # spatial map on the first net, plain float on the second
    g = torch.Generator().manual_seed(9)
    smap = torch.rand(16, 16, generator=g)
    eng.prepare(ctx, hints, [smap, 0.5])
    with torch.no_grad():
        ref = O.apply_model(usd, ut, [(sd, ct) for sd in csds], x, torch.full((2,), t), ctx, hints, [smap, 0.5])
    assert (eng.eps(x, t) - ref).abs().max().item() < 2e-4
    # back to plain floats
    eng.prepare(ctx, hints, [0.7, 1.0])
    assert (eng.eps(x, t) - plain).abs().max().item() < 2e-4


def test_graph_cache_is_a_bounded_lru():
    from editanything_b200._graphs import GraphLRU
    c = GraphLRU(2)
    c.put("a", 1), c.put("b", 2)
    assert c.get("a") == 1            # refreshes "a"
    c.put("c", 3)                     # evicts "b", the least recently used
    assert "b" not in c and "a" in c and "c" in c and len(c) == 2 and c.get("b") is None
