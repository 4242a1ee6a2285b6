"""Host logic: the diffusers<->ldm key map and per-image sharding with a real
world_size-2 gloo group on CPU."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

from editanything_b200 import sharding
from editanything_b200.unet_spec import SD15, SD21, TINY, param_shapes
from editanything_b200.weights import diffusers_to_ldm, key_map


@pytest.mark.parametrize("cfg", [SD15, SD21])
def test_key_map_is_a_bijection_with_known_anchor_names(cfg):
    for kind in ("unet", "controlnet"):
        shapes = param_shapes(cfg, kind)
        km = key_map(cfg, kind, shapes.keys())
        assert len(set(km.values())) == len(km)                   # injective
        sd = {dk: torch.empty(0) for dk in km.values()}
        back = diffusers_to_ldm(sd, kind, cfg)
        assert set(back) == set(shapes)
    km = key_map(cfg, "unet", param_shapes(cfg, "unet").keys())
    # anchors from SURVEY.md Appendix B
    assert km["input_blocks.0.0.weight"] == "conv_in.weight"
    assert km["time_embed.2.bias"] == "time_embedding.linear_2.bias"
    assert km["input_blocks.1.0.in_layers.2.weight"] == "down_blocks.0.resnets.0.conv1.weight"
    assert km["input_blocks.4.0.skip_connection.weight"] == "down_blocks.1.resnets.0.conv_shortcut.weight"
    assert km["input_blocks.5.1.transformer_blocks.0.attn2.to_k.weight"] == \
        "down_blocks.1.attentions.1.transformer_blocks.0.attn2.to_k.weight"
    assert km["input_blocks.6.0.op.weight"] == "down_blocks.1.downsamplers.0.conv.weight"
    assert km["middle_block.2.emb_layers.1.bias"] == "mid_block.resnets.1.time_emb_proj.bias"
    assert km["output_blocks.2.1.conv.weight"] == "up_blocks.0.upsamplers.0.conv.weight"
    assert km["output_blocks.5.2.conv.weight"] == "up_blocks.1.upsamplers.0.conv.weight"
    assert km["output_blocks.11.1.proj_out.bias"] == "up_blocks.3.attentions.2.proj_out.bias"
    assert km["out.2.weight"] == "conv_out.weight"
    kc = key_map(cfg, "controlnet", param_shapes(cfg, "controlnet").keys())
    assert kc["input_hint_block.0.weight"] == "controlnet_cond_embedding.conv_in.weight"
    assert kc["input_hint_block.2.weight"] == "controlnet_cond_embedding.blocks.0.weight"
    assert kc["input_hint_block.14.bias"] == "controlnet_cond_embedding.conv_out.bias"
    assert kc["zero_convs.11.0.weight"] == "controlnet_down_blocks.11.weight"
    assert kc["middle_block_out.0.bias"] == "controlnet_mid_block.bias"


def test_shard_indices_cover_everything_once():
    for n in (0, 1, 5, 8, 33):
        for w in (1, 2, 3, 8):
            seen = sorted(i for r in range(w) for i in sharding.shard_indices(n, r, w))
            assert seen == list(range(n))
    with pytest.raises(ValueError):
        sharding.shard_indices(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_items, q):
    import torch.distributed as dist
    from editanything_b200.denoise import DenoiseEngine, ddim_schedule
    from editanything_b200.unet_spec import make_state_dict
    from oracle.inputs import make_inputs
    from tests import cpu_ops
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    cfg = TINY
    usd = make_state_dict(cfg, "unet", 41)
    csds = [make_state_dict(cfg, "controlnet", 42)]
    eng = DenoiseEngine(cfg, usd, csds, torch.device("cpu"), backend=cpu_ops)
    ts, a, ap = ddim_schedule(50)
    outs = []
    for item in sharding.shard_indices(n_items, rank, world):      # one image per step batch (+CFG)
        x, ctx, hints = make_inputs(cfg, 2, 8, 7, 100 + item, n_controlnets=1)
        eng.prepare(ctx, hints, [0.7])
        eng.begin(x[:1], guidance=7.5, use_graph=False)
        for i in range(2):
            eng.step(int(ts[i]), float(a[i]), float(ap[i]))
        outs.append(eng.latents()[0])
    local = torch.stack(outs) if outs else torch.empty(0, 4, 8, 8)
    full = sharding.gather_sharded(local, n_items, rank, world)
    if rank == 0:
        q.put(full.clone())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_sharded_denoise_matches_single_process():
    n_items = 3                                  # ragged: rank 0 owns 2 images, rank 1 owns 1
    ctx = mp.get_context("spawn")
    res = {}
    for world in (1, 2):
        for attempt in range(2):                 # one retry: the probed port can be taken between probe and bind
            q = ctx.Queue()
            port = _free_port()
            procs = [ctx.Process(target=_worker, args=(r, world, port, n_items, q)) for r in range(world)]
            for p in procs:
                p.start()
            try:
                res[world] = q.get(timeout=240)
            except Exception:
                res[world] = None
            for p in procs:
                p.join(timeout=60)
                if p.is_alive():
                    p.kill()
            if res[world] is not None and all(p.exitcode == 0 for p in procs):
                break
            assert attempt == 0, f"world {world}: worker exit codes {[p.exitcode for p in procs]}"
    assert res[2].shape == (n_items, 4, 8, 8)
    assert torch.equal(res[1], res[2])           # sharding must not change any image
