"""Shapes of the first-stage (AutoencoderKL) decoder the reference decodes latents with
(SURVEY.md §8a R17, §8f N1): `decode_latents` utils/stable_diffusion_controlnet_inpaint.py:718-724 →
`vae.decode`; same network as ldm `Decoder` (ldm/modules/diffusionmodules/model.py:546-652) behind
`post_quant_conv` (ldm/models/autoencoder.py `AutoencoderKL.decode`).  Parameter names are the ldm
ones (`decoder.*`, `post_quant_conv.*`); editanything_b200.weights.vae_diffusers_to_ldm maps the
diffusers spelling onto them.
"""
from dataclasses import dataclass
from typing import Tuple

import torch


@dataclass(frozen=True)
class VaeConfig:
    ch: int = 128
    out_ch: int = 3
    ch_mult: Tuple[int, ...] = (1, 2, 4, 4)
    num_res_blocks: int = 2
    z_channels: int = 4
    embed_dim: int = 4
    scaling_factor: float = 0.18215
    num_groups: int = 32
    eps: float = 1e-6

    @property
    def num_resolutions(self):
        return len(self.ch_mult)

    @property
    def block_out_channels(self):       # diffusers' vae.config.block_out_channels (pipeline reads its length)
        return tuple(self.ch * m for m in self.ch_mult)


SD_VAE = VaeConfig()                                                       # kl-f8 (SD1.5 / SD2.1)
VAE_TINY = VaeConfig(ch=64, ch_mult=(1, 2), num_res_blocks=1)              # test-sized, same code paths (conv Cin % 64 == 0)


def decoder_blocks(cfg: VaeConfig):
    """[(kind, prefix, cin, cout)] in execution order (model.py:623-652): mid res / attn / res, then per
    level (highest first) num_res_blocks+1 ResnetBlocks and, except at level 0, Upsample."""
    out = []
    block_in = cfg.ch * cfg.ch_mult[-1]
    out += [("res", "decoder.mid.block_1", block_in, block_in), ("attn", "decoder.mid.attn_1", block_in, block_in),
            ("res", "decoder.mid.block_2", block_in, block_in)]
    for lvl in reversed(range(cfg.num_resolutions)):
        block_out = cfg.ch * cfg.ch_mult[lvl]
        for i in range(cfg.num_res_blocks + 1):
            out.append(("res", f"decoder.up.{lvl}.block.{i}", block_in, block_out))
            block_in = block_out
        if lvl != 0:
            out.append(("up", f"decoder.up.{lvl}.upsample", block_in, block_in))
    return out, block_in


def encoder_blocks(cfg: VaeConfig):
    """[(kind, prefix, cin, cout)] in execution order (Encoder.forward, model.py:519-543): per level
    num_res_blocks ResnetBlocks and, except at the last level, Downsample; then mid res / attn / res."""
    out = []
    block_in = cfg.ch
    for lvl in range(cfg.num_resolutions):
        block_out = cfg.ch * cfg.ch_mult[lvl]
        for i in range(cfg.num_res_blocks):
            out.append(("res", f"encoder.down.{lvl}.block.{i}", block_in, block_out))
            block_in = block_out
        if lvl != cfg.num_resolutions - 1:
            out.append(("down", f"encoder.down.{lvl}.downsample", block_in, block_in))
    out += [("res", "encoder.mid.block_1", block_in, block_in), ("attn", "encoder.mid.attn_1", block_in, block_in),
            ("res", "encoder.mid.block_2", block_in, block_in)]
    return out, block_in


def _block_param_shapes(ps, blocks):
    for kind, p, cin, cout in blocks:
        if kind == "res":
            ps[p + ".norm1.weight"], ps[p + ".norm1.bias"] = ((cin,), "g"), ((cin,), "nb")
            ps[p + ".conv1.weight"], ps[p + ".conv1.bias"] = ((cout, cin, 3, 3), "w"), ((cout,), "b")
            ps[p + ".norm2.weight"], ps[p + ".norm2.bias"] = ((cout,), "g"), ((cout,), "nb")
            ps[p + ".conv2.weight"], ps[p + ".conv2.bias"] = ((cout, cout, 3, 3), "w"), ((cout,), "b")
            if cin != cout:
                ps[p + ".nin_shortcut.weight"], ps[p + ".nin_shortcut.bias"] = ((cout, cin, 1, 1), "w"), ((cout,), "b")
        elif kind == "attn":
            ps[p + ".norm.weight"], ps[p + ".norm.bias"] = ((cin,), "g"), ((cin,), "nb")
            for n in ("q", "k", "v", "proj_out"):
                ps[f"{p}.{n}.weight"], ps[f"{p}.{n}.bias"] = ((cin, cin, 1, 1), "w"), ((cin,), "b")
        else:   # up: upsample.conv, down: downsample.conv
            ps[p + ".conv.weight"], ps[p + ".conv.bias"] = ((cin, cin, 3, 3), "w"), ((cin,), "b")


def vae_encoder_param_shapes(cfg: VaeConfig, in_channels=3):
    """quant_conv + Encoder (double_z): name -> (shape, role)."""
    ps = {"encoder.conv_in.weight": ((cfg.ch, in_channels, 3, 3), "w"), "encoder.conv_in.bias": ((cfg.ch,), "b")}
    blocks, last = encoder_blocks(cfg)
    _block_param_shapes(ps, blocks)
    ps["encoder.norm_out.weight"], ps["encoder.norm_out.bias"] = ((last,), "g"), ((last,), "nb")
    ps["encoder.conv_out.weight"] = ((2 * cfg.z_channels, last, 3, 3), "w")
    ps["encoder.conv_out.bias"] = ((2 * cfg.z_channels,), "b")
    ps["quant_conv.weight"] = ((2 * cfg.embed_dim, 2 * cfg.z_channels, 1, 1), "w")
    ps["quant_conv.bias"] = ((2 * cfg.embed_dim,), "b")
    return ps


def vae_decoder_param_shapes(cfg: VaeConfig):
    """name -> (shape, role); role in {w, b, g (norm scale), nb (norm shift)}."""
    ps = {"post_quant_conv.weight": ((cfg.z_channels, cfg.embed_dim, 1, 1), "w"),
          "post_quant_conv.bias": ((cfg.z_channels,), "b")}
    top = cfg.ch * cfg.ch_mult[-1]
    ps["decoder.conv_in.weight"] = ((top, cfg.z_channels, 3, 3), "w")
    ps["decoder.conv_in.bias"] = ((top,), "b")
    blocks, last = decoder_blocks(cfg)
    _block_param_shapes(ps, blocks)
    ps["decoder.norm_out.weight"], ps["decoder.norm_out.bias"] = ((last,), "g"), ((last,), "nb")
    ps["decoder.conv_out.weight"] = ((cfg.out_ch, last, 3, 3), "w")
    ps["decoder.conv_out.bias"] = ((cfg.out_ch,), "b")
    return ps


def make_vae_state_dict(cfg: VaeConfig, seed: int, dtype=torch.float32, device="cpu", part="decoder"):
    """Deterministic synthetic weights (no checkpoint in this environment) of the decoder side
    (`part="decoder"`: post_quant_conv + decoder) or the encoder side (`part="encoder"`: encoder +
    quant_conv): conv weights ~ N(0, 1/fan_in), small biases, norm scales around 1 — activations stay
    O(1) through the stack."""
    g = torch.Generator(device=device).manual_seed(seed)
    sd = {}
    shapes = vae_decoder_param_shapes(cfg) if part == "decoder" else vae_encoder_param_shapes(cfg)
    for name, (shape, role) in shapes.items():
        if role == "w":
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            t = torch.randn(shape, generator=g, device=device) * (fan_in ** -0.5)
        elif role == "b":
            t = torch.randn(shape, generator=g, device=device) * 0.05
        elif role == "g":
            t = 1.0 + 0.1 * torch.randn(shape, generator=g, device=device)
        else:
            t = 0.1 * torch.randn(shape, generator=g, device=device)
        sd[name] = t.to(dtype)
    return sd
