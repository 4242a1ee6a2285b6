"""SAM image-encoder (ViT) configuration + synthetic weights.

The reference depends on the un-vendored, un-pinned `segment_anything` package
(editany_lora.py:37-50, README.md:235) and builds `sam_model_registry["default"]` = ViT-H
(editany_lora.py:82-95).  Parameter names follow upstream segment_anything's
`ImageEncoderViT` state dict (`patch_embed.proj`, `pos_embed`, `blocks.{i}.{norm1,attn.qkv,
attn.proj,attn.rel_pos_h,attn.rel_pos_w,norm2,mlp.lin1,mlp.lin2}`, `neck.{0,1,2,3}`) so a real
`sam_vit_h_4b8939.pth` (keys prefixed `image_encoder.`) loads unchanged.  The in-container HF port
(`transformers/models/sam/modeling_sam.py:700-1075`) is the secondary oracle; key map in
oracle/sam_oracle.py.
"""
from dataclasses import dataclass
from typing import Tuple

import torch


@dataclass(frozen=True)
class SamEncoderConfig:
    img_size: int = 1024
    patch_size: int = 16
    in_chans: int = 3
    embed_dim: int = 1280
    depth: int = 32
    num_heads: int = 16
    mlp_dim: int = 5120
    out_chans: int = 256
    window_size: int = 14
    global_attn_indexes: Tuple[int, ...] = (7, 15, 23, 31)
    ln_eps: float = 1e-6

    @property
    def grid(self):
        return self.img_size // self.patch_size

    @property
    def head_dim(self):
        return self.embed_dim // self.num_heads


SAM_VIT_H = SamEncoderConfig()
# reduced configuration for fast parity tests: same head_dim (80) as ViT-H, a 16x16 token grid with
# 6x6 windows (16 % 6 != 0 -> exercises the zero-padded windows), 2 windowed + 2 global blocks
SAM_TINY = SamEncoderConfig(img_size=256, embed_dim=160, depth=4, num_heads=2, mlp_dim=320, out_chans=64,
                            window_size=6, global_attn_indexes=(1, 3))


def sam_param_shapes(cfg: SamEncoderConfig):
    P = {}
    D, g = cfg.embed_dim, cfg.grid
    P["patch_embed.proj.weight"] = ((D, cfg.in_chans, cfg.patch_size, cfg.patch_size), "w")
    P["patch_embed.proj.bias"] = ((D,), "b")
    P["pos_embed"] = ((1, g, g, D), "pos")
    for i in range(cfg.depth):
        p = f"blocks.{i}"
        S = g if i in cfg.global_attn_indexes else cfg.window_size
        P[p + ".norm1.weight"] = ((D,), "g")
        P[p + ".norm1.bias"] = ((D,), "nb")
        P[p + ".attn.qkv.weight"] = ((3 * D, D), "w")
        P[p + ".attn.qkv.bias"] = ((3 * D,), "b")
        P[p + ".attn.proj.weight"] = ((D, D), "w")
        P[p + ".attn.proj.bias"] = ((D,), "b")
        P[p + ".attn.rel_pos_h"] = ((2 * S - 1, cfg.head_dim), "rel")
        P[p + ".attn.rel_pos_w"] = ((2 * S - 1, cfg.head_dim), "rel")
        P[p + ".norm2.weight"] = ((D,), "g")
        P[p + ".norm2.bias"] = ((D,), "nb")
        P[p + ".mlp.lin1.weight"] = ((cfg.mlp_dim, D), "w")
        P[p + ".mlp.lin1.bias"] = ((cfg.mlp_dim,), "b")
        P[p + ".mlp.lin2.weight"] = ((D, cfg.mlp_dim), "w")
        P[p + ".mlp.lin2.bias"] = ((D,), "b")
    P["neck.0.weight"] = ((cfg.out_chans, D, 1, 1), "w")
    P["neck.1.weight"] = ((cfg.out_chans,), "g")
    P["neck.1.bias"] = ((cfg.out_chans,), "nb")
    P["neck.2.weight"] = ((cfg.out_chans, cfg.out_chans, 3, 3), "w")
    P["neck.3.weight"] = ((cfg.out_chans,), "g")
    P["neck.3.bias"] = ((cfg.out_chans,), "nb")
    return P


def make_sam_state_dict(cfg: SamEncoderConfig, seed: int, dtype=torch.float32, device="cpu"):
    """Deterministic synthetic weights (no checkpoint exists in this environment).  Linear/conv
    weights ~ N(0, 1/fan_in); the relative-position tables and pos_embed (zero-initialised upstream)
    are drawn non-zero so the decomposed rel-pos bias path is actually exercised."""
    g = torch.Generator(device=device).manual_seed(seed)
    sd = {}
    for name, (shape, role) in sam_param_shapes(cfg).items():
        if role == "w":
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            t = torch.randn(shape, generator=g, device=device) * (fan_in ** -0.5)
        elif role == "b":
            t = torch.randn(shape, generator=g, device=device) * 0.05
        elif role == "g":
            t = 1.0 + 0.1 * torch.randn(shape, generator=g, device=device)
        elif role == "nb":
            t = 0.1 * torch.randn(shape, generator=g, device=device)
        elif role == "pos":
            t = 0.5 * torch.randn(shape, generator=g, device=device)
        else:  # rel
            t = 0.15 * torch.randn(shape, generator=g, device=device)
        sd[name] = t.to(dtype)
    return sd
