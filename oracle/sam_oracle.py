"""CPU ORACLE for the SAM ViT image encoder (test infrastructure — NOT a product path).

Plain-PyTorch fp32 functional restatement of upstream `segment_anything/modeling/image_encoder.py`
(`ImageEncoderViT`, `Block`, `Attention`, `window_partition/unpartition`, `get_rel_pos`,
`add_decomposed_rel_pos`, `PatchEmbed`, `LayerNorm2d`), which the reference calls through
`SamAutomaticMaskGenerator.generate` (editany_lora.py:82-95,522-525).  segment_anything is NOT
vendored in /root/reference and not installed here (un-pinned git dependency, README.md:235), so
each function instead cites the in-container HF port of the same code,
`transformers/models/sam/modeling_sam.py` (abbreviated HF:line).

Parity status: PINNED against the HF port — `tests/test_sam_oracle.py` runs
`transformers.SamVisionEncoder` with the same weights (key map `to_hf_state_dict`) and requires
agreement to fp32 round-off; `tests/golden/sam_*.pt` hold HF-generated outputs
(oracle/make_golden_sam.py).  Against upstream segment_anything itself: unpinned (absent).
"""
import torch
import torch.nn.functional as F


def get_rel_pos(q_size, k_size, rel_pos):
    """HF:729-759.  Tables have length 2*S-1 for S = q_size = k_size (no interpolation needed)."""
    max_rel_dist = int(2 * max(q_size, k_size) - 1)
    if rel_pos.shape[0] != max_rel_dist:
        rel_pos = F.interpolate(rel_pos.reshape(1, rel_pos.shape[0], -1).permute(0, 2, 1), size=max_rel_dist,
                                mode="linear").reshape(-1, max_rel_dist).permute(1, 0)
    q_coords = torch.arange(q_size)[:, None] * max(k_size / q_size, 1.0)
    k_coords = torch.arange(k_size)[None, :] * max(q_size / k_size, 1.0)
    rel = (q_coords - k_coords) + (k_size - 1) * max(q_size / k_size, 1.0)
    return rel_pos[rel.long()]          # [q, k, d]


def attention(x, sd, p, heads):
    """Attention.forward with decomposed rel-pos (HF:762-831): x [B, H, W, C]."""
    B, H, W, C = x.shape
    d = C // heads
    qkv = F.linear(x, sd[p + ".qkv.weight"], sd[p + ".qkv.bias"]).reshape(B, H * W, 3, heads, d).permute(2, 0, 3, 1, 4)
    q, k, v = qkv.reshape(3, B * heads, H * W, d).unbind(0)
    attn = (q * d ** -0.5) @ k.transpose(-2, -1)
    Rh = get_rel_pos(H, H, sd[p + ".rel_pos_h"])
    Rw = get_rel_pos(W, W, sd[p + ".rel_pos_w"])
    rq = q.reshape(B * heads, H, W, d)                       # NB: the UNSCALED q (HF:812-816)
    rel_h = torch.einsum("bhwc,hkc->bhwk", rq, Rh)
    rel_w = torch.einsum("bhwc,wkc->bhwk", rq, Rw)
    attn = (attn.view(B * heads, H, W, H, W) + rel_h[:, :, :, :, None] + rel_w[:, :, :, None, :]).view(
        B * heads, H * W, H * W)
    attn = attn.softmax(dim=-1)
    out = (attn @ v).view(B, heads, H, W, d).permute(0, 2, 3, 1, 4).reshape(B, H, W, C)
    return F.linear(out, sd[p + ".proj.weight"], sd[p + ".proj.bias"])


def window_partition(x, ws):
    """HF:900-922 — zero pad (after LayerNorm) to a multiple of ws, then tile."""
    B, H, W, C = x.shape
    ph, pw = (ws - H % ws) % ws, (ws - W % ws) % ws
    x = F.pad(x, (0, 0, 0, pw, 0, ph))
    Hp, Wp = H + ph, W + pw
    x = x.view(B, Hp // ws, ws, Wp // ws, ws, C).permute(0, 1, 3, 2, 4, 5).reshape(-1, ws, ws, C)
    return x, (Hp, Wp)


def window_unpartition(w, ws, pad_hw, hw):
    """HF:924-952."""
    Hp, Wp = pad_hw
    H, W = hw
    B = w.shape[0] // (Hp * Wp // ws // ws)
    x = w.view(B, Hp // ws, Wp // ws, ws, ws, -1).permute(0, 1, 3, 2, 4, 5).reshape(B, Hp, Wp, -1)
    return x[:, :H, :W, :]


def block(x, sd, p, heads, ws, eps):
    """Block.forward (HF:954-972)."""
    shortcut = x
    x = F.layer_norm(x, (x.shape[-1],), sd[p + ".norm1.weight"], sd[p + ".norm1.bias"], eps)
    if ws > 0:
        H, W = x.shape[1], x.shape[2]
        x, pad_hw = window_partition(x, ws)
    x = attention(x, sd, p + ".attn", heads)
    if ws > 0:
        x = window_unpartition(x, ws, pad_hw, (H, W))
    x = shortcut + x
    y = F.layer_norm(x, (x.shape[-1],), sd[p + ".norm2.weight"], sd[p + ".norm2.bias"], eps)
    y = F.linear(F.gelu(F.linear(y, sd[p + ".mlp.lin1.weight"], sd[p + ".mlp.lin1.bias"])),
                 sd[p + ".mlp.lin2.weight"], sd[p + ".mlp.lin2.bias"])            # exact GELU, HF:132-143
    return x + y


def layernorm2d(x, w, b, eps):
    """LayerNorm2d over channels of NCHW (HF SamLayerNorm channels_first)."""
    u = x.mean(1, keepdim=True)
    s = (x - u).pow(2).mean(1, keepdim=True)
    return w[None, :, None, None] * ((x - u) / torch.sqrt(s + eps)) + b[None, :, None, None]


def image_encoder(sd, cfg, img, return_tokens=False):
    """ImageEncoderViT.forward (HF:1050-1075): img fp32 [B, 3, S, S] (already normalised and
    padded by Sam.preprocess) -> [B, out_chans, S/16, S/16]."""
    x = F.conv2d(img, sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"], stride=cfg.patch_size)
    x = x.permute(0, 2, 3, 1) + sd["pos_embed"]
    for i in range(cfg.depth):
        ws = 0 if i in cfg.global_attn_indexes else cfg.window_size
        x = block(x, sd, f"blocks.{i}", cfg.num_heads, ws, cfg.ln_eps)
    tokens = x
    x = F.conv2d(x.permute(0, 3, 1, 2), sd["neck.0.weight"])
    x = layernorm2d(x, sd["neck.1.weight"], sd["neck.1.bias"], 1e-6)
    x = F.conv2d(x, sd["neck.2.weight"], padding=1)
    x = layernorm2d(x, sd["neck.3.weight"], sd["neck.3.bias"], 1e-6)
    return (x, tokens) if return_tokens else x


# --------------------------------------------------------------------------- HF pinning
def to_hf_state_dict(sd):
    """segment_anything ImageEncoderViT names -> transformers SamVisionEncoder names."""
    out = {}
    for k, v in sd.items():
        k2 = k.replace("patch_embed.proj", "patch_embed.projection")
        k2 = k2.replace("blocks.", "layers.").replace(".norm1.", ".layer_norm1.").replace(".norm2.", ".layer_norm2.")
        k2 = (k2.replace("neck.0.", "neck.conv1.").replace("neck.1.", "neck.layer_norm1.")
              .replace("neck.2.", "neck.conv2.").replace("neck.3.", "neck.layer_norm2."))
        out[k2] = v
    return out


def hf_encoder(cfg, sd):
    """transformers.SamVisionEncoder carrying the given weights (eager attention, fp32)."""
    from transformers import SamVisionConfig
    from transformers.models.sam.modeling_sam import SamVisionEncoder
    hc = SamVisionConfig(hidden_size=cfg.embed_dim, output_channels=cfg.out_chans, num_hidden_layers=cfg.depth,
                         num_attention_heads=cfg.num_heads, num_channels=cfg.in_chans, image_size=cfg.img_size,
                         patch_size=cfg.patch_size, layer_norm_eps=cfg.ln_eps, window_size=cfg.window_size,
                         global_attn_indexes=list(cfg.global_attn_indexes), mlp_dim=cfg.mlp_dim,
                         attn_implementation="eager")
    m = SamVisionEncoder(hc).eval()
    missing, unexpected = m.load_state_dict(to_hf_state_dict(sd), strict=True)
    return m
