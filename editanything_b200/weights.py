"""State-dict key maps between diffusers (`UNet2DConditionModel` / `ControlNetModel`, the layout of
the checkpoints the reference downloads: editany_lora.py:72-79,360,372-377) and the ldm/cldm naming
this package (and the reference's vendored `cldm/`, `ldm/`) uses.

The two are the same network (the reference converts with diffusers' own converter,
tools/convert_controlnet_to_diffusers.py:19,80-91); the table is SURVEY.md Appendix B, generated here
from the block topology instead of being written out by hand.
"""
from .unet_spec import HINT_STRIDES, UNetConfig, build_topology

_RES = {"in_layers.0": "norm1", "in_layers.2": "conv1", "emb_layers.1": "time_emb_proj",
        "out_layers.0": "norm2", "out_layers.3": "conv2", "skip_connection": "conv_shortcut"}


def ldm_prefix_map(cfg: UNetConfig, kind: str):
    """{ldm module prefix: diffusers module prefix} for kind in {'unet', 'controlnet'}."""
    topo = build_topology(cfg, with_decoder=(kind == "unet"))
    m = {"time_embed.0": "time_embedding.linear_1", "time_embed.2": "time_embedding.linear_2"}
    level, j = 0, 0
    for layers in topo.input_blocks:
        for b in layers:
            if b.kind == "conv_in":
                m[b.prefix] = "conv_in"
            elif b.kind == "res":
                m[b.prefix] = f"down_blocks.{level}.resnets.{j}"
            elif b.kind == "attn":
                m[b.prefix] = f"down_blocks.{level}.attentions.{j}"
            elif b.kind == "down":
                m[b.prefix + ".op"] = f"down_blocks.{level}.downsamplers.0.conv"
        kinds = [b.kind for b in layers]
        if "res" in kinds:
            j += 1
        if "down" in kinds:
            level, j = level + 1, 0
    m["middle_block.0"], m["middle_block.1"], m["middle_block.2"] = (
        "mid_block.resnets.0", "mid_block.attentions.0", "mid_block.resnets.1")
    if kind == "unet":
        per = cfg.num_res_blocks + 1
        for oi, layers in enumerate(topo.output_blocks):
            i, jj = oi // per, oi % per
            for b in layers:
                if b.kind == "res":
                    m[b.prefix] = f"up_blocks.{i}.resnets.{jj}"
                elif b.kind == "attn":
                    m[b.prefix] = f"up_blocks.{i}.attentions.{jj}"
                elif b.kind == "up":
                    m[b.prefix + ".conv"] = f"up_blocks.{i}.upsamplers.0.conv"
        m["out.0"], m["out.2"] = "conv_norm_out", "conv_out"
    else:
        n = len(HINT_STRIDES)
        for k in range(n):
            src = f"input_hint_block.{2 * k}"
            m[src] = ("controlnet_cond_embedding.conv_in" if k == 0 else
                      "controlnet_cond_embedding.conv_out" if k == n - 1 else
                      f"controlnet_cond_embedding.blocks.{k - 1}")
        for i in range(len(topo.input_chans)):
            m[f"zero_convs.{i}.0"] = f"controlnet_down_blocks.{i}"
        m["middle_block_out.0"] = "controlnet_mid_block"
    return m


def ldm_to_diffusers_key(key: str, prefix_map):
    """Translate one parameter name; longest matching module prefix wins."""
    best = None
    for src in prefix_map:
        if key == src or key.startswith(src + "."):
            if best is None or len(src) > len(best):
                best = src
    if best is None:
        raise KeyError(key)
    rest = key[len(best):].lstrip(".")
    for a, b in _RES.items():                 # ResBlock-internal names
        if rest == a + ".weight" or rest == a + ".bias":
            rest = b + rest[len(a):]
            break
    return prefix_map[best] + ("." + rest if rest else "")


def key_map(cfg: UNetConfig, kind: str, ldm_keys):
    pm = ldm_prefix_map(cfg, kind)
    return {k: ldm_to_diffusers_key(k, pm) for k in ldm_keys}


def diffusers_to_ldm(state_dict, kind: str, cfg: UNetConfig):
    """Re-key a diffusers UNet2DConditionModel / ControlNetModel state dict to ldm/cldm names.
    diffusers stores the transformer `proj_in/proj_out` of SD1.x as 1x1 convs exactly like ldm; SD2.x
    (`use_linear_projection`) as Linear like `use_linear_in_transformer` — shapes carry over."""
    from .unet_spec import param_shapes
    fwd = key_map(cfg, kind, param_shapes(cfg, kind).keys())
    out = {}
    missing = []
    for lk, dk in fwd.items():
        if dk in state_dict:
            out[lk] = state_dict[dk]
        else:
            missing.append(dk)
    if missing:
        raise KeyError(f"{len(missing)} diffusers keys missing, e.g. {missing[:3]}")
    return out


# ---- first-stage decoder (AutoencoderKL) ------------------------------------------------------------
def vae_ldm_to_diffusers_names(vcfg, attn_style="to_q", part="decoder"):
    """{ldm parameter name: diffusers AutoencoderKL parameter name} for post_quant_conv + decoder
    (`part="decoder"`) or encoder + quant_conv (`part="encoder"`).
    diffusers numbers the up blocks from the lowest resolution (up_blocks.0 = ldm up.{L-1}; down blocks keep
    their order), calls the 1x1 shortcut `conv_shortcut`, the output norm `conv_norm_out`, and stores the mid
    attention as Linear layers named to_q/to_k/to_v/to_out.0 (`attn_style="to_q"`, diffusers >= 0.18) or
    query/key/value/proj_attn (`attn_style="query"`, diffusers <= 0.17, the reference's pin)."""
    from .vae_spec import vae_decoder_param_shapes, vae_encoder_param_shapes
    attn = ({"q": "to_q", "k": "to_k", "v": "to_v", "proj_out": "to_out.0", "norm": "group_norm"} if attn_style == "to_q"
            else {"q": "query", "k": "key", "v": "value", "proj_out": "proj_attn", "norm": "group_norm"})
    L_ = vcfg.num_resolutions
    out = {}
    shapes = vae_decoder_param_shapes(vcfg) if part == "decoder" else vae_encoder_param_shapes(vcfg)
    for k in shapes:
        parts = k.split(".")
        side = parts[0]                                   # "decoder" | "encoder" | "(post_)quant_conv"
        if side in ("post_quant_conv", "quant_conv") or parts[1] in ("conv_in", "conv_out"):
            out[k] = k
        elif parts[1] == "norm_out":
            out[k] = f"{side}.conv_norm_out." + parts[-1]
        elif parts[1] == "mid":
            if parts[2].startswith("block_"):
                rest = ".".join(parts[3:]).replace("nin_shortcut", "conv_shortcut")
                out[k] = f"{side}.mid_block.resnets.{int(parts[2][-1]) - 1}.{rest}"
            else:
                out[k] = f"{side}.mid_block.attentions.0.{attn[parts[3]]}.{parts[-1]}"
        elif parts[1] == "up":   # decoder.up.<lvl>.block.<i>.<...> | decoder.up.<lvl>.upsample.conv.<...>
            ub = L_ - 1 - int(parts[2])
            if parts[3] == "block":
                rest = ".".join(parts[5:]).replace("nin_shortcut", "conv_shortcut")
                out[k] = f"decoder.up_blocks.{ub}.resnets.{parts[4]}.{rest}"
            else:
                out[k] = f"decoder.up_blocks.{ub}.upsamplers.0.conv.{parts[-1]}"
        else:                    # encoder.down.<lvl>.block.<i>.<...> | encoder.down.<lvl>.downsample.conv.<...>
            if parts[3] == "block":
                rest = ".".join(parts[5:]).replace("nin_shortcut", "conv_shortcut")
                out[k] = f"encoder.down_blocks.{parts[2]}.resnets.{parts[4]}.{rest}"
            else:
                out[k] = f"encoder.down_blocks.{parts[2]}.downsamplers.0.conv.{parts[-1]}"
    return out


def vae_diffusers_to_ldm(state_dict, vcfg, parts=("decoder",)):
    """Re-key a diffusers AutoencoderKL state dict (either attention spelling) to the ldm names
    editanything_b200.vae consumes, for the requested parts ("decoder", "encoder"); Linear attention
    weights become 1x1 convs; keys of the other part are ignored."""
    out = {}
    for part in parts:
        done = False
        for style in ("to_q", "query"):
            names = vae_ldm_to_diffusers_names(vcfg, style, part)
            if all(v in state_dict for v in names.values()):
                for lk, dk in names.items():
                    t = state_dict[dk]
                    if ".attn_1." in lk and lk.endswith("weight") and ".norm." not in lk and t.dim() == 2:
                        t = t.reshape(t.shape[0], t.shape[1], 1, 1)
                    out[lk] = t
                done = True
                break
        if not done:
            missing = [v for v in vae_ldm_to_diffusers_names(vcfg, "to_q", part).values() if v not in state_dict][:5]
            raise KeyError(f"not a diffusers AutoencoderKL {part} state dict; missing e.g. {missing}")
    return out
