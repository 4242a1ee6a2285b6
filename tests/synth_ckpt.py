"""Synthetic checkpoints in the on-disk formats the reference downloads (test infrastructure): diffusers-layout
pipeline / ControlNet directories and an upstream-format SAM .pth, all test-sized with seeded random weights."""
import json
import os
import string

import torch

from editanything_b200 import weights as W
from editanything_b200.unet_spec import UNetConfig, make_state_dict
from editanything_b200.vae import make_vae_state_dict
from editanything_b200.vae_spec import VaeConfig


def _unet_json(cfg: UNetConfig, kind):
    mc = cfg.model_channels
    boc = [mc * m for m in cfg.channel_mult]
    n = len(boc)
    ahd = cfg.num_heads if cfg.num_head_channels == -1 else [b // cfg.num_head_channels for b in boc]
    c = {"in_channels": cfg.in_channels, "block_out_channels": boc, "layers_per_block": cfg.num_res_blocks,
         "cross_attention_dim": cfg.context_dim, "attention_head_dim": ahd,
         "use_linear_projection": cfg.use_linear_in_transformer,
         "down_block_types": ["CrossAttnDownBlock2D"] * (n - 1) + ["DownBlock2D"]}
    if kind == "unet":
        c["out_channels"] = cfg.out_channels
    return c


def _save(sd, path):
    from safetensors.torch import save_file
    save_file({k: v.contiguous() for k, v in sd.items()}, path)


def write_unet_like(dirpath, cfg, kind, seed):
    os.makedirs(dirpath, exist_ok=True)
    sd = make_state_dict(cfg, kind, seed)
    km = W.key_map(cfg, kind, sd.keys())
    _save({km[k]: v for k, v in sd.items()}, os.path.join(dirpath, "diffusion_pytorch_model.safetensors"))
    json.dump(_unet_json(cfg, kind), open(os.path.join(dirpath, "config.json"), "w"))
    return sd


def write_vae(dirpath, vcfg: VaeConfig, seed):
    os.makedirs(dirpath, exist_ok=True)
    sd = dict(make_vae_state_dict(vcfg, seed, part="encoder"))
    sd.update(make_vae_state_dict(vcfg, seed + 1))
    out = {}
    for part in ("decoder", "encoder"):
        for lk, dk in W.vae_ldm_to_diffusers_names(vcfg, "to_q", part).items():
            t = sd[lk]
            if ".attn_1." in lk and lk.endswith("weight") and ".norm." not in lk:
                t = t.reshape(t.shape[0], t.shape[1])            # diffusers stores the mid attention as Linear
            out[dk] = t
    _save(out, os.path.join(dirpath, "diffusion_pytorch_model.safetensors"))
    json.dump({"block_out_channels": list(vcfg.block_out_channels), "layers_per_block": vcfg.num_res_blocks,
               "latent_channels": vcfg.z_channels, "scaling_factor": vcfg.scaling_factor, "out_channels": vcfg.out_ch},
              open(os.path.join(dirpath, "config.json"), "w"))
    return sd


def write_text_encoder_and_tokenizer(root, hidden):
    from transformers import CLIPTextConfig, CLIPTextModel, CLIPTokenizer
    chars = list(string.ascii_lowercase + string.digits + ",.!?-")
    vocab = {}
    for c in chars:
        vocab[c] = len(vocab)
    for c in chars:
        vocab[c + "</w>"] = len(vocab)
    vocab["<|startoftext|>"] = len(vocab)
    vocab["<|endoftext|>"] = len(vocab)
    tdir = os.path.join(root, "tokenizer")
    os.makedirs(tdir, exist_ok=True)
    json.dump(vocab, open(os.path.join(tdir, "vocab.json"), "w"))
    open(os.path.join(tdir, "merges.txt"), "w").write("#version: 0.2\n")
    CLIPTokenizer(os.path.join(tdir, "vocab.json"), os.path.join(tdir, "merges.txt"), model_max_length=16).save_pretrained(tdir)
    torch.manual_seed(0)
    cfg = CLIPTextConfig(vocab_size=len(vocab), hidden_size=hidden, intermediate_size=2 * hidden, num_hidden_layers=2,
                         num_attention_heads=4, max_position_embeddings=16, bos_token_id=vocab["<|startoftext|>"],
                         eos_token_id=vocab["<|endoftext|>"], pad_token_id=vocab["<|endoftext|>"])
    CLIPTextModel(cfg).save_pretrained(os.path.join(root, "text_encoder"))


def write_pipeline(root, cfg, vcfg, seed=51):
    """A diffusers-layout base model directory; returns (unet ldm state dict, vae ldm state dict)."""
    os.makedirs(root, exist_ok=True)
    usd = write_unet_like(os.path.join(root, "unet"), cfg, "unet", seed)
    vsd = write_vae(os.path.join(root, "vae"), vcfg, seed + 10)
    write_text_encoder_and_tokenizer(root, cfg.context_dim)
    os.makedirs(os.path.join(root, "scheduler"), exist_ok=True)
    json.dump({"_class_name": "PNDMScheduler", "beta_start": 0.00085, "beta_end": 0.012, "beta_schedule": "scaled_linear",
               "num_train_timesteps": 1000}, open(os.path.join(root, "scheduler", "scheduler_config.json"), "w"))
    json.dump({"_class_name": "StableDiffusionPipeline"}, open(os.path.join(root, "model_index.json"), "w"))
    return usd, vsd


def write_sam(path, cfg, seed=5):
    from editanything_b200.sam_spec import make_sam_state_dict
    from editanything_b200.segment_anything import MaskDecoder, PromptEncoder, TwoWayTransformer
    torch.manual_seed(seed)
    c = cfg.out_chans
    pe = PromptEncoder(embed_dim=c, image_embedding_size=(cfg.grid, cfg.grid), input_image_size=(cfg.img_size, cfg.img_size))
    md = MaskDecoder(transformer_dim=c, transformer=TwoWayTransformer(2, c, 8, 2048))
    sd = {"image_encoder." + k: v for k, v in make_sam_state_dict(cfg, seed).items()}
    sd.update({"prompt_encoder." + k: v for k, v in pe.state_dict().items()})
    sd.update({"mask_decoder." + k: v for k, v in md.state_dict().items()})
    os.makedirs(os.path.dirname(path), exist_ok=True)
    torch.save(sd, path)
    return sd
