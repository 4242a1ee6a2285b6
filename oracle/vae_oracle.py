"""CPU ORACLE for the first-stage (AutoencoderKL) decoder (test infrastructure — NOT a product path).

Plain-PyTorch fp32 functional restatement of the reference's decode path
    decode_latents       utils/stable_diffusion_controlnet_inpaint.py:718-724
    AutoencoderKL.decode ldm/models/autoencoder.py:88-91
    Decoder.forward      ldm/modules/diffusionmodules/model.py:623-652
operating directly on an ldm-named state dict.  Each function cites the reference lines it follows.

Parity status: PINNED — tests/test_vae_cpu.py::test_oracle_matches_reference_decoder instantiates the
reference's own `Decoder` + `post_quant_conv` (imported from /root/reference in the build container) with
the same weights and requires agreement to fp32 round-off; tests/golden/vae_*.pt hold outputs generated
by that reference module (oracle/make_golden_vae.py), which is what the GPU-box tests compare with.
"""
import torch
import torch.nn.functional as F

from editanything_b200.vae_spec import VaeConfig, decoder_blocks


def normalize(x, sd, p, cfg):
    """Normalize = GroupNorm(32, C, eps=1e-6, affine) (model.py:46-47)."""
    return F.group_norm(x, cfg.num_groups, sd[p + ".weight"], sd[p + ".bias"], eps=cfg.eps)


def nonlinearity(x):
    """swish (model.py:41-43)."""
    return x * torch.sigmoid(x)


def resnet_block(x, sd, p, cfg):
    """ResnetBlock.forward with temb = None (model.py:129-148)."""
    h = F.conv2d(nonlinearity(normalize(x, sd, p + ".norm1", cfg)), sd[p + ".conv1.weight"], sd[p + ".conv1.bias"], padding=1)
    h = F.conv2d(nonlinearity(normalize(h, sd, p + ".norm2", cfg)), sd[p + ".conv2.weight"], sd[p + ".conv2.bias"], padding=1)
    if p + ".nin_shortcut.weight" in sd:
        x = F.conv2d(x, sd[p + ".nin_shortcut.weight"], sd[p + ".nin_shortcut.bias"])
    return x + h


def attn_block(x, sd, p, cfg):
    """AttnBlock.forward (model.py:181-210): single head over all C channels."""
    h_ = normalize(x, sd, p + ".norm", cfg)
    q = F.conv2d(h_, sd[p + ".q.weight"], sd[p + ".q.bias"])
    k = F.conv2d(h_, sd[p + ".k.weight"], sd[p + ".k.bias"])
    v = F.conv2d(h_, sd[p + ".v.weight"], sd[p + ".v.bias"])
    b, c, h, w = q.shape
    q = q.reshape(b, c, h * w).permute(0, 2, 1)
    k = k.reshape(b, c, h * w)
    w_ = torch.bmm(q, k) * (int(c) ** (-0.5))
    w_ = F.softmax(w_, dim=2)
    v = v.reshape(b, c, h * w)
    h_ = torch.bmm(v, w_.permute(0, 2, 1)).reshape(b, c, h, w)
    h_ = F.conv2d(h_, sd[p + ".proj_out.weight"], sd[p + ".proj_out.bias"])
    return x + h_


def upsample(x, sd, p):
    """Upsample.forward with_conv (model.py:60-64)."""
    x = F.interpolate(x, scale_factor=2.0, mode="nearest")
    return F.conv2d(x, sd[p + ".conv.weight"], sd[p + ".conv.bias"], padding=1)


def decode(z, sd, cfg: VaeConfig):
    """AutoencoderKL.decode (autoencoder.py:88-91) = Decoder.forward(post_quant_conv(z))."""
    z = F.conv2d(z, sd["post_quant_conv.weight"], sd["post_quant_conv.bias"])
    h = F.conv2d(z, sd["decoder.conv_in.weight"], sd["decoder.conv_in.bias"], padding=1)
    blocks, _ = decoder_blocks(cfg)
    for kind, p, cin, cout in blocks:
        if kind == "res":
            h = resnet_block(h, sd, p, cfg)
        elif kind == "attn":
            h = attn_block(h, sd, p, cfg)
        else:
            h = upsample(h, sd, p)
    h = nonlinearity(normalize(h, sd, "decoder.norm_out", cfg))
    return F.conv2d(h, sd["decoder.conv_out.weight"], sd["decoder.conv_out.bias"], padding=1)


def decode_latents(latents, sd, cfg: VaeConfig):
    """decode_latents (utils/stable_diffusion_controlnet_inpaint.py:718-722) without the numpy plumbing."""
    image = decode(latents / cfg.scaling_factor, sd, cfg)
    return (image / 2 + 0.5).clamp(0, 1)


def reference_decoder(cfg: VaeConfig, sd):
    """The reference's OWN modules (build container only): ldm Decoder + post_quant_conv loaded with `sd`."""
    from oracle import ref_shim
    ref_shim.load()
    from ldm.modules.diffusionmodules.model import Decoder
    dec = Decoder(ch=cfg.ch, out_ch=cfg.out_ch, ch_mult=cfg.ch_mult, num_res_blocks=cfg.num_res_blocks,
                  attn_resolutions=[], dropout=0.0, in_channels=3, resolution=256, z_channels=cfg.z_channels,
                  attn_type="vanilla").eval()
    dec.load_state_dict({k[len("decoder."):]: v for k, v in sd.items() if k.startswith("decoder.")}, strict=True)
    pqc = torch.nn.Conv2d(cfg.embed_dim, cfg.z_channels, 1)
    pqc.load_state_dict({"weight": sd["post_quant_conv.weight"], "bias": sd["post_quant_conv.bias"]})

    def run(z):
        with torch.no_grad():
            return dec(pqc(z))
    return run


# ---- encode side -----------------------------------------------------------------------------------
def downsample(x, sd, p):
    """Downsample.forward with_conv (model.py:79-86): asymmetric pad (0,1,0,1), stride-2 pad-0 conv."""
    x = F.pad(x, (0, 1, 0, 1), mode="constant", value=0)
    return F.conv2d(x, sd[p + ".conv.weight"], sd[p + ".conv.bias"], stride=2, padding=0)


def encode_moments(x, sd, cfg: VaeConfig):
    """AutoencoderKL.encode up to the distribution (autoencoder.py:82-86): quant_conv(Encoder(x))
    (Encoder.forward model.py:519-543)."""
    from editanything_b200.vae_spec import encoder_blocks
    h = F.conv2d(x, sd["encoder.conv_in.weight"], sd["encoder.conv_in.bias"], padding=1)
    blocks, _ = encoder_blocks(cfg)
    for kind, p, cin, cout in blocks:
        if kind == "res":
            h = resnet_block(h, sd, p, cfg)
        elif kind == "attn":
            h = attn_block(h, sd, p, cfg)
        else:
            h = downsample(h, sd, p)
    h = nonlinearity(normalize(h, sd, "encoder.norm_out", cfg))
    h = F.conv2d(h, sd["encoder.conv_out.weight"], sd["encoder.conv_out.bias"], padding=1)
    return F.conv2d(h, sd["quant_conv.weight"], sd["quant_conv.bias"])


def reference_encoder(cfg: VaeConfig, sd):
    """The reference's OWN modules (build container only): ldm Encoder + quant_conv loaded with `sd`."""
    from oracle import ref_shim
    ref_shim.load()
    from ldm.modules.diffusionmodules.model import Encoder
    enc = Encoder(ch=cfg.ch, out_ch=cfg.out_ch, ch_mult=cfg.ch_mult, num_res_blocks=cfg.num_res_blocks,
                  attn_resolutions=[], dropout=0.0, in_channels=3, resolution=256, z_channels=cfg.z_channels,
                  double_z=True, attn_type="vanilla").eval()
    enc.load_state_dict({k[len("encoder."):]: v for k, v in sd.items() if k.startswith("encoder.")}, strict=True)
    qc = torch.nn.Conv2d(2 * cfg.z_channels, 2 * cfg.embed_dim, 1)
    qc.load_state_dict({"weight": sd["quant_conv.weight"], "bias": sd["quant_conv.bias"]})

    def run(x):
        with torch.no_grad():
            return qc(enc(x))
    return run
