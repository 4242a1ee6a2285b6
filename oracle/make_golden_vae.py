"""Generate tests/golden/vae_*.pt by RUNNING the reference's own Decoder
(ldm/modules/diffusionmodules/model.py:546-652, imported from /root/reference; build container only).

    python -m oracle.make_golden_vae [--full]

Weights are the deterministic synthetic ones of editanything_b200.vae_spec.make_vae_state_dict; latents
are seeded N(0,1) scaled like the sampler's output.  The stored tensor is decode_latents' image
((x / 2 + 0.5).clamp(0, 1), utils/stable_diffusion_controlnet_inpaint.py:718-722)."""
import os
import sys

import torch

from editanything_b200.vae_spec import SD_VAE, VAE_TINY, make_vae_state_dict
from oracle import vae_oracle as V

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
CASES = {"vae_tiny": (VAE_TINY, 2, 16, 401, 7)}          # name -> (cfg, batch, latent side, weight seed, latent seed)
FULL_CASES = {"vae_sd": (SD_VAE, 1, 64, 402, 8)}


def make_latents(cfg, B, side, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, cfg.z_channels, side, side, generator=g) * cfg.scaling_factor * 4.0


def run_case(name, spec):
    cfg, B, side, wseed, lseed = spec
    sd = make_vae_state_dict(cfg, wseed)
    lat = make_latents(cfg, B, side, lseed)
    ref = V.reference_decoder(cfg, sd)
    raw = ref(lat / cfg.scaling_factor)
    img = (raw / 2 + 0.5).clamp(0, 1)
    rec = {"meta": dict(name=name, B=B, side=side, weight_seed=wseed, latent_seed=lseed,
                        generator="reference ldm Decoder + post_quant_conv", raw_abs_max=float(raw.abs().max()),
                        raw_std=float(raw.std())),
           "image": img.to(torch.float16) if side >= 64 else img.clone(),
           "raw_center": raw[:, :, raw.shape[2] // 2 - 8:raw.shape[2] // 2 + 8, raw.shape[3] // 2 - 8:raw.shape[3] // 2 + 8].clone()}
    torch.save(rec, os.path.join(GOLD, name + ".pt"))
    print(name, tuple(img.shape), "raw abs max", float(raw.abs().max()), "raw std", float(raw.std()),
          "fraction clamped", float(((raw / 2 + 0.5) < 0).float().mean() + ((raw / 2 + 0.5) > 1).float().mean()))


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    for n, s in (FULL_CASES if "--full" in sys.argv else CASES).items():
        run_case(n, s)
