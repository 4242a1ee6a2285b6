"""`StableDiffusionControlNetInpaintPipeline` host logic (call surface, check_inputs errors, latent /
mask / conditioning preparation, loop + blend semantics) against a straight re-enactment of the
reference loop (utils/stable_diffusion_controlnet_inpaint.py:1540-1664) on the CPU oracle.  The
engine's operators are emulated on CPU (tests/cpu_ops.py)."""
import math
from types import SimpleNamespace

import pytest
import torch
import torch.nn.functional as F

from editanything_b200.denoise import DenoiseEngine
from editanything_b200.pipeline import (DDIMScheduler, StableDiffusionControlNetInpaintPipeline,
                                        prepare_controlnet_conditioning_image, prepare_mask_image)
from editanything_b200.unet_spec import TINY, build_topology, make_state_dict
from oracle import unet_oracle as O
from tests import cpu_ops


class FakeVAE:
    """Deterministic stand-in with the diffusers AutoencoderKL surface the pipeline touches."""
    config = SimpleNamespace(scaling_factor=0.18215, latent_channels=4, block_out_channels=(1, 2, 3, 4))

    def encode(self, x):
        z = F.avg_pool2d(x, 8)
        z = torch.cat([z, z.mean(1, keepdim=True)], 1)
        return SimpleNamespace(latent_dist=SimpleNamespace(sample=lambda generator=None: z))

    def decode(self, z):
        return SimpleNamespace(sample=F.interpolate(z[:, :3], scale_factor=8, mode="nearest"))


def _setup(n_cn=2):
    cfg = TINY
    usd = make_state_dict(cfg, "unet", 51)
    csds = [make_state_dict(cfg, "controlnet", 52 + i) for i in range(n_cn)]
    eng = DenoiseEngine(cfg, usd, csds, torch.device("cpu"), backend=cpu_ops)
    pipe = StableDiffusionControlNetInpaintPipeline(eng, vae=FakeVAE())
    g = torch.Generator().manual_seed(0)
    H = W = 64
    image = torch.rand(1, 3, H, W, generator=g) * 2 - 1
    mask = torch.zeros(1, 1, H, W)
    mask[:, :, 16:48, 8:40] = 1.0
    conds = [torch.randint(0, 256, (1, 3, H, W), generator=g).float(), torch.rand(1, 3, H, W, generator=g)][:n_cn]
    pe = torch.randn(1, 9, cfg.context_dim, generator=g)
    ne = torch.randn(1, 9, cfg.context_dim, generator=g)
    return cfg, usd, csds, pipe, image, mask, conds, pe, ne


def _reference_loop(cfg, usd, csds, image, mask, conds, pe, ne, steps, gs, scales, seed, alignment_ratio, n_img=1):
    """The reference __call__ re-enacted with the oracle networks (in_channels == 4 branch)."""
    vae = FakeVAE()
    sch = DDIMScheduler()
    sch.set_timesteps(steps)
    ts = sch.timesteps
    ctx = torch.cat([ne.repeat(n_img, 1, 1), pe.repeat(n_img, 1, 1)])
    hints = [torch.cat([c.repeat_interleave(n_img, 0)] * 2) for c in conds]
    lat = torch.randn((n_img, 4, 8, 8), generator=torch.Generator().manual_seed(seed))
    noise = lat
    init = 0.18215 * vae.encode(image).latent_dist.sample().repeat(n_img, 1, 1, 1)
    m = 1 - F.interpolate((mask >= 0.5).float(), (8, 8), mode="nearest")
    ut, ct = build_topology(cfg), build_topology(cfg, with_decoder=False)
    for i, t in enumerate(ts):
        x_in = torch.cat([lat] * 2)
        with torch.no_grad():
            e = O.apply_model(usd, ut, [(sd, ct) for sd in csds], x_in, torch.full((2 * n_img,), int(t)), ctx, hints,
                              scales)
        e = e[:n_img] + gs * (e[n_img:] - e[:n_img])
        lat = sch.step(e, t, lat).prev_sample
        if alignment_ratio is not None and i < len(ts) * alignment_ratio:
            lat = sch.add_noise(init, noise, ts[i + 1]) * m + lat * (1 - m)
    if alignment_ratio is None or alignment_ratio == 1.0:
        lat = init * m + lat * (1 - m)
    return lat


@pytest.mark.parametrize("alignment_ratio", [None, 0.5])
def test_call_matches_reference_loop(alignment_ratio):
    cfg, usd, csds, pipe, image, mask, conds, pe, ne = _setup()
    out = pipe(image=image, mask_image=mask, controlnet_conditioning_image=conds, height=64, width=64,
               num_inference_steps=4, guidance_scale=9.0, generator=torch.manual_seed(7), prompt_embeds=pe,
               negative_prompt_embeds=ne, output_type="latent", controlnet_conditioning_scale=[0.5, 1.0],
               alignment_ratio=alignment_ratio, num_images_per_prompt=1)
    ref = _reference_loop(cfg, usd, csds, image, mask, conds, pe, ne, 4, 9.0, [0.5, 1.0], 7, alignment_ratio)
    assert out.images.shape == (1, 4, 8, 8)
    assert (out.images - ref).abs().max().item() < 2e-4


def test_generic_scheduler_path_and_decode():
    """A scheduler that is not the built-in DDIM (the demos install UniPC) goes through
    engine.eps + scheduler.step; output_type='np' decodes through the attached VAE."""
    cfg, usd, csds, pipe, image, mask, conds, pe, ne = _setup(n_cn=1)

    class OtherDDIM:                      # same math, different class -> generic path
        def __init__(self):
            self._s = DDIMScheduler()
            self.order, self.init_noise_sigma = 1, 1.0

        def set_timesteps(self, n, device=None):
            self._s.set_timesteps(n)
            self.timesteps = self._s.timesteps

        def scale_model_input(self, x, t):
            return x

        def step(self, e, t, x, **kw):
            return self._s.step(e, t, x)

        def add_noise(self, a, b, t):
            return self._s.add_noise(a, b, t)

    kw = dict(image=image, mask_image=mask, controlnet_conditioning_image=conds, height=64, width=64,
              num_inference_steps=4, guidance_scale=5.0, prompt_embeds=pe, negative_prompt_embeds=ne,
              controlnet_conditioning_scale=[0.8], alignment_ratio=None)
    a = pipe(generator=torch.manual_seed(3), output_type="latent", **kw).images
    pipe.scheduler = OtherDDIM()
    b = pipe(generator=torch.manual_seed(3), output_type="latent", **kw).images
    assert (a - b).abs().max().item() < 1e-4
    imgs = pipe(generator=torch.manual_seed(3), output_type="np", return_dict=False, **kw)[0]
    assert imgs.shape == (1, 64, 64, 3) and imgs.min() >= 0 and imgs.max() <= 1


def test_check_inputs_errors_match_reference_types():
    cfg, usd, csds, pipe, image, mask, conds, pe, ne = _setup()
    ok = dict(image=image, mask_image=mask, controlnet_conditioning_image=conds, height=64, width=64,
              prompt_embeds=pe, negative_prompt_embeds=ne, controlnet_conditioning_scale=[0.5, 1.0],
              num_inference_steps=1, output_type="latent")
    with pytest.raises(ValueError):
        pipe(**{**ok, "height": 60})
    with pytest.raises(ValueError):
        pipe(**{**ok, "callback_steps": 0})
    with pytest.raises(ValueError):
        pipe(**{**ok, "prompt": "a cat"})                           # both prompt and prompt_embeds
    with pytest.raises(ValueError):
        pipe(**{**ok, "prompt_embeds": None})
    with pytest.raises(ValueError):
        pipe(**{**ok, "negative_prompt_embeds": ne[:, :5]})
    with pytest.raises(TypeError):
        pipe(**{**ok, "controlnet_conditioning_image": conds[0]})   # multi-controlnet needs a list
    with pytest.raises(ValueError):
        pipe(**{**ok, "controlnet_conditioning_image": conds[:1]})
    with pytest.raises(ValueError):
        pipe(**{**ok, "controlnet_conditioning_scale": [1.0]})
    with pytest.raises(ValueError):
        pipe(**{**ok, "image": image * 3})
    with pytest.raises(ValueError):
        pipe(**{**ok, "mask_image": mask[:, :, :32]})
    with pytest.raises(IndexError):
        pipe(**{**ok, "alignment_ratio": 1.0})                      # reference quirk: timesteps[i + 1]
    with pytest.raises(NotImplementedError):
        pipe(**{**ok, "ref_image": image})


def test_preparation_helpers():
    m = prepare_mask_image(torch.tensor([[0.2, 0.7], [0.5, 0.49]]))
    assert m.shape == (1, 1, 2, 2) and m.flatten().tolist() == [0.0, 1.0, 1.0, 0.0]
    c = prepare_controlnet_conditioning_image(torch.full((1, 3, 8, 8), 200.0), 8, 8, 3, 3, torch.float32, True)
    assert c.shape == (6, 3, 8, 8) and float(c.max()) == 200.0     # un-normalised, repeated, CFG-doubled
    s = DDIMScheduler()
    s.set_timesteps(30)
    assert len(s.timesteps) == 31 and int(s.timesteps[0]) == 991    # reference quirk: S=30 -> 31 steps


def test_decode_through_the_vae_engine():
    """output_type='np' with editanything_b200.vae.VaeDecoderEngine attached as `pipe.vae` == the oracle's
    decode_latents (utils/...inpaint.py:718-724) of the latents the same call returns with output_type='latent'."""
    from editanything_b200.vae import VAE_TINY, VaeDecoderEngine, make_vae_state_dict
    from oracle import vae_oracle as V
    cfg, usd, csds, pipe, image, mask, conds, pe, ne = _setup(n_cn=1)
    vsd = make_vae_state_dict(VAE_TINY, 77)
    kw = dict(image=image, mask_image=mask, controlnet_conditioning_image=conds, height=64, width=64,
              num_inference_steps=4, guidance_scale=7.0, prompt_embeds=pe, negative_prompt_embeds=ne,
              controlnet_conditioning_scale=1.0, num_images_per_prompt=1, latents=torch.randn(1, 4, 8, 8, generator=torch.manual_seed(3)))
    lat = pipe(output_type="latent", **kw).images
    pipe.vae = _EncDec(VaeDecoderEngine(VAE_TINY, vsd, torch.device("cpu"), backend=cpu_ops))
    out = pipe(output_type="np", **kw).images
    with torch.no_grad():
        ref = V.decode_latents(lat.float(), vsd, VAE_TINY).permute(0, 2, 3, 1).numpy()
    # the VAE_TINY decoder has one upsampling level: 8x8 latents -> 16x16 image
    assert out.shape == ref.shape == (1, 16, 16, 3)
    assert abs(out - ref).max() < 1e-3


class _EncDec:
    """The decoder engine plus FakeVAE's encode (the encode side stays PyTorch)."""

    def __init__(self, dec):
        self._dec, self.config = dec, dec.config
        self.decode_latents, self.decode, self.encode = dec.decode_latents, dec.decode, FakeVAE().encode


def test_full_call_with_the_vae_engine_on_both_sides():
    """`pipe.vae = VaeEngine`: the masked source image is encoded by the engine (prepare_masked_image_latents,
    utils/...inpaint.py:1056-1105), the final blend uses those latents and the result is decoded by the engine
    (decode_latents :718-724).  Re-enacted with the oracle VAE + oracle networks, same generator order."""
    from editanything_b200.vae import VaeEngine, make_vae_state_dict
    from editanything_b200.vae_spec import VaeConfig
    from oracle import vae_oracle as V
    vcfg = VaeConfig(ch=64, ch_mult=(1, 1, 1, 1), num_res_blocks=1)          # f = 8 like kl-f8, test-sized
    vsd = dict(make_vae_state_dict(vcfg, 61, part="encoder"))
    vsd.update(make_vae_state_dict(vcfg, 62))
    cfg, usd, csds, pipe, image, mask, conds, pe, ne = _setup(n_cn=1)
    pipe = StableDiffusionControlNetInpaintPipeline(pipe.engine, vae=VaeEngine(vcfg, vsd, torch.device("cpu"), backend=cpu_ops))
    assert pipe.vae_scale_factor == 8
    steps, gs = 4, 7.0
    out = pipe(image=image, mask_image=mask, controlnet_conditioning_image=conds, height=64, width=64,
               num_inference_steps=steps, guidance_scale=gs, generator=torch.manual_seed(11), prompt_embeds=pe,
               negative_prompt_embeds=ne, output_type="np", controlnet_conditioning_scale=1.0,
               num_images_per_prompt=1).images
    # --- re-enactment -------------------------------------------------------------------------------
    gen = torch.manual_seed(11)
    lat = torch.randn((1, 4, 8, 8), generator=gen)                                       # prepare_latents
    with torch.no_grad():                      # in_channels == 4 branch: the source image itself is encoded (:1469-1476)
        mom = V.encode_moments(image, vsd, vcfg)
    mean, logvar = mom.chunk(2, 1)
    init = vcfg.scaling_factor * (mean + torch.exp(0.5 * logvar.clamp(-30, 20)) * torch.randn(mean.shape, generator=gen))
    m = 1 - F.interpolate((mask >= 0.5).float(), (8, 8), mode="nearest")
    sch = DDIMScheduler()
    sch.set_timesteps(steps)
    ctx = torch.cat([ne, pe])
    hints = [torch.cat([c] * 2) for c in conds]
    ut, ct = build_topology(cfg), build_topology(cfg, with_decoder=False)
    for t in sch.timesteps:
        with torch.no_grad():
            e = O.apply_model(usd, ut, [(sd, ct) for sd in csds], torch.cat([lat] * 2), torch.full((2,), int(t)), ctx, hints, [1.0])
        lat = sch.step(e[:1] + gs * (e[1:] - e[:1]), t, lat).prev_sample
    lat = init * m + lat * (1 - m)
    with torch.no_grad():
        ref = V.decode_latents(lat, vsd, vcfg).permute(0, 2, 3, 1).numpy()
    assert out.shape == ref.shape == (1, 64, 64, 3)
    assert abs(out - ref).max() < 2e-3



def test_callback_sees_unblended_latents():
    """The reference calls `callback(i, t, latents)` before the inpaint blend of step i
    (utils/...inpaint.py:1640-1656); with a callback the blend therefore runs after it, outside the fused step."""
    cfg, usd, csds, pipe, image, mask, conds, pe, ne = _setup()
    seen = []
    kw = dict(image=image, mask_image=mask, controlnet_conditioning_image=conds, height=64, width=64,
              num_inference_steps=4, guidance_scale=9.0, prompt_embeds=pe, negative_prompt_embeds=ne,
              output_type="latent", controlnet_conditioning_scale=[0.5, 1.0], alignment_ratio=0.5,
              num_images_per_prompt=1)
    a = pipe(generator=torch.manual_seed(7), callback=lambda i, t, x: seen.append(x.clone()), **kw).images
    b = pipe(generator=torch.manual_seed(7), **kw).images
    ref = _reference_loop(cfg, usd, csds, image, mask, conds, pe, ne, 4, 9.0, [0.5, 1.0], 7, 0.5)
    assert len(seen) == 4
    assert (a - ref).abs().max().item() < 2e-4 and (b - ref).abs().max().item() < 2e-4
    # step 0's callback latents differ from the blended ones inside the kept region
    assert (seen[0] - a).abs().max().item() > 1e-3


@pytest.mark.parametrize("alignment_ratio", [None, 0.5])
def test_unipc_fused_update_matches_scheduler_object_loop(alignment_ratio):
    """`pipe.scheduler = UniPCMultistepScheduler.from_config(pipe.scheduler.config)` (editany_lora.py:383): the fused
    device-side multistep update (coefficient rows + history buffers) == the reference loop re-enacted with the
    scheduler OBJECT's step() on the oracle networks."""
    from editanything_b200.schedulers import UniPCMultistepScheduler
    cfg, usd, csds, pipe, image, mask, conds, pe, ne = _setup()
    pipe.scheduler = UniPCMultistepScheduler.from_config(pipe.scheduler.config)
    steps, gs = 6, 7.0
    out = pipe(image=image, mask_image=mask, controlnet_conditioning_image=conds, height=64, width=64,
               num_inference_steps=steps, guidance_scale=gs, generator=torch.manual_seed(7), prompt_embeds=pe,
               negative_prompt_embeds=ne, output_type="latent", controlnet_conditioning_scale=[0.5, 1.0],
               alignment_ratio=alignment_ratio, num_images_per_prompt=1).images
    # re-enactment with the scheduler object
    vae = FakeVAE()
    sch = UniPCMultistepScheduler.from_config(DDIMScheduler().config)
    sch.set_timesteps(steps)
    ts = sch.timesteps
    ctx = torch.cat([ne, pe])
    hints = [torch.cat([c] * 2) for c in conds]
    lat = torch.randn((1, 4, 8, 8), generator=torch.Generator().manual_seed(7))
    noise = lat
    init = 0.18215 * vae.encode(image).latent_dist.sample()
    m = 1 - F.interpolate((mask >= 0.5).float(), (8, 8), mode="nearest")
    ut, ct = build_topology(cfg), build_topology(cfg, with_decoder=False)
    for i, t in enumerate(ts):
        with torch.no_grad():
            e = O.apply_model(usd, ut, [(sd, ct) for sd in csds], torch.cat([lat] * 2), torch.full((2,), int(t)), ctx, hints,
                              [0.5, 1.0])
        lat = sch.step(e[:1] + gs * (e[1:] - e[:1]), t, lat).prev_sample
        if alignment_ratio is not None and i < len(ts) * alignment_ratio:
            lat = sch.add_noise(init, noise, ts[i + 1]) * m + lat * (1 - m)
    if alignment_ratio is None:
        lat = init * m + lat * (1 - m)
    assert (out - lat).abs().max().item() < 5e-4, (out - lat).abs().max().item()
