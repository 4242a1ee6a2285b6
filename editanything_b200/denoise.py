"""The per-step engine: ControlNet(s) -> UNet -> CFG -> DDIM on the B200 kernels.

This is the seam SURVEY.md §8b calls B3: what
utils/stable_diffusion_controlnet_inpaint.py:1607-1636 does with `self.controlnet(...)`,
`self.unet(...)`, the guidance combine and `scheduler.step(...)`, as ONE stream-ordered sequence
of libea_b200 launches (optionally replayed as a CUDA graph).
"""
import math

import numpy as np
import torch

from . import ops as _cuda_ops
from .nets import PackedNet, UNetRunner
from .unet_spec import UNetConfig


def ddim_schedule(num_steps, linear_start=0.00085, linear_end=0.012, T=1000):
    """DDIM 'uniform' schedule of the reference sampler: betas (util.py:21-25, cldm_v21.yaml:4-5),
    alphas_cumprod (ddpm.py:145-147), timesteps range(0,T,T//S)+1 (util.py:46-60), a_prev
    (util.py:63-74).  Returned in SAMPLING order (descending t).  S=30 yields 31 steps like the
    reference (1000//30 = 33)."""
    betas = np.linspace(linear_start ** 0.5, linear_end ** 0.5, T, dtype=np.float64) ** 2
    ac = np.cumprod(1.0 - betas, axis=0)
    ts = np.asarray(list(range(0, T, T // num_steps))) + 1
    a = ac[ts]
    a_prev = np.asarray([ac[0]] + ac[ts[:-1]].tolist())
    return ts[::-1].copy(), a[::-1].copy(), a_prev[::-1].copy()


class DenoiseEngine:
    def __init__(self, cfg: UNetConfig, unet_sd, controlnet_sds, device, backend=None):
        self.cfg, self.dev = cfg, device
        self.ops = backend or _cuda_ops
        self.unet = PackedNet(cfg, "unet", unet_sd, device, backend)
        self.cns = [PackedNet(cfg, "controlnet", sd, device, backend) for sd in controlnet_sds]
        self.runner = UNetRunner(self.unet, self.cns, device)
        self.hdt = self.unet.hdt
        self._graph = None
        self.launches_per_step = None

    def weight_bytes(self):
        return self.unet.weight_bytes() + sum(c.weight_bytes() for c in self.cns)

    # -- per request -------------------------------------------------------------------------
    def _keep(self, name, new):
        """Store `new` under self.<name>, reusing the existing buffer (same address => a captured
        CUDA graph stays valid) when shape/dtype match; otherwise invalidate the graph."""
        old = getattr(self, name, None)
        if torch.is_tensor(old) and torch.is_tensor(new) and old.shape == new.shape and old.dtype == new.dtype:
            old.copy_(new)
            return old
        self._graph = None
        setattr(self, name, new)
        return new

    def prepare(self, ctx, hints, scales):
        """ctx: [B, L, D] prompt embeddings ([negative; positive] stacked for CFG,
        utils/stable_diffusion_controlnet_inpaint.py:1339-1347); hints: list of NCHW conditioning
        images [B, 3, 8h, 8w] (un-normalised, editany_lora.py:771-778,814-828); scales: list."""
        ctx = ctx.to(self.dev)
        self.B = ctx.shape[0]
        nets = [self.unet] + self.cns
        caches = [n.precompute_context(ctx) for n in nets]
        old = getattr(self, "ctx_cache", None)
        if old is not None and len(old) == len(caches) and all(
                o["L"] == c["L"] and o["B"] == c["B"] for o, c in zip(old, caches)):
            for o, c in zip(old, caches):
                for k in o["kv"]:
                    o["kv"][k].copy_(c["kv"][k])
        else:
            self.ctx_cache = caches
            self._graph = None
        new_hints = [c.precompute_hint(h.to(self.dev)) for c, h in zip(self.cns, hints)]
        old_h = getattr(self, "hints", None)
        if old_h is not None and len(old_h) == len(new_hints) and all(o.shape == n.shape for o, n in zip(old_h, new_hints)):
            for o, n in zip(old_h, new_hints):
                o.copy_(n)
        else:
            self.hints = new_hints
            self._graph = None
        new_scales = [float(s) for s in scales]
        if getattr(self, "scales", None) != new_scales:
            self._graph = None        # scales are baked into the zero-conv launches
        self.scales = new_scales
        if not hasattr(self, "t_dev") or self.t_dev.shape[0] != self.B:
            self._emb_cache = {}
            nets = [self.unet] + self.cns
            self.emb_bufs = [torch.zeros(self.B, n.emb_total, device=self.dev, dtype=torch.float32) for n in nets]
            self.t_dev = torch.zeros(self.B, device=self.dev, dtype=torch.float32)
            self.coef_dev = torch.zeros(4, device=self.dev, dtype=torch.float32)
            self.gn_ws = torch.zeros(self.B * (32 * 2 + 2), device=self.dev, dtype=torch.float32)
            self._graph = None

    def _fill_emb(self, t):
        """Time-embedding rows for timestep t into the fixed buffers the (captured) step reads.
        Computed once per distinct t (util.py:154-174 + time_embed + every ResBlock emb_layers,
        openaimodel.py:526-531,204-210) and cached: the DDIM table repeats for every image."""
        key = float(t)
        rows = self._emb_cache.get(key)
        if rows is None:
            self.t_dev.fill_(key)
            rows = [e.clone() for e in self.runner.compute_embs(self.t_dev, self.B)]
            if len(self._emb_cache) < 1100:
                self._emb_cache[key] = rows
        for buf, r in zip(self.emb_bufs, rows):
            buf.copy_(r)

    # -- parity API: the network output itself ------------------------------------------------
    def eps(self, x_nchw, t):
        """eps = unet(x, t, ctx, control=sum_k scale_k * controlnet_k(x, hint_k, t, ctx)) as fp32
        NCHW — the quantity the reference calls `noise_pred` before guidance."""
        B, C, H, W_ = x_nchw.shape
        xh = x_nchw.to(self.dev).permute(0, 2, 3, 1).contiguous().to(self.hdt)
        self._fill_emb(t)
        xn = self.runner.eps_features(xh, self.t_dev, self.ctx_cache, self.hints, self.scales, self.gn_ws,
                                      embs=self.emb_bufs)
        eps = torch.empty(B, H, W_, 4, device=self.dev, dtype=torch.float32)
        self.ops.out_cfg_ddim(xn, self.unet.w["out.w"], self.unet.w["out.cb"], eps_out=eps, Nimg=B // 2, H=H,
                              W=W_, C_=self.cfg.model_channels)
        return eps.permute(0, 3, 1, 2).contiguous()

    # -- production API: one fused denoising step ----------------------------------------------
    def _step_body(self):
        H, W_ = self.lat.shape[1], self.lat.shape[2]
        xn = self.runner.eps_features(self.x_half, self.t_dev, self.ctx_cache, self.hints, self.scales, self.gn_ws,
                                      embs=self.emb_bufs)
        self.ops.out_cfg_ddim(xn, self.unet.w["out.w"], self.unet.w["out.cb"], latents=self.lat,
                              coef=self.coef_dev, guidance=self.guidance, known=self.known, mask=self.mask,
                              lat_half_out=self.x_half, Nimg=self.B // 2, H=H, W=W_, C_=self.cfg.model_channels)

    def begin(self, latents_nchw, guidance, known_nchw=None, mask_n1hw=None, use_graph=True):
        """latents: fp32 [N, 4, h, w] initial noise (N = B/2 images).  known/mask: optional inpaint
        blend tensors (mask == 1 keeps `known`, utils/...inpaint.py:1484-1489,1647-1664)."""
        lat = latents_nchw.to(self.dev, torch.float32).permute(0, 2, 3, 1).contiguous()
        self._keep("lat", lat)
        self._keep("x_half", torch.cat([lat, lat]).to(self.hdt).contiguous())
        if getattr(self, "guidance", None) != float(guidance):
            self._graph = None
        self.guidance = float(guidance)
        if known_nchw is None:
            if getattr(self, "known", None) is not None:
                self._graph = None
            self.known = self.mask = None
        else:
            self._keep("known", known_nchw.to(self.dev, torch.float32).permute(0, 2, 3, 1).contiguous())
            self._keep("mask", mask_n1hw.to(self.dev, torch.float32).reshape(lat.shape[0], lat.shape[1],
                                                                            lat.shape[2]).contiguous())
        use = use_graph and self.ops is _cuda_ops
        if use != getattr(self, "_use_graph", None):
            self._graph = None
        self._use_graph = use

    def set_known(self, known_nchw):
        self.known.copy_(known_nchw.to(self.dev, torch.float32).permute(0, 2, 3, 1))

    def step(self, t, a_t, a_prev):
        """One DDIM (eta=0) step at timestep t (cldm/ddim_hacked.py:181-231)."""
        self._fill_emb(t)
        self.coef_dev.copy_(torch.tensor([math.sqrt(a_t), math.sqrt(1.0 - a_t), math.sqrt(a_prev),
                                          math.sqrt(1.0 - a_prev)], dtype=torch.float32), non_blocking=True)
        if not self._use_graph:
            self._step_body()
            return
        if self._graph is None:
            # warm-up on a side stream (allocator + cudaFuncSetAttribute), then capture once
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            lat0, xh0 = self.lat.clone(), self.x_half.clone()
            with torch.cuda.stream(s):
                self._step_body()
            torch.cuda.current_stream().wait_stream(s)
            self.lat.copy_(lat0)
            self.x_half.copy_(xh0)
            g = torch.cuda.CUDAGraph()
            n0 = self.ops.launch_count()
            with torch.cuda.graph(g):
                self._step_body()
            self.launches_per_step = self.ops.launch_count() - n0
            self._graph = g
            self.lat.copy_(lat0)
            self.x_half.copy_(xh0)
        self._graph.replay()

    def latents(self):
        return self.lat.permute(0, 3, 1, 2).contiguous()
