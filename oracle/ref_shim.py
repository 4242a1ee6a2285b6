"""Import the REAL reference modules from /root/reference (build container only).

The reference needs two third-party packages that are absent here only for import-time glue
(SURVEY.md §8c): omegaconf (ListConfig type check) and pytorch_lightning (base class of modules we
never instantiate).  They are stubbed; the arithmetic modules (ldm/cldm) run unmodified.
Used by oracle/make_golden.py and tests/test_oracle_vs_reference.py; never on the GPU box.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("EA_REFERENCE_ROOT", "/root/reference")
_LOCAL = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")     # oracle/build_ref.py (travels to the GPU box)
if not os.path.isdir(os.path.join(REFERENCE_ROOT, "cldm")) and os.path.isdir(os.path.join(_LOCAL, "cldm")):
    REFERENCE_ROOT = _LOCAL


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "cldm"))


def load():
    """Returns the reference's cldm.cldm module (ControlledUnetModel, ControlNet)."""
    import torch
    if not available():
        raise RuntimeError("reference tree not present")
    if "omegaconf" not in sys.modules:
        oc = types.ModuleType("omegaconf")
        lc = types.ModuleType("omegaconf.listconfig")

        class ListConfig(list):
            pass
        lc.ListConfig = oc.ListConfig = ListConfig
        oc.listconfig = lc
        oc.OmegaConf = object
        sys.modules["omegaconf"] = oc
        sys.modules["omegaconf.listconfig"] = lc
    if "pytorch_lightning" not in sys.modules:
        pl = types.ModuleType("pytorch_lightning")
        pl.LightningModule = torch.nn.Module
        plu = types.ModuleType("pytorch_lightning.utilities")
        pld = types.ModuleType("pytorch_lightning.utilities.distributed")
        pld.rank_zero_only = lambda f: f
        sys.modules.update({"pytorch_lightning": pl, "pytorch_lightning.utilities": plu,
                            "pytorch_lightning.utilities.distributed": pld})
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import contextlib
    with contextlib.redirect_stdout(sys.stderr):   # the reference prints at import ("No module 'xformers' ..."):
        import cldm.cldm as C                       # keep stdout clean for bench.py's one JSON line
    return C


def ctor_kwargs(cfg):
    """UNetConfig -> keyword arguments of the reference constructors."""
    kw = dict(image_size=32, in_channels=cfg.in_channels, model_channels=cfg.model_channels,
              attention_resolutions=list(cfg.attention_resolutions), num_res_blocks=cfg.num_res_blocks,
              channel_mult=list(cfg.channel_mult), use_spatial_transformer=True, transformer_depth=1,
              context_dim=cfg.context_dim, legacy=False, use_checkpoint=False,
              use_linear_in_transformer=cfg.use_linear_in_transformer)
    if cfg.num_head_channels == -1:
        kw["num_heads"] = cfg.num_heads
    else:
        kw["num_head_channels"] = cfg.num_head_channels
    return kw


def build_reference_nets(cfg, unet_sd, cn_sds):
    """Instantiate the reference UNet + ControlNets and load the given state dicts (strict)."""
    C = load()
    kw = ctor_kwargs(cfg)
    unet = C.ControlledUnetModel(out_channels=cfg.out_channels, **kw).eval()
    unet.load_state_dict(unet_sd, strict=True)
    cns = []
    for sd in cn_sds:
        cn = C.ControlNet(hint_channels=cfg.hint_channels, **kw).eval()
        cn.load_state_dict(sd, strict=True)
        cns.append(cn)
    return unet, cns


def reference_apply_model(unet, cns, x, t, ctx, hints, scales):
    """ControlLDM.apply_model arithmetic (cldm/cldm.py:328-341) with the reference modules; several
    ControlNets are summed like diffusers' MultiControlNetModel."""
    import torch
    with torch.no_grad():
        control = None
        for cn, hint, s in zip(cns, hints, scales):
            outs = cn(x=x, hint=hint, timesteps=t, context=ctx)
            outs = [o * s for o in outs]
            control = outs if control is None else [a + b for a, b in zip(control, outs)]
        return unet(x=x, timesteps=t, context=ctx, control=list(control) if control else None,
                    only_mid_control=False), control
