"""Import-compatible stand-in for the `segment_anything` package the reference depends on (an un-vendored git
dependency: editany_lora.py:44-66):

    from segment_anything import sam_model_registry, SamAutomaticMaskGenerator, SamPredictor      # reference
    from editanything_b200.segment_anything import ...                                             # this backend

`sam_model_registry["default"](checkpoint=path)` builds a `Sam` whose image encoder is the B200 engine
(`editanything_b200.sam.SamEncoderEngine`, built when the model is moved to a CUDA device) and whose prompt encoder /
mask decoder are PyTorch modules with upstream parameter names, so the official .pth checkpoints load as they are."""
import torch

from ..sam_spec import SAM_VIT_H, SamEncoderConfig
from .amg import SamAutomaticMaskGenerator, remove_small_regions  # noqa: F401
from .modeling import MaskDecoder, PromptEncoder, Sam, TwoWayTransformer, _EncoderModule
from .predictor import ResizeLongestSide, SamPredictor  # noqa: F401

SAM_VIT_L = SamEncoderConfig(embed_dim=1024, depth=24, num_heads=16, mlp_dim=4096, global_attn_indexes=(5, 11, 17, 23))
SAM_VIT_B = SamEncoderConfig(embed_dim=768, depth=12, num_heads=12, mlp_dim=3072, global_attn_indexes=(2, 5, 8, 11))


def build_sam_from_state_dict(cfg: SamEncoderConfig, state_dict):
    """`state_dict`: upstream names (image_encoder.*, prompt_encoder.*, mask_decoder.*)."""
    pe_dim = cfg.out_chans
    enc_sd = {k: v for k, v in state_dict.items() if k.startswith("image_encoder.")}
    sam = Sam(_EncoderModule(cfg, enc_sd),
              PromptEncoder(embed_dim=pe_dim, image_embedding_size=(cfg.grid, cfg.grid),
                            input_image_size=(cfg.img_size, cfg.img_size), mask_in_chans=16),
              MaskDecoder(transformer_dim=pe_dim, transformer=TwoWayTransformer(2, pe_dim, 8, 2048)))
    rest = {k: v for k, v in state_dict.items() if not k.startswith("image_encoder.")}
    if rest:
        missing, unexpected = sam.load_state_dict(rest, strict=False)
        bad = [k for k in missing if not k.startswith("image_encoder.")] + list(unexpected)
        if bad:
            raise KeyError(f"SAM checkpoint does not match: {bad[:5]}")
    return sam.eval()


def _builder(cfg):
    def build(checkpoint=None):
        if checkpoint is None:
            raise ValueError("a checkpoint path is required (synthetic weights: build_sam_from_state_dict)")
        sd = torch.load(checkpoint, map_location="cpu", weights_only=True)
        return build_sam_from_state_dict(cfg, sd)
    return build


build_sam_vit_h, build_sam_vit_l, build_sam_vit_b = _builder(SAM_VIT_H), _builder(SAM_VIT_L), _builder(SAM_VIT_B)
build_sam = build_sam_vit_h
sam_model_registry = {"default": build_sam_vit_h, "vit_h": build_sam_vit_h, "vit_l": build_sam_vit_l, "vit_b": build_sam_vit_b}
