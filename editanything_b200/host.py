"""Host-side semantics around the hot path: the id-coded SAM control map, image /
mask preparation and long-prompt embedding windows, restated from the reference's behaviour (not its code) so the
orchestrator in `editanything_b200.app` produces the same tensors the reference feeds its pipeline.

    HWC3 / resize_image / resize_points / get_bounding_box      annotator/util.py:9-71
    show_anns                                                   editany_lora.py:426-449   (ordering quirk kept)
    make_inpaint_condition                                      editany_lora.py:332-340
    get_pipeline_embeds                                         editany_lora.py:110-194
    seed_everything                                             pytorch_lightning.seed_everything as used at :789
Pinned against the reference's own functions in tests/test_host_semantics_cpu.py (run where /root/reference exists).
"""
import os
import random

import numpy as np
import torch

try:
    import cv2
except Exception:  # pragma: no cover
    cv2 = None
try:
    from PIL import Image
except Exception:  # pragma: no cover
    Image = None


def HWC3(x):
    """uint8 image of 1 / 3 / 4 channels (or H x W) -> H x W x 3; alpha is composited over white."""
    if x.dtype != np.uint8:
        raise AssertionError("HWC3 expects uint8")
    if x.ndim == 2:
        x = x[:, :, None]
    if x.ndim != 3 or x.shape[2] not in (1, 3, 4):
        raise AssertionError(f"HWC3: unsupported shape {x.shape}")
    c = x.shape[2]
    if c == 3:
        return x
    if c == 1:
        return np.concatenate([x, x, x], axis=2)
    rgb = x[:, :, :3].astype(np.float32)
    alpha = x[:, :, 3:4].astype(np.float32) / 255.0
    return (rgb * alpha + 255.0 * (1.0 - alpha)).clip(0, 255).astype(np.uint8)


def resize_image(input_image, resolution):
    """Short side -> `resolution`, both sides rounded to multiples of 64; Lanczos when enlarging, area when
    shrinking (annotator/util.py:28-38)."""
    h, w = float(input_image.shape[0]), float(input_image.shape[1])
    k = float(resolution) / min(h, w)
    nh = int(np.round(h * k / 64.0)) * 64
    nw = int(np.round(w * k / 64.0)) * 64
    return cv2.resize(input_image, (nw, nh), interpolation=cv2.INTER_LANCZOS4 if k > 1 else cv2.INTER_AREA)


def resize_points(clicked_points, original_shape, resolution):
    k = float(resolution) / min(float(original_shape[0]), float(original_shape[1]))
    return [(int(round(x * k)), int(round(y * k)), lab) for x, y, lab in clicked_points]


def get_bounding_box(mask):
    """[xmin, ymin, xmax, ymax] of the non-zero region of channel 0."""
    m = np.array(mask).astype(np.uint8)[:, :, 0]
    xs = np.where(np.any(m, axis=0))[0]
    ys = np.where(np.any(m, axis=1))[0]
    return [xs[0], ys[0], xs[-1], ys[-1]]


def show_anns(anns):
    """SAM masks -> (colour preview PIL image, id map `res` float64 [H, W, 3] with ch0 = id % 256, ch1 = id // 256).

    Quirk kept on purpose: the reference sorts the annotations by area but then indexes the
    UNSORTED list, so ids follow the generator's output order and later masks overwrite earlier ones
    (editany_lora.py:429-439).  One np.random.random((1, 3)) draw per mask, in that order."""
    if len(anns) == 0:
        return None
    full_img, idmap = None, None
    for i in range(len(anns)):
        m = anns[i]["segmentation"]
        if full_img is None:
            full_img = np.zeros((m.shape[0], m.shape[1], 3))
            idmap = np.zeros((m.shape[0], m.shape[1]), dtype=np.uint16)
        idmap[m != 0] = i + 1
        full_img[m != 0] = np.random.random((1, 3)).tolist()[0]
    res = np.zeros((idmap.shape[0], idmap.shape[1], 3))
    res[:, :, 0] = idmap % 256
    res[:, :, 1] = idmap // 256
    return Image.fromarray(np.uint8(full_img * 255)), res


def make_inpaint_condition(image, image_mask):
    """Inpaint-ControlNet conditioning: image / 255 with the masked pixels (mask > 128, any channel layout that
    broadcasts like the reference's boolean index) set to -1 -> float64 tensor [1, 3, H, W] (:332-340)."""
    img = image / 255.0
    if img.shape[0:1] != image_mask.shape[0:1]:
        raise AssertionError("image and image_mask must have the same image size")
    img[image_mask > 128] = -1.0
    return torch.from_numpy(np.expand_dims(img, 0).transpose(0, 3, 1, 2))


def get_pipeline_embeds(pipeline, prompt, negative_prompt, device):
    """Prompt embeddings without the 77-token limit: tokenise both prompts untruncated, pad the shorter one to the
    longer one's length, run the text encoder over windows of `model_max_length` tokens and concatenate along the
    sequence (editany_lora.py:160-194).  Returns (prompt_embeds, negative_prompt_embeds), both [1, L, D]."""
    tok, enc = pipeline.tokenizer, pipeline.text_encoder
    max_length = tok.model_max_length
    ids = tok(prompt, return_tensors="pt", truncation=False).input_ids.to(device)
    neg = tok(negative_prompt, return_tensors="pt", truncation=False).input_ids.to(device)
    total = max(ids.shape[-1], neg.shape[-1])
    if ids.shape[-1] > neg.shape[-1]:
        neg = tok(negative_prompt, truncation=False, padding="max_length", max_length=total,
                  return_tensors="pt").input_ids.to(device)
    else:
        ids = tok(prompt, return_tensors="pt", truncation=False, padding="max_length",
                  max_length=total).input_ids.to(device)
    pos_parts, neg_parts = [], []
    for i in range(0, total, max_length):
        pos_parts.append(enc(ids[:, i:i + max_length])[0])
        neg_parts.append(enc(neg[:, i:i + max_length])[0])
    return torch.cat(pos_parts, dim=1), torch.cat(neg_parts, dim=1)


def seed_everything(seed):
    """python / numpy / torch (CPU + CUDA) seeds and PL_GLOBAL_SEED, like pytorch_lightning.seed_everything."""
    seed = int(seed)
    os.environ["PL_GLOBAL_SEED"] = str(seed)
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    return seed
