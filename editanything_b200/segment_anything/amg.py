"""`SamAutomaticMaskGenerator` (call site editany_lora.py:523: `.generate(np.uint8 HWC RGB)` -> list of dicts with
"segmentation" (bool H x W), "area", "bbox", "predicted_iou", "point_coords", "stability_score", "crop_box").

The automatic mask generation procedure of the upstream package, restated: a points_per_side^2 grid of single-point
prompts per crop in batches of points_per_batch, three masks per point, filtered by predicted IoU, by the stability
score (IoU of the mask thresholded at +-offset), by touching a crop edge, then box NMS; crops beyond layer 0 and
small-region post-processing as upstream.  Masks stay on the device until the final conversion.  Upstream is absent
in this container: parity unpinned against it (the decoder arithmetic is pinned against the HF port)."""
import math

import numpy as np
import torch

try:
    from torchvision.ops.boxes import batched_nms, box_area
except Exception:  # pragma: no cover
    batched_nms = box_area = None

from .predictor import SamPredictor


def build_point_grid(n_per_side):
    off = 1 / (2 * n_per_side)
    pts = np.linspace(off, 1 - off, n_per_side)
    return np.stack([np.tile(pts[None, :], (n_per_side, 1)), np.tile(pts[:, None], (1, n_per_side))], axis=-1).reshape(-1, 2)


def build_all_layer_point_grids(n_per_side, n_layers, scale_per_layer):
    return [build_point_grid(int(n_per_side / (scale_per_layer ** i))) for i in range(n_layers + 1)]


def generate_crop_boxes(im_size, n_layers, overlap_ratio):
    im_h, im_w = im_size
    short = min(im_h, im_w)
    boxes, layer_idxs = [[0, 0, im_w, im_h]], [0]

    def crop_len(orig_len, n_crops, overlap):
        return int(math.ceil((overlap * (n_crops - 1) + orig_len) / n_crops))
    for i_layer in range(n_layers):
        n = 2 ** (i_layer + 1)
        overlap = int(overlap_ratio * short * (2 / n))
        cw, ch = crop_len(im_w, n, overlap), crop_len(im_h, n, overlap)
        for x0 in [int((cw - overlap) * i) for i in range(n)]:
            for y0 in [int((ch - overlap) * i) for i in range(n)]:
                boxes.append([x0, y0, min(x0 + cw, im_w), min(y0 + ch, im_h)])
                layer_idxs.append(i_layer + 1)
    return boxes, layer_idxs


def calculate_stability_score(masks, mask_threshold, offset):
    inter = (masks > (mask_threshold + offset)).sum(-1, dtype=torch.int16).sum(-1, dtype=torch.int32)
    union = (masks > (mask_threshold - offset)).sum(-1, dtype=torch.int16).sum(-1, dtype=torch.int32)
    return inter / union


def batched_mask_to_box(masks):
    """bool [..., H, W] -> XYXY boxes [..., 4]; empty masks -> [0, 0, 0, 0]."""
    if masks.numel() == 0:
        return torch.zeros(*masks.shape[:-2], 4, device=masks.device)
    shape = masks.shape
    h, w = shape[-2:]
    m = masks.flatten(0, -3) if len(shape) > 2 else masks.unsqueeze(0)
    in_h, _ = torch.max(m, dim=-1)
    hc = in_h * torch.arange(h, device=m.device)[None, :]
    bottom, _ = torch.max(hc, dim=-1)
    top, _ = torch.min(hc + h * (~in_h), dim=-1)
    in_w, _ = torch.max(m, dim=-2)
    wc = in_w * torch.arange(w, device=m.device)[None, :]
    right, _ = torch.max(wc, dim=-1)
    left, _ = torch.min(wc + w * (~in_w), dim=-1)
    empty = (right < left) | (bottom < top)
    out = torch.stack([left, top, right, bottom], dim=-1) * (~empty).unsqueeze(-1)
    return out.reshape(*shape[:-2], 4) if len(shape) > 2 else out[0]


def is_box_near_crop_edge(boxes, crop_box, orig_box, atol=20.0):
    crop_t = torch.as_tensor(crop_box, dtype=torch.float, device=boxes.device)
    orig_t = torch.as_tensor(orig_box, dtype=torch.float, device=boxes.device)
    b = uncrop_boxes_xyxy(boxes, crop_box).float()
    near_crop = torch.isclose(b, crop_t[None, :], atol=atol, rtol=0)
    near_img = torch.isclose(b, orig_t[None, :], atol=atol, rtol=0)
    return torch.any(torch.logical_and(near_crop, ~near_img), dim=1)


def uncrop_boxes_xyxy(boxes, crop_box):
    x0, y0 = crop_box[0], crop_box[1]
    off = torch.tensor([[x0, y0, x0, y0]], device=boxes.device)
    if len(boxes.shape) == 3:
        off = off.unsqueeze(1)
    return boxes + off


def uncrop_points(points, crop_box):
    off = torch.tensor([[crop_box[0], crop_box[1]]], device=points.device)
    if len(points.shape) == 3:
        off = off.unsqueeze(1)
    return points + off


def uncrop_masks(masks, crop_box, orig_h, orig_w):
    x0, y0, x1, y1 = crop_box
    if x0 == 0 and y0 == 0 and x1 == orig_w and y1 == orig_h:
        return masks
    return torch.nn.functional.pad(masks, (x0, orig_w - (x1 - x0) - x0, y0, orig_h - (y1 - y0) - y0), value=0)


def remove_small_regions(mask, area_thresh, mode):
    """Remove small disconnected regions ("islands") or fill small holes ("holes") of a bool mask (needs cv2)."""
    import cv2
    if mode not in ("holes", "islands"):
        raise AssertionError(mode)
    correct_holes = mode == "holes"
    working = (correct_holes ^ mask).astype(np.uint8)
    n, regions, stats, _ = cv2.connectedComponentsWithStats(working, 8)
    sizes = stats[:, -1][1:]
    small = [i + 1 for i, s in enumerate(sizes) if s < area_thresh]
    if len(small) == 0:
        return mask, False
    fill = [0] + small
    if not correct_holes:
        fill = [i for i in range(n) if i not in fill]
        if len(fill) == 0:
            fill = [int(np.argmax(sizes)) + 1]
    return np.isin(regions, fill), True


class SamAutomaticMaskGenerator:
    def __init__(self, model, points_per_side=32, points_per_batch=64, pred_iou_thresh=0.88, stability_score_thresh=0.95,
                 stability_score_offset=1.0, box_nms_thresh=0.7, crop_n_layers=0, crop_nms_thresh=0.7,
                 crop_overlap_ratio=512 / 1500, crop_n_points_downscale_factor=1, point_grids=None,
                 min_mask_region_area=0, output_mode="binary_mask"):
        if (points_per_side is None) == (point_grids is None):
            raise AssertionError("Exactly one of points_per_side or point_grid must be provided.")
        if output_mode != "binary_mask":
            raise NotImplementedError("output_mode: only 'binary_mask' (what editany_lora.py:429-439 consumes)")
        self.point_grids = (build_all_layer_point_grids(points_per_side, crop_n_layers, crop_n_points_downscale_factor)
                            if points_per_side is not None else point_grids)
        self.predictor = SamPredictor(model)
        self.points_per_batch, self.pred_iou_thresh = points_per_batch, pred_iou_thresh
        self.stability_score_thresh, self.stability_score_offset = stability_score_thresh, stability_score_offset
        self.box_nms_thresh, self.crop_n_layers, self.crop_nms_thresh = box_nms_thresh, crop_n_layers, crop_nms_thresh
        self.crop_overlap_ratio, self.crop_n_points_downscale_factor = crop_overlap_ratio, crop_n_points_downscale_factor
        self.min_mask_region_area, self.output_mode = min_mask_region_area, output_mode

    @torch.no_grad()
    def generate(self, image):
        data = self._generate_masks(image)
        if self.min_mask_region_area > 0:
            data = self._postprocess_small_regions(data, self.min_mask_region_area, max(self.box_nms_thresh, self.crop_nms_thresh))
        masks = data["masks"].cpu().numpy()
        boxes = data["boxes"].cpu()
        xywh = boxes.clone()
        xywh[:, 2] -= xywh[:, 0]
        xywh[:, 3] -= xywh[:, 1]
        out = []
        for i in range(masks.shape[0]):
            out.append({"segmentation": masks[i], "area": int(masks[i].sum()), "bbox": xywh[i].tolist(),
                        "predicted_iou": float(data["iou_preds"][i]), "point_coords": [data["points"][i].tolist()],
                        "stability_score": float(data["stability_score"][i]),
                        "crop_box": [data["crop_boxes"][i][0], data["crop_boxes"][i][1],
                                     data["crop_boxes"][i][2] - data["crop_boxes"][i][0],
                                     data["crop_boxes"][i][3] - data["crop_boxes"][i][1]]})
        return out

    # ---------------------------------------------------------------------------------------------------------
    @staticmethod
    def _cat(parts):
        keys = parts[0].keys()
        out = {}
        for k in keys:
            v = [p[k] for p in parts]
            out[k] = torch.cat(v, 0) if torch.is_tensor(v[0]) else sum(v, [])
        return out

    @staticmethod
    def _filter(d, keep):
        idx = keep.nonzero().flatten().tolist() if keep.dtype == torch.bool else keep.tolist()
        return {k: (v[keep] if torch.is_tensor(v) else [v[i] for i in idx]) for k, v in d.items()}

    def _generate_masks(self, image):
        orig_size = image.shape[:2]
        crop_boxes, layer_idxs = generate_crop_boxes(orig_size, self.crop_n_layers, self.crop_overlap_ratio)
        parts = [self._process_crop(image, cb, li, orig_size) for cb, li in zip(crop_boxes, layer_idxs)]
        data = self._cat(parts)
        if len(crop_boxes) > 1 and data["boxes"].shape[0] > 0:     # prefer masks from smaller crops
            cb = torch.tensor(data["crop_boxes"], device=data["boxes"].device, dtype=torch.float)
            scores = 1 / box_area(cb)
            keep = batched_nms(data["boxes"].float(), scores, torch.zeros_like(data["boxes"][:, 0]), iou_threshold=self.crop_nms_thresh)
            data = self._filter(data, keep)
        return data

    def _process_crop(self, image, crop_box, crop_layer_idx, orig_size):
        x0, y0, x1, y1 = crop_box
        cropped = image[y0:y1, x0:x1, :]
        ch, cw = cropped.shape[:2]
        self.predictor.set_image(cropped)
        pts = self.point_grids[crop_layer_idx] * np.array([cw, ch])[None, :]
        parts = []
        for i in range(0, len(pts), self.points_per_batch):
            parts.append(self._process_batch(pts[i:i + self.points_per_batch], (ch, cw), crop_box, orig_size))
        self.predictor.reset_image()
        data = self._cat(parts)
        keep = batched_nms(data["boxes"].float(), data["iou_preds"], torch.zeros_like(data["boxes"][:, 0]),
                           iou_threshold=self.box_nms_thresh)
        data = self._filter(data, keep)
        data["boxes"] = uncrop_boxes_xyxy(data["boxes"], crop_box)
        data["points"] = uncrop_points(data["points"], crop_box)
        data["masks"] = uncrop_masks(data["masks"], crop_box, orig_size[0], orig_size[1])
        data["crop_boxes"] = [list(crop_box) for _ in range(data["boxes"].shape[0])]
        return data

    def _process_batch(self, points, im_size, crop_box, orig_size):
        orig_h, orig_w = orig_size
        dev = self.predictor.device
        tp = self.predictor.transform.apply_coords(points, im_size)
        in_points = torch.as_tensor(tp, device=dev, dtype=torch.float)
        in_labels = torch.ones(in_points.shape[0], dtype=torch.int, device=dev)
        masks, iou, _ = self.predictor.predict_torch(in_points[:, None, :], in_labels[:, None], multimask_output=True,
                                                     return_logits=True)
        data = {"masks": masks.flatten(0, 1), "iou_preds": iou.flatten(0, 1),
                "points": torch.as_tensor(np.repeat(points, masks.shape[1], axis=0), device=dev)}
        del masks
        if self.pred_iou_thresh > 0.0:
            data = self._filter(data, data["iou_preds"] > self.pred_iou_thresh)
        data["stability_score"] = calculate_stability_score(data["masks"], self.predictor.model.mask_threshold,
                                                            self.stability_score_offset)
        if self.stability_score_thresh > 0.0:
            data = self._filter(data, data["stability_score"] >= self.stability_score_thresh)
        data["masks"] = data["masks"] > self.predictor.model.mask_threshold
        data["boxes"] = batched_mask_to_box(data["masks"])
        keep = ~is_box_near_crop_edge(data["boxes"], crop_box, [0, 0, orig_w, orig_h])
        if not torch.all(keep):
            data = self._filter(data, keep)
        return data

    def _postprocess_small_regions(self, data, min_area, nms_thresh):
        if data["masks"].shape[0] == 0:
            return data
        new_masks, scores = [], []
        for m in data["masks"].cpu().numpy():
            m, changed = remove_small_regions(m, min_area, mode="holes")
            unchanged = not changed
            m, changed = remove_small_regions(m, min_area, mode="islands")
            unchanged = unchanged and not changed
            new_masks.append(torch.as_tensor(m).unsqueeze(0))
            scores.append(float(unchanged))          # unchanged masks win the NMS
        masks = torch.cat(new_masks, dim=0).to(data["masks"].device)
        boxes = batched_mask_to_box(masks)
        keep = batched_nms(boxes.float(), torch.as_tensor(scores, device=boxes.device), torch.zeros_like(boxes[:, 0]),
                           iou_threshold=nms_thresh)
        for i in keep.tolist():
            if scores[i] == 0.0:
                data["masks"][i] = masks[i]
                data["boxes"][i] = boxes[i]
        return self._filter(data, keep)
