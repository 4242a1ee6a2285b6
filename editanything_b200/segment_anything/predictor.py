"""`SamPredictor` (call sites editany_lora.py:528-541: `set_image(img)`, `predict(point_coords, point_labels,
multimask_output=False)` -> `(masks[C,H,W] bool, scores, low_res_logits)`) and the longest-side resize transform,
following the upstream segment_anything semantics."""
import numpy as np
import torch
from PIL import Image


class ResizeLongestSide:
    def __init__(self, target_length):
        self.target_length = target_length

    @staticmethod
    def get_preprocess_shape(oldh, oldw, long_side_length):
        scale = long_side_length * 1.0 / max(oldh, oldw)
        return int(oldh * scale + 0.5), int(oldw * scale + 0.5)

    def apply_image(self, image):
        """uint8 HWC -> uint8 HWC with the longest side = target_length (PIL bilinear, like upstream's
        torchvision resize of the PIL image)."""
        h, w = self.get_preprocess_shape(image.shape[0], image.shape[1], self.target_length)
        return np.array(Image.fromarray(image).resize((w, h), Image.BILINEAR))

    def apply_coords(self, coords, original_size):
        oh, ow = original_size
        nh, nw = self.get_preprocess_shape(oh, ow, self.target_length)
        coords = np.array(coords, dtype=float, copy=True)
        coords[..., 0] = coords[..., 0] * (nw / ow)
        coords[..., 1] = coords[..., 1] * (nh / oh)
        return coords

    def apply_boxes(self, boxes, original_size):
        return self.apply_coords(np.asarray(boxes).reshape(-1, 2, 2), original_size).reshape(-1, 4)

    def apply_coords_torch(self, coords, original_size):
        oh, ow = original_size
        nh, nw = self.get_preprocess_shape(oh, ow, self.target_length)
        coords = coords.clone().to(torch.float)
        coords[..., 0] = coords[..., 0] * (nw / ow)
        coords[..., 1] = coords[..., 1] * (nh / oh)
        return coords

    def apply_boxes_torch(self, boxes, original_size):
        return self.apply_coords_torch(boxes.reshape(-1, 2, 2), original_size).reshape(-1, 4)


class SamPredictor:
    def __init__(self, sam_model):
        self.model = sam_model
        self.transform = ResizeLongestSide(sam_model.image_encoder.img_size)
        self.reset_image()

    @property
    def device(self):
        return self.model.device

    def reset_image(self):
        self.is_image_set = False
        self.features = None
        self.orig_h = self.orig_w = self.input_h = self.input_w = None
        self.original_size = self.input_size = None

    def set_image(self, image, image_format="RGB"):
        if image_format not in ("RGB", "BGR"):
            raise AssertionError(f"image_format must be in ['RGB', 'BGR'], is {image_format}.")
        if image_format != self.model.image_format:
            image = image[..., ::-1]
        x = torch.as_tensor(self.transform.apply_image(np.ascontiguousarray(image)), device=self.device)
        self.set_torch_image(x.permute(2, 0, 1).contiguous()[None], image.shape[:2])

    @torch.no_grad()
    def set_torch_image(self, transformed_image, original_image_size):
        s = self.model.image_encoder.img_size
        if len(transformed_image.shape) != 4 or transformed_image.shape[1] != 3 or max(transformed_image.shape[2:]) != s:
            raise AssertionError(f"set_torch_image input must be BCHW with long side {s}.")
        self.reset_image()
        self.original_size = tuple(original_image_size)
        self.input_size = tuple(transformed_image.shape[-2:])
        self.features = self.model.image_encoder(self.model.preprocess(transformed_image.float()))
        self.is_image_set = True

    def predict(self, point_coords=None, point_labels=None, box=None, mask_input=None, multimask_output=True,
                return_logits=False):
        if not self.is_image_set:
            raise RuntimeError("An image must be set with .set_image(...) before mask prediction.")
        coords_t = labels_t = box_t = mask_t = None
        if point_coords is not None:
            if point_labels is None:
                raise AssertionError("point_labels must be supplied if point_coords is supplied.")
            pc = self.transform.apply_coords(point_coords, self.original_size)
            coords_t = torch.as_tensor(pc, dtype=torch.float, device=self.device)[None]
            labels_t = torch.as_tensor(point_labels, dtype=torch.int, device=self.device)[None]
        if box is not None:
            box_t = torch.as_tensor(self.transform.apply_boxes(box, self.original_size), dtype=torch.float,
                                    device=self.device)[None]
        if mask_input is not None:
            mask_t = torch.as_tensor(mask_input, dtype=torch.float, device=self.device)[None]
        masks, iou, low = self.predict_torch(coords_t, labels_t, box_t, mask_t, multimask_output, return_logits)
        return masks[0].detach().cpu().numpy(), iou[0].detach().cpu().numpy(), low[0].detach().cpu().numpy()

    @torch.no_grad()
    def predict_torch(self, point_coords, point_labels, boxes=None, mask_input=None, multimask_output=True,
                      return_logits=False):
        if not self.is_image_set:
            raise RuntimeError("An image must be set with .set_image(...) before mask prediction.")
        points = (point_coords, point_labels) if point_coords is not None else None
        sparse, dense = self.model.prompt_encoder(points=points, boxes=boxes, masks=mask_input)
        low, iou = self.model.mask_decoder(image_embeddings=self.features, image_pe=self.model.prompt_encoder.get_dense_pe(),
                                           sparse_prompt_embeddings=sparse, dense_prompt_embeddings=dense,
                                           multimask_output=multimask_output)
        masks = self.model.postprocess_masks(low, self.input_size, self.original_size)
        if not return_logits:
            masks = masks > self.model.mask_threshold
        return masks, iou, low

    def get_image_embedding(self):
        if not self.is_image_set:
            raise RuntimeError("An image must be set with .set_image(...) to generate an embedding.")
        return self.features
