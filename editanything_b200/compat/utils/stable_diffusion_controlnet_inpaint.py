from editanything_b200.pipeline import (StableDiffusionControlNetInpaintPipeline, StableDiffusionPipelineOutput,  # noqa: F401
                                        prepare_controlnet_conditioning_image, prepare_image, prepare_mask_image)
