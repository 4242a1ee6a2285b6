from editanything_b200.segment_anything import *  # noqa: F401,F403
from editanything_b200.segment_anything import (SamAutomaticMaskGenerator, SamPredictor, build_sam,  # noqa: F401
                                                sam_model_registry)
