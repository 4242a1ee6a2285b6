"""One SAM ViT-H encode under cudaProfilerStart/Stop (for `ncu --profile-from-start off`)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from editanything_b200.sam import SamEncoderEngine, SAM_VIT_H, make_sam_state_dict  # noqa: E402

dev = torch.device("cuda:0")
sam = SamEncoderEngine(SAM_VIT_H, make_sam_state_dict(SAM_VIT_H, 201, device=dev), dev)
img = torch.randn(1, 3, 1024, 1024, device=dev)
sam.encode(img)
torch.cuda.synchronize()
torch.cuda.profiler.start()
sam.encode(img)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
