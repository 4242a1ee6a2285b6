set -x
cd /root/repo
mkdir -p gpurun_out
timeout 300 python - <<'PY' 2>&1 | tail -20
import torch
from editanything_b200 import ops, _lib as L
dt = ops.half_dtype()
a = torch.randn(256, 768, device='cuda').to(dt); w = torch.randn(160, 768, device='cuda').to(dt)
for fp in (-1, 1, 2):
    try:
        o = ops.gemm(a, w, force_persistent=fp)
        torch.cuda.synchronize()
        print('force_persistent', fp, 'ok', float((o.float() - a.float() @ w.float().t()).abs().max()))
    except Exception as e:
        print('force_persistent', fp, 'FAILED', e)
import ctypes
print(torch.cuda.get_device_properties(0))
PY
cuobjdump -res-usage editanything_b200/lib/libea_b200.so 2>/dev/null | grep -A1 "persistent" | grep -E "Function|REG" | head -8
