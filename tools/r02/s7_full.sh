# round 2, GPU session 7: session 6 re-run with the persistent kernel back at 168 registers (3 warps per SM sub-partition cap), exp2 polynomial in attention
set -x
cd /root/repo
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q ) > gpurun_out/s7_pytest_gpu.log 2>&1
tail -12 gpurun_out/s7_pytest_gpu.log
grep -h "max-abs" gpurun_out/s7_pytest_gpu.log | tail -6
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/s7_bench.json 2> gpurun_out/s7_bench.err; tail -3 gpurun_out/s7_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/s7_bench.json'))
print('ms_step', d['ms_per_step'], 'launches', d['launches_per_step'], 'roofline', d['roofline']['frac'], d['roofline']['gemm_ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_image'], 'image_ms', d['config']['image_ms'], 'value', d['value'], 'prepare_ms', d['config'].get('prepare_ms_per_request'), 'sam', d['config']['sam_ms_per_image'], 'vae', d['config']['vae_decode_ms_per_image'], d['config']['vae_encode_ms_per_image'])
print('batch4', d['config'].get('batch4'))
PY
EA_HINT_TC=0 timeout 600 python - <<'PY'
import sys, torch, json
sys.path.insert(0, '.')
import bench
from editanything_b200.denoise import DenoiseEngine
from editanything_b200.unet_spec import SD15, make_state_dict
dev = torch.device('cuda:0')
eng = DenoiseEngine(SD15, make_state_dict(SD15, 'unet', 101, device=dev), [make_state_dict(SD15, 'controlnet', 102, device=dev), make_state_dict(SD15, 'controlnet', 103, device=dev)], dev)
x, ctx, hints = bench.make_inputs(SD15, 2, 64, 77, 11)
for dup in (False, True):
    for _ in range(2): eng.prepare(ctx, hints, [0.5, 1.0], cfg_duplicated=dup)
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): eng.prepare(ctx, hints, [0.5, 1.0], cfg_duplicated=dup)
    e1.record(); torch.cuda.synchronize()
    print('prepare with the CUDA-core hint stack, cfg_duplicated', dup, round(e0.elapsed_time(e1) / 3, 3), 'ms')
PY
timeout 500 python tools/gemm_breakdown.py gpurun_out/s7_gemm_breakdown.json 2>&1 | head -24
