set -x
cd /root/repo
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_vae.py -m gpu -q -s ) > gpurun_out/pytest_vae.log 2>&1
grep -v "^$" gpurun_out/pytest_vae.log | tail -30 | cut -c1-400
( time timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_vae.py ) > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log | head -3
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench.err | cut -c1-1300; tail -3 gpurun_out/bench.err
