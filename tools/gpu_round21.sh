set -x
cd /root/repo
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_vae.py -m gpu -q -s ) > gpurun_out/pytest_vae.log 2>&1
grep -v "^$" gpurun_out/pytest_vae.log | tail -25 | cut -c1-400

