set -x
mkdir -p gpurun_out/r2v
timeout 900 python -m pytest tests/test_gpu_configs.py -x -q -k "code_default" > gpurun_out/r2v/pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2v/pytest.log
tail -15 gpurun_out/r2v/pytest.log
