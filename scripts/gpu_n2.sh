set -u
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_sharded_gpu.py tests/test_crd_gpu.py -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/r3i_pytest_2gpu.txt
bash scripts/gpu_sharded.sh r3i 2 bench
