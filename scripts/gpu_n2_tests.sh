set -u
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_sharded_gpu.py tests/test_crd_knn_gpu.py -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/r4g_pytest_2gpu.txt
