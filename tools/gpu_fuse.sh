mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_models.py -m gpu -q --timeout=300 -x -p no:cacheprovider > gpurun_out/models.log 2>&1; echo "exit $?"; tail -3 gpurun_out/models.log
timeout 300 python tools/ab_total.py pdl 0 1 2>&1 | tail -6
