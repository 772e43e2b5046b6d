mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_models.py -m gpu -q --timeout=300 -x -p no:cacheprovider > gpurun_out/models.log 2>&1; echo "models exit $?"; tail -4 gpurun_out/models.log
timeout 300 python tools/ab_option.py xpose2 0 1 2 2>&1 | tail -6
