mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_models.py -m gpu -q --timeout=300 -x -p no:cacheprovider -k "apply_r_main" > gpurun_out/models.log 2>&1; echo "exit $?"; tail -12 gpurun_out/models.log
