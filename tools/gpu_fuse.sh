mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q --timeout=600 -x -p no:cacheprovider > gpurun_out/train.log 2>&1; echo "train tests exit $?"; tail -5 gpurun_out/train.log
timeout 300 python tools/exp_train.py 32 2>&1 | tail -2
