mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_models.py tests/test_gpu_l2.py -m gpu -q --timeout=300 -x -p no:cacheprovider > gpurun_out/models.log 2>&1; echo "models exit $?"; tail -4 gpurun_out/models.log
timeout 300 python tools/ab_option.py fuse_conv3 1 2>&1 | tail -2
GEOM=3,64,64,256,4096 timeout 300 python tools/ab_option.py fuse_conv3 1 2>&1 | tail -2
