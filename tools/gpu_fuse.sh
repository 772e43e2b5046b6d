mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_models.py -m gpu -q --timeout=300 -x -p no:cacheprovider > gpurun_out/models.log 2>&1; echo "models exit $?"; tail -3 gpurun_out/models.log
for rep in 1 2 3; do
echo "--- new"; timeout 300 python tools/ab_option.py tma_store 1 2>&1 | tail -1
echo "--- before (pool_direct=0, cvt packs)"; GANREV_CUDA_LIB=$PWD/gan-reverser_b200/libganrev_cuda_trace.so timeout 300 python tools/ab_option.py tma_store 1 2>&1 | tail -1
done
