for rep in 1 2 3; do
echo "--- store_early=1 (default lib)"; timeout 300 python tools/ab_option.py tma_store 1 2>&1 | tail -1
echo "--- store_early=0"; GANREV_CUDA_LIB=$PWD/gan-reverser_b200/libganrev_cuda_trace.so timeout 300 python tools/ab_option.py tma_store 1 2>&1 | tail -1
done
