mkdir -p gpurun_out
timeout 300 python tools/ab_option.py tma_hybrid 0 1 2>&1 | tail -4
