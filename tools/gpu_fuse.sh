mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_models.py -m gpu -q --timeout=300 -x -p no:cacheprovider -k "R_matches or benchmark or resident" > gpurun_out/models.log 2>&1; echo "exit $?"; tail -3 gpurun_out/models.log
for rep in 1 2 3; do timeout 300 python tools/ab_option.py tma_store 1 2>&1 | tail -1; done
