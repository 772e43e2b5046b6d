#!/bin/bash
# first GPU contact: exact kernels first, then the conv path (CUDA-core, then tcgen05), then smoke + a short bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for t in test_gpu_l2.py test_gpu_db.py; do
  timeout 600 python -m pytest tests/$t -m gpu -q --timeout=300 -p no:cacheprovider > gpurun_out/$t.log 2>&1; echo "$t exit $?"; tail -15 gpurun_out/$t.log
done
timeout 900 python -m pytest tests/test_gpu_models.py -m gpu -q --timeout=300 -k "cudacore" -p no:cacheprovider > gpurun_out/models_cudacore.log 2>&1; echo "models cudacore exit $?"; tail -25 gpurun_out/models_cudacore.log
timeout 900 python -m pytest tests/test_gpu_models.py -m gpu -q --timeout=300 -k "not cudacore" -p no:cacheprovider > gpurun_out/models_tc.log 2>&1; echo "models tcgen05 exit $?"; tail -40 gpurun_out/models_tc.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -5 gpurun_out/smoke.log
timeout 600 python bench.py --images 65536 --steps 2 --warmup 1 > gpurun_out/bench_small.log 2>&1; echo "bench exit $?"; tail -3 gpurun_out/bench_small.log
