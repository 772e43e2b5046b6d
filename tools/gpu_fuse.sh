mkdir -p gpurun_out/final
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider > gpurun_out/final/pytest_gpu.log 2>&1; echo "pytest gpu exit $?"; tail -3 gpurun_out/final/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/final/smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/final/smoke.log
