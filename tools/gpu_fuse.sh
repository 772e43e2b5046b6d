mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 77 python -m pytest tests/test_gpu_train.py -m gpu -q --timeout=800 -x -p no:cacheprovider -k "matches_torch and 8-False" > gpurun_out/racecheck_train.log 2>&1; echo "racecheck train exit $?"; grep -E "RACECHECK SUMMARY|hazard|passed|failed" gpurun_out/racecheck_train.log | head -10
