mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -q --timeout=300 -x -p no:cacheprovider > gpurun_out/train.log 2>&1; echo "train tests exit $?"; tail -5 gpurun_out/train.log
timeout 300 python tools/exp_train.py 32 2>&1 | tail -2
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400 --csv --log-file gpurun_out/train_launches.csv python tools/exp_train.py 32 > gpurun_out/train_ncu.log 2>&1; echo "ncu exit $?"
python tools/launch_list.py gpurun_out/train_launches.csv 2>/dev/null | head -16
