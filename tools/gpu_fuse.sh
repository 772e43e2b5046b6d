mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400 --csv --log-file gpurun_out/train_launches.csv python tools/exp_train.py 32 > gpurun_out/train_ncu.log 2>&1; echo "ncu exit $?"
python tools/launch_list.py gpurun_out/train_launches.csv | head -24
