mkdir -p gpurun_out/final
timeout 900 python bench.py --config 5 > gpurun_out/final/bench_cfg5.json 2> gpurun_out/final/bench_cfg5.err; echo "cfg5 exit $?"; tail -c 300 gpurun_out/final/bench_cfg5.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/final/bench_cfg5.json').read().strip().splitlines()[-1])
print({k: d.get(k) for k in ('value','ms_per_step','verified','clocks')}); print('e2e', d['e2e']['value']); print('config5', json.dumps(d.get('config5'))[:700])
for k, v in d['kernels'].items():
    if v['share'] > 0.02: print("  %-22s ms=%-9.3f share=%.3f  TF=%-8.2f" % (k, v['ms'], v['share'], v['tflops_executed']))
PY
timeout 600 ncu --set full --clock-control none -k regex:"conv3x3|bn_stats|bn_bwd_reduce|linear_bwd_data" -s 60 -c 24 -f -o /tmp/prof_train python tools/exp_train.py 32 > gpurun_out/final/ncu_train.log 2>&1; echo "ncu train exit $?"
python tools/ncu_summary.py /tmp/prof_train.ncu-rep > gpurun_out/final/ncu_train.md; cat gpurun_out/final/ncu_train.md | cut -c1-220
