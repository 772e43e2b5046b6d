#!/bin/bash
# iteration run: conv parity (tcgen05), smoke, short bench, optional ncu capture
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_models.py -m gpu -q --timeout=300 -k "not cudacore" -p no:cacheprovider > gpurun_out/models_tc.log 2>&1; echo "models tcgen05 exit $?"; tail -5 gpurun_out/models_tc.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --images ${IMAGES:-131072} --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_small.log 2>&1; echo "bench exit $?"
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/bench_small.log').read().strip().splitlines()[-1])
    print("value %.0f img/s  e2e %.0f  ms/step %.1f  clocks %s" % (d['value'], d['e2e']['value'], d['ms_per_step'], d['clocks']))
    print("roofline", {k: (round(v, 3) if isinstance(v, float) else v) for k, v in d['roofline'].items() if k != 'note'})
    for k, v in d['kernels'].items():
        print("  %-22s n=%-5d ms=%-9.3f share=%.3f  TF=%-8.2f GB/s=%.1f" % (k, v['launches'], v['ms'], v['share'], v['tflops_executed'], v['gbs']))
except Exception as e:
    print("parse failed", e); print(open('gpurun_out/bench_small.log').read()[-3000:])
PY
if [ -n "$NCU" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s ${NCU_SKIP:-20} -c ${NCU_COUNT:-10} -f -o gpurun_out/prof python bench.py --images 8192 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu.log 2>&1; echo "ncu exit $?"; tail -3 gpurun_out/ncu.log
fi
