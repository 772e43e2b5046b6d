#!/bin/bash
# full GPU pass: all gpu tests, smoke, the default bench (1M faces), the reference arm, ncu launch list + full capture
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest gpu exit $?"; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; echo "bench exit $?"; tail -c 600 gpurun_out/bench_1gpu.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref exit $?"; cat gpurun_out/bench_ref.json | cut -c1-400
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_1gpu.json').read().strip().splitlines()[-1])
print("value %.0f img/s  e2e %.0f  ms/step %.1f  search q/s %.0f clocks %s" % (d['value'], d['e2e']['value'], d['ms_per_step'], d['search_queries_per_sec'], d['clocks']))
print("cpu_baseline", d.get('cpu_baseline'))
print("roofline", {k: (round(v, 3) if isinstance(v, float) else v) for k, v in d['roofline'].items() if k != 'note'})
for k, v in d['kernels'].items():
    print("  %-22s n=%-5d ms=%-9.3f share=%.3f  TF=%-8.2f GB/s=%.1f" % (k, v['launches'], v['ms'], v['share'], v['tflops_executed'], v['gbs']))
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --images 16384 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 0 -c 11 -f -o gpurun_out/prof_r01c python bench.py --images 8192 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu full exit $?"
