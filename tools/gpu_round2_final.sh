#!/bin/bash
# round-2 full pass: all gpu tests, smoke, default bench, reference arm, launch list, ncu --set full of one chunk of G and R
mkdir -p gpurun_out/final
O=gpurun_out/final
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "pytest gpu exit $?"; tail -4 $O/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke exit $?"; tail -2 $O/smoke.log
timeout 900 python bench.py > $O/bench_1gpu.json 2> $O/bench_1gpu.err; echo "bench exit $?"; tail -c 600 $O/bench_1gpu.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "bench ref exit $?"; cut -c1-300 $O/bench_ref.json
python - <<'PY'
import json
d = json.loads(open('gpurun_out/final/bench_1gpu.json').read().strip().splitlines()[-1])
print("value %.0f img/s  e2e %.0f  ms/step %.1f  search q/s %.0f clocks %s verified %s" % (d['value'], d['e2e']['value'], d['ms_per_step'], d['search_queries_per_sec'], d['clocks'], d.get('verified')))
print("roofline", {k: (round(v, 3) if isinstance(v, float) else v) for k, v in d['roofline'].items() if k != 'note'})
for k, v in d['kernels'].items():
    print("  %-22s n=%-5d ms=%-9.3f share=%.3f  TF=%-8.2f GB/s=%.1f" % (k, v['launches'], v['ms'], v['share'], v['tflops_executed'], v['gbs']))
PY
if [ -n "$NCU" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches.csv python bench.py --images 16384 --steps 1 --warmup 1 --no-cpu-baseline > $O/ncu_launches.log 2>&1; echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_tc|gather|r_conv1" -s 0 -c 13 -f -o /tmp/prof_final python bench.py --images 8192 --steps 1 --warmup 1 --no-cpu-baseline > $O/ncu_full.log 2>&1; echo "ncu full exit $?"
python tools/ncu_summary.py /tmp/prof_final.ncu-rep > $O/ncu_full_final.md 2>> $O/ncu_full.log; cat $O/ncu_full_final.md
fi
