mkdir -p gpurun_out/final
timeout 900 python bench.py --config 5 > gpurun_out/final/bench_cfg5.json 2> gpurun_out/final/bench_cfg5.err; echo "cfg5 exit $?"; tail -c 300 gpurun_out/final/bench_cfg5.err
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/final/bench_cfg5.json').read().splitlines() if l.startswith('{"')][-1])
print({k: d.get(k) for k in ('value','ms_per_step','verified','clocks')}); print('e2e', d['e2e']['value']); print('kmeans', json.dumps(d['config5']['kmeans'])); print('search ms', d['config5']['search']['ms'])
PY
