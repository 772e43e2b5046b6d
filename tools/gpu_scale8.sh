#!/bin/bash
# N-GPU strong-scaling record: our arm and the reference arm as the driver launches them
N=${N:-8}
mkdir -p gpurun_out/scale
O=gpurun_out/scale
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > $O/bench_${N}gpu.json 2> $O/bench_${N}gpu.err; echo "bench N=$N exit $?"; tail -c 400 $O/bench_${N}gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 1 --warmup 1 > $O/bench_ref_${N}gpu.json 2> $O/bench_ref_${N}gpu.err; echo "ref N=$N exit $?"; cut -c1-200 $O/bench_ref_${N}gpu.json
python - <<PY
import json
d = json.loads([l for l in open('$O/bench_${N}gpu.json').read().strip().splitlines() if l.startswith('{')][-1])
print({k: d.get(k) for k in ('value','n_gpus','ms_per_step','scaling','verified','sharded_parity','collectives_ms_per_step','clocks','kmeans_leg','search_ms_per_step')})
print('e2e', d.get('e2e'))
PY
