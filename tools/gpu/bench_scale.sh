#!/bin/bash
# bench.py at N = 2, 4, 8 ranks on one box (needs gpurun --gpus 8): weak-scaling headline + the C4 strong-scaling config
mkdir -p gpurun_out
for n in ${NS:-2 4 8}; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520 + n)) \
    bench.py --gpus $n --steps 5 --warmup 3 --no-nrsfm --no-stream --no-matching --no-cpu-baseline > gpurun_out/bench_r02_n$n.json 2> gpurun_out/bench_r02_n$n.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/bench_r02_n$n.json").read().strip().splitlines()[-1])
c4 = d["configs"]["C4"]
print("N=$n value %.0f e2e %.0f | C4 strong: %.2f ms, %.0f solves/s, e2e %.2f ms" % (d["value"], d["e2e"]["value"], c4["batch_latency_ms"], c4["value"], c4["e2e"]["batch_latency_ms"]))
PY
done
