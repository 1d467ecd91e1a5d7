#!/bin/bash
# multi-GPU bench of the shipped build (gpurun --gpus N): one bench.py run at N ranks
N=${1:-2}
mkdir -p gpurun_out
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $N --master-port 29533 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_${N}gpu_final.json 2> gpurun_out/bench_${N}gpu_final.err; echo "bench $N exit $?"
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${N}gpu_final.json").read().strip().splitlines()[-1])
print("value %.1fM e2e %.2fM frac %s pcm16 %.2fM spectra %.2fM" % (d["value"]/1e6, d["e2e"]["value"]/1e6, d["e2e"].get("frac_of_h2d_ceiling"), d["e2e_pcm16"]["value"]/1e6, d["e2e_spectra"]["value"]/1e6))
PY
