#!/bin/bash
# multi-GPU run (gpurun --gpus N): H2D probe, cfg4 exact mode over NCCL, bench e2e scaling with NUMA-local pinned buffers
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1; lscpu | grep -i "numa\|socket\|model name\|^CPU(s)" >> gpurun_out/topo_$N.txt; free -g | head -2 >> gpurun_out/topo_$N.txt
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
P=29500
for k in 1 2 4 8; do
  [ $k -le $N ] || continue
  P=$((P+1))
  timeout -s KILL 300 $TR --nproc-per-node $k --master-port $P tools/h2d_probe.py 2 6 > gpurun_out/h2d_probe_$k.json 2> gpurun_out/h2d_probe_$k.err; echo "probe $k exit $?"; tail -c 300 gpurun_out/h2d_probe_$k.json
done
P=$((P+1))
[ -n "$SKIP_CFG4" ] || timeout -s KILL 600 $TR --nproc-per-node $N --master-port $P tools/cfg4_state_modes.py 3600 4096 > gpurun_out/cfg4_${N}gpu.json 2> gpurun_out/cfg4_${N}gpu.err; echo "cfg4 exit $?"; tail -c 600 gpurun_out/cfg4_${N}gpu.err | tail -5
for k in ${BENCH_NS:-2 4 8}; do
  [ $k -le $N ] || continue
  P=$((P+1))
  timeout -s KILL 600 $TR --nproc-per-node $k --master-port $P bench.py --gpus $k --steps 3 --warmup 3 > gpurun_out/bench_${k}gpu.json 2> gpurun_out/bench_${k}gpu.err; echo "bench $k exit $?"
  P=$((P+1))
  AA_NO_NUMA_BIND=1 timeout -s KILL 600 $TR --nproc-per-node $k --master-port $P bench.py --gpus $k --steps 3 --warmup 3 > gpurun_out/bench_${k}gpu_nobind.json 2> gpurun_out/bench_${k}gpu_nobind.err; echo "bench nobind $k exit $?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_*gpu*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.1fM e2e %.2fM pcm16 %.2fM" % (d["value"]/1e6, d["e2e"]["value"]/1e6, d["e2e_pcm16"]["value"]/1e6))
    except Exception as e: print(f, "ERR", e)
PY
