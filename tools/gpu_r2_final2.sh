#!/bin/bash
# round 2, final pass over the shipped build: full GPU suite, smoke, bench (both arms), ncu (launch list + full captures),
# conditioning chain, parity soak
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/final_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/final_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/final_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/final_smoke.log
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference > gpurun_out/bench_final_reference.json 2> gpurun_out/bench_final_reference.err; echo "reference exit $?"
B="--no-e2e --no-cpu --steps 1 --warmup 3"
ncu --set full --clock-control none --import-source on -k regex:analyze_kernel -s 3 -c 1 -f -o gpurun_out/r02_v19_4096 python bench.py $B > gpurun_out/ncu_v19_4096.log 2>&1; echo "ncu 4096 exit $?"
B2="$B --n 2048 --sr 44100 --seconds 10 --clips 4096"
timeout 300 python bench.py $B2 > gpurun_out/bench_final_2048.json 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:analyze_kernel -s 3 -c 1 -f -o gpurun_out/r02_v19_2048 python bench.py $B2 > gpurun_out/ncu_v19_2048.log 2>&1; echo "ncu 2048 exit $?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/bench_launches_v19.csv python bench.py --no-cpu --steps 2 --warmup 3 > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches exit $?"
timeout 600 python tools/bench_cond.py --reps 3 > gpurun_out/bench_cond_final.json 2> gpurun_out/cond.err; echo "cond exit $?"
ncu --set full --clock-control none --import-source on -k regex:cond_cluster -c 1 -f -o gpurun_out/cond_cluster_final python tools/bench_cond.py --clips 256 --seconds 4 --reps 1 > gpurun_out/ncu_cond.log 2>&1; echo "ncu cond exit $?"
timeout 900 python tools/parity_soak.py --clips 24 > gpurun_out/soak_final.json 2> gpurun_out/soak.err; echo "soak exit $?"
