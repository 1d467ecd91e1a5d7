#!/bin/bash
# ncu dram throughput of the batched forward FFT at n = 4096 / 2048 / 1024 / 512 / 256 (round-1 verdict item 7); plain run first
mkdir -p gpurun_out
cat > /tmp/one_fft.py <<'PY'
import importlib, sys, torch
sys.path.insert(0, ".")
aa = importlib.import_module("audio-analyzer-rs_b200")
n = int(sys.argv[1]); batch = int(sys.argv[2])
x = torch.randn(batch, n, device="cuda"); out = torch.empty(batch, n // 2 + 1, 2, device="cuda")
f = aa.FftProcessor(n)
for _ in range(3):
    f.forward_device(x.data_ptr(), batch, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
print("ok", n, batch)
PY
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__inst_executed.avg.per_cycle_elapsed,launch__registers_per_thread,launch__block_size,launch__grid_size
for n in 4096 2048 1024 512 256; do
  B=$((1638400000/n/4))
  python /tmp/one_fft.py $n $B > /dev/null 2>&1 || { echo "plain $n failed"; continue; }
  ncu --metrics $M --clock-control none -k regex:fft_forward -s 2 -c 1 --csv --log-file gpurun_out/fft_ncu_$n.csv python /tmp/one_fft.py $n $B > /dev/null 2>&1
  echo "ncu $n exit $?"
done
