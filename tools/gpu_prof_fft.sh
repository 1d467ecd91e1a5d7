#!/bin/bash
# ncu --set full of the batched forward FFT at n = 4096 and 2048
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
PY
for n in 4096 2048; do
  ncu --set full --clock-control none --import-source on -k regex:fft_forward -s 2 -c 1 -f -o gpurun_out/fft$n python /tmp/one_fft.py $n $((1600000000/n/4)) > gpurun_out/ncu_fft$n.log 2>&1
  echo "ncu $n exit $?"
done
