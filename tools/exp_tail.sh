B="--no-e2e --no-cpu --steps 3 --warmup 3"
show() { grep '^{' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['value']/1e6,2), 'Mframes/s')"; }
for F in 0 4 2 6 1 9 15; do python bench.py $B --features $F 2>&1 | show "features=$F"; done
cp audio-analyzer-rs_b200/libaa_gpu.so /tmp/keep.so
cp variants/libaa_gpu_noscore.so audio-analyzer-rs_b200/libaa_gpu.so
for F in 1 15; do python bench.py $B --features $F 2>&1 | show "noscore features=$F"; done
cp /tmp/keep.so audio-analyzer-rs_b200/libaa_gpu.so
