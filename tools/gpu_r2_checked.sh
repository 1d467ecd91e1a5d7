#!/bin/bash
# Checked build of the analysis kernel (compute-sanitizer is closed on this pool): build the variants first with
#   tools/build_variants.sh "checked:AA_DEF_CHECKED=1" "checkfail:AA_DEF_CHECKED=2" "checkfail3:AA_DEF_CHECKED=3"
# then the whole GPU suite runs against the build whose kernel range-checks its data-dependent indices (any violated
# check fails the API call), and one test runs against the self-test build whose check is wrong on purpose.
mkdir -p gpurun_out
SO=audio-analyzer-rs_b200/libaa_gpu.so
cp $SO /tmp/keep.so
cp variants/libaa_gpu_checked.so $SO
timeout -s KILL 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --deselect tests/test_build_properties.py > gpurun_out/checked_pytest.log 2>&1
echo "checked build: pytest exit $?"; tail -3 gpurun_out/checked_pytest.log
timeout -s KILL 300 python tools/parity_soak.py --clips 6 > gpurun_out/checked_soak.json 2> gpurun_out/checked_soak.err; echo "checked build: soak exit $?"
timeout -s KILL 300 python bench.py --no-e2e --no-cpu --steps 2 --warmup 1 > gpurun_out/checked_bench.json 2> gpurun_out/checked_bench.err; echo "checked build: full-size bench exit $?"
cp variants/libaa_gpu_checkfail.so $SO
timeout -s KILL 300 python -m pytest tests/test_gpu_analyze.py -m gpu -q --tb=line -x -p no:cacheprovider -k cfg1_sine > gpurun_out/checkfail_pytest.log 2>&1
echo "self-test build (must FAIL): pytest exit $?"; grep -m2 "index check failed" gpurun_out/checkfail_pytest.log
if [ -f variants/libaa_gpu_checkfail3.so ]; then
  cp variants/libaa_gpu_checkfail3.so $SO
  timeout -s KILL 300 python -m pytest tests/test_gpu_analyze.py -m gpu -q --tb=line -x -p no:cacheprovider -k cfg1_sine > gpurun_out/checkfail3_pytest.log 2>&1
  echo "protocol self-test build (must FAIL with bit 14 = 0x4000): pytest exit $?"; grep -m1 "index check failed" gpurun_out/checkfail3_pytest.log
fi
cp /tmp/keep.so $SO
