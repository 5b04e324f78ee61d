# K1 geometry sweep (one gpurun call): correctness of every compiled variant on the small parity tests, then the
# dominant kernel's time on the bench workload per variant and per claim depth.
mkdir -p gpurun_out
out=gpurun_out/r2_k1_variants.txt
: > $out
for v in ${VARIANTS:-0 5 1 2 3 4}; do
  echo "== variant $v: parity" | tee -a $out
  SILO_K1_VARIANT=$v timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py -q -m gpu -x -k "not baseline_sizes" 2>&1 | tail -2 | tee -a $out
  echo "== variant $v: probe" | tee -a $out
  SILO_K1_VARIANT=$v timeout 300 python profiles/k1_probe.py 2>&1 | tail -4 | tee -a $out
done
for c in ${CLAIMS:-0 2}; do
  echo "== variant 0, claims in flight $c" | tee -a $out
  SILO_K1_CLAIMS=$c timeout 300 python profiles/k1_probe.py 2>&1 | tail -4 | tee -a $out
done
echo "== variant 0, stream only" | tee -a $out
SILO_K1_STREAM_ONLY=1 timeout 300 python profiles/k1_probe.py 2>&1 | tail -4 | tee -a $out
