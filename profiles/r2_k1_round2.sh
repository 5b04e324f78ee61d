mkdir -p gpurun_out
out=gpurun_out/r2_k1_round2.txt
: > $out
for v in ${VARIANTS:-0 5 1}; do
  echo "== variant $v: parity" | tee -a $out
  SILO_K1_VARIANT=$v timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py -q -m gpu -x -k "not baseline_sizes" 2>&1 | tail -2 | tee -a $out
  echo "== variant $v: probe" | tee -a $out
  SILO_K1_VARIANT=$v timeout 300 python profiles/k1_probe.py 2>&1 | tail -4 | tee -a $out
done
echo "== variant 0, claims 0" | tee -a $out
SILO_K1_CLAIMS=0 timeout 300 python profiles/k1_probe.py 2>&1 | tail -4 | tee -a $out
echo "== variant 0, debug times" | tee -a $out
SILO_K1_DEBUG=1 timeout 300 python profiles/k1_probe.py 2>&1 | grep -v "^$" | tail -45 | cut -c1-400 | tee -a $out
echo "== ncu" | tee -a $out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:containerAndCountKernel -s 4 -c 1 -f -o gpurun_out/r2_k1_full \
  python bench.py --steps 3 --warmup 3 --skip-cpu-baseline --eager > gpurun_out/r2_k1_full.log 2>&1; tail -3 gpurun_out/r2_k1_full.log | cut -c1-300
