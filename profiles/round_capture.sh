# One GPU call that produces every artifact profiles/summarize_round.py reads (run under gpurun from the repo root):
#   bash profiles/round_capture.sh r1
tag=${1:-r1}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc $?" | tee -a gpurun_out/${tag}_pytest_gpu.log; tail -3 gpurun_out/${tag}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/${tag}_smoke.log
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -c 3000 gpurun_out/${tag}_bench.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err; tail -c 1500 gpurun_out/${tag}_bench_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 3 --warmup 3 --skip-cpu-baseline --eager > gpurun_out/${tag}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:containerAndCountKernel -s 4 -c 1 -f -o gpurun_out/${tag}_k1_full \
  python bench.py --steps 3 --warmup 3 --skip-cpu-baseline --eager > gpurun_out/${tag}_k1_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'evalProgramKernel|coverageDiffKernel|prepare|finalizeCountsKernel' -s 12 -c 4 -f -o gpurun_out/${tag}_other_full \
  python bench.py --steps 3 --warmup 3 --skip-cpu-baseline --eager > gpurun_out/${tag}_other_full.log 2>&1
ls -la gpurun_out | tail -20
if [ "$tag" != "r1" ]; then
  # round 2: the other workloads of BASELINE.json, the threshold sweep kernel, and compute-sanitizer over the parity suites
  for w in nof aa reads cooc; do
    python bench.py --workload $w > gpurun_out/${tag}_bench_${w}.json 2> gpurun_out/${tag}_bench_${w}.err; tail -c 600 gpurun_out/${tag}_bench_${w}.json
  done
  ncu --set full --clock-control none --import-source on -k regex:thresholdSweepKernel -s 1 -c 1 -f -o gpurun_out/${tag}_sweep_full \
    python profiles/config3_ncu_target.py > gpurun_out/${tag}_sweep_full.log 2>&1
  ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${tag}_config3_launches.csv \
    python profiles/config3_ncu_target.py > gpurun_out/${tag}_config3_launches.log 2>&1
  timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/${tag}_sanitizer_memcheck.log \
    python -m pytest tests/test_gpu_kernels.py -m gpu -x -q > gpurun_out/${tag}_sanitizer_memcheck_pytest.log 2>&1; echo "memcheck rc $?" | tee -a gpurun_out/${tag}_sanitizer_memcheck_pytest.log
  timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/${tag}_sanitizer_memcheck_parity.log \
    python -m pytest tests/test_gpu_parity.py -m gpu -q -k "not baseline and not twelve_genes and not tree_data and not concurrent and not shards_sum" > gpurun_out/${tag}_sanitizer_memcheck_parity_pytest.log 2>&1; echo "memcheck parity rc $?" | tee -a gpurun_out/${tag}_sanitizer_memcheck_parity_pytest.log
  timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file gpurun_out/${tag}_sanitizer_racecheck.log \
    python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "not baseline" > gpurun_out/${tag}_sanitizer_racecheck_pytest.log 2>&1; echo "racecheck rc $?" | tee -a gpurun_out/${tag}_sanitizer_racecheck_pytest.log
  tail -n 3 gpurun_out/${tag}_sanitizer_memcheck.log; tail -n 3 gpurun_out/${tag}_sanitizer_racecheck.log
  SILO_QUERY_TRACE=1 python profiles/query_trace_probe.py > gpurun_out/${tag}_query_trace.txt 2>&1; tail -n 4 gpurun_out/${tag}_query_trace.txt
fi
ls -la gpurun_out | tail -30
