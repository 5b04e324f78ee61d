# 1 -> 8 GPUs on one box (run under `gpurun --gpus 8`): the bench line at N = 1, 2, 4, 8 with the library's shard group
# (--reduce peer, the default) and, for comparison, the NCCL all-reduce path at the N listed in NCCL_AT (default: 8);
# the co-occurrence workload at N = 1 and 8. profiles/summarize_round.py reads gpurun_out/r2_scale_n*.json.
set -x
cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -3
python bench.py --skip-cpu-baseline > gpurun_out/r2_scale_n1.json 2> gpurun_out/r2_scale_n1.err
for n in 2 4 8; do
  reduces="peer"
  case " ${NCCL_AT:-8} " in *" $n "*) reduces="peer nccl";; esac
  for red in $reduces; do
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 3 --skip-cpu-baseline --reduce $red > gpurun_out/r2_scale_n${n}_${red}.json 2> gpurun_out/r2_scale_n${n}_${red}.err
    echo "n=$n $red rc=$?"
  done
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29528 bench.py --gpus 8 --workload cooc > gpurun_out/r2_bench_cooc_n8.json 2> gpurun_out/r2_bench_cooc_n8.err
echo "cooc n=8 rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_scale_n*.json')) + ['gpurun_out/r2_bench_cooc_n8.json']:
    try:
        line=[l for l in open(f) if l.startswith('{')][-1]
        j=json.loads(line)
        print(f, j['n_gpus'], 'ms/step', round(j['ms_per_step'],5), 'value %.3e'%j['value'], 'e2e ms', round(j['e2e']['ms_per_step'],5), 'e2e %.3e'%j['e2e']['value'])
    except Exception as e:
        print(f, 'ERR', e)
PY
