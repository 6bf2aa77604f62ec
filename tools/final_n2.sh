#!/bin/bash
# 2-GPU evidence run (under gpurun --gpus 2):  tools/final_n2.sh <tag>
tag=${1:-r2b}
out=gpurun_out
mkdir -p $out
timeout 300 python -m pytest tests/test_gpu_multirank.py -m gpu -q > $out/pytest_multirank_$tag.log 2>&1; echo "multirank rc=$?"; tail -2 $out/pytest_multirank_$tag.log
n=2
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29502 \
    bench.py --gpus $n 2> $out/bench_${tag}_n${n}_stderr.log | tail -1 > $out/bench_${tag}_n${n}.json
echo "N=$n rc=$?"
python -c "
import json; d=json.load(open('$out/bench_${tag}_n${n}.json')); r=d.get('rowblock') or {}
print('N', d['n_gpus'], 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'pts-only', round(d['e2e_points_only']['value']), 'rowblock ms', r.get('ms_per_pair'), 'n1', r.get('ms_per_pair_n1'), 'eff', r.get('eff_vs_n1'), 'T identical', r.get('T_identical_across_ranks'))"
