#!/bin/bash
# Multi-GPU evidence run of a round (under gpurun --gpus 8):  tools/final_n8.sh <tag>
tag=${1:-r2b}
out=gpurun_out
mkdir -p $out
nvidia-smi topo -m > $out/topo_n8_$tag.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_multirank.py -m gpu -q > $out/pytest_multirank_$tag.log 2>&1; echo "multirank rc=$?"; tail -2 $out/pytest_multirank_$tag.log
for n in 8 4 2; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
      bench.py --gpus $n 2> $out/bench_${tag}_n${n}_stderr.log | tail -1 > $out/bench_${tag}_n${n}.json
  echo "N=$n rc=$?"
  python -c "
import json; d=json.load(open('$out/bench_${tag}_n${n}.json')); r=d.get('rowblock') or {}
print('N', d['n_gpus'], 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'h2d GB/s/rank', round(d['e2e'].get('h2d_gbs', 0), 1), 'pts-only', round(d['e2e_points_only']['value']), 'rowblock ms', r.get('ms_per_pair'), 'n1', r.get('ms_per_pair_n1'), 'eff', r.get('eff_vs_n1'), 'T identical', r.get('T_identical_across_ranks'))"
done
