#!/bin/bash
# Final single-GPU evidence run of a round (under gpurun):  tools/final_n1.sh <tag>   -> gpurun_out/*_<tag>.*
tag=${1:-r2b}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -q > $out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?" 
python bench.py 2> $out/bench_${tag}_stderr.log | tail -1 > $out/bench_${tag}_n1.json; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > $out/bench_${tag}_reference_arm.json
for w in c3 c5 c4; do python tools/bench_configs.py --workload $w 2>/dev/null | tail -1; done > $out/secondary_workloads_$tag.jsonl
python tools/time_topk.py > $out/topk_$tag.txt 2>&1
timeout 200 python tools/fuzz_parity.py --seconds 100 > $out/fuzz_$tag.txt 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $out/launches_$tag.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/ncu_launches_$tag.log 2>&1
python tools/topk_once.py 32 > /dev/null 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:"match_tc_filter_kernel|topk_" -c 5 -f -o $out/topk_$tag \
      python tools/topk_once.py 32 > $out/ncu_topk_$tag.log 2>&1
tail -3 $out/pytest_gpu_$tag.log
python -c "
import json; d=json.load(open('$out/bench_${tag}_n1.json')); print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'pts-only', d['e2e_points_only']['value'], 'filter ms', d['roofline']['ms_per_launch'], 'frac', d['roofline']['frac'], 'knn ms', d['roofline_knn']['ms_per_step'], d['clocks'])"
cat $out/secondary_workloads_$tag.jsonl | cut -c1-400
tail -2 $out/fuzz_$tag.txt
