#!/bin/bash
# One gpurun call that refreshes everything under profiles/ for the round:  profiles/final_round.sh
# (tests, smoke, bench, ncu launch list, ncu --set full of the dominant kernel, per-kernel rooflines, epilogue variants)
set -u
O=gpurun_out
mkdir -p $O
timeout 500 python -m pytest tests -q -m gpu > $O/final_pytest.log 2>&1; tail -2 $O/final_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/final_smoke.log 2>&1; tail -3 $O/final_smoke.log
timeout 300 python bench.py > $O/final_bench.json 2> $O/final_bench.err; cut -c1-300 $O/final_bench.json
timeout 300 python bench.py --workload cfg5 > $O/final_cfg5.json 2> $O/final_cfg5.err; cut -c1-300 $O/final_cfg5.json
timeout 300 python profiles/kernel_bench.py --out $O/final_kernels.jsonl > $O/final_kernels.log 2>&1; cut -c1-110 $O/final_kernels.jsonl
[ "${LIGHT:-0}" = 1 ] || { timeout 200 python profiles/k4_epilogue_sweep.py --variants=-1,0,4,304,303,305,104 --out $O/final_k4_epi.jsonl > $O/final_k4_epi.log 2>&1; cat $O/final_k4_epi.jsonl; }
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/final_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/final_ncu_bench.log 2>&1
[ "${LIGHT:-0}" = 1 ] && exit 0       # LIGHT=1: tests, smoke, bench lines, per-kernel rooflines and the launch list only
for v in 304; do timeout 120 profiles/ncu_k4_epi.sh $v; done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k4_score_bf16_cg2 -s 1 -c 1 -f -o $O/k4_r1_final2 \
  python profiles/k4_epilogue_sweep.py --variants=304 --iters 1 --out $O/tmp_epi.jsonl > $O/final_ncu_k4.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k3_gru_bf16_cluster -s 2 -c 1 -f -o $O/k3_r1_cluster \
  python profiles/kernel_bench.py --only k3 --out $O/tmp_k3.jsonl > $O/final_ncu_k3.log 2>&1
ls -la $O/*.ncu-rep | tail -3
