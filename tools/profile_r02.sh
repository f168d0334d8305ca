#!/bin/bash
# Round-2 profiling pass (one GPU box call): launch list of a short bench run + `ncu --set full`
# captures of the three heaviest kernels inside a 32-streams-per-launch batch.  Numbers printed by
# runs under ncu are never bench values; the plain run comes first.
mkdir -p gpurun_out
BENCH="python bench.py --no-cpu --no-e2e --no-others --steps 1 --warmup 1 --passes 1 --batch 32"
timeout 600 $BENCH > gpurun_out/r02_plain.json 2> gpurun_out/r02_plain.err || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_4k.csv $BENCH > gpurun_out/r02_ncu_list.log 2>&1
echo "launch list rc=$?"
for k in k_lf_sb k_inter_pred_sb; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:^$k -s 40 -c 1 -f -o gpurun_out/r02_full_$k $BENCH > gpurun_out/r02_ncu_$k.log 2>&1
  echo "$k rc=$?"
done
