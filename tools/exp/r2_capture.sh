#!/bin/bash
# round-2 evidence run on one B200 (gpurun): ncu --set full captures of the bench kernel, odd-ts timings, the bench's
# launch list, host ceilings.  Numbers printed by commands under ncu are never bench values.
set -x
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -k regex:fg_cta -c 1"
python tools/kbench.py --batch 8192 --steps 20 > gpurun_out/r2_kb.txt 2>&1
$NCU -s 3 -f -o gpurun_out/r2_S10_B8192 python tools/kbench.py --batch 8192 --steps 2 > gpurun_out/ncu_a.log 2>&1
$NCU -s 2 -f -o gpurun_out/r2_S10_B65536 python tools/kbench.py --batch 65536 --steps 2 --warmup 2 > gpurun_out/ncu_b.log 2>&1
$NCU -s 3 -f -o gpurun_out/r2_G7_B4096 python tools/kbench.py --workload G7_skywalker_ts100 --batch 4096 --steps 2 > gpurun_out/ncu_c.log 2>&1
$NCU -s 3 -f -o gpurun_out/r2_S10_ts199_B8192 python tools/kbench.py --ts 199 --batch 8192 --steps 2 > gpurun_out/ncu_d.log 2>&1
rm -f gpurun_out/r2_oddts.txt
for t in 200 199 45 44 33 32; do for ov in 0 2; do python tools/kbench.py --ts $t --batch 65536 --steps 20 --overlap $ov >> gpurun_out/r2_oddts.txt 2>&1; done; done
for t in 100 99; do python tools/kbench.py --workload G7_skywalker_ts100 --ts $t --batch 65536 --steps 20 >> gpurun_out/r2_oddts.txt 2>&1; done
python bench.py --steps 5 --warmup 3 --e2e-steps 1 --no-cpu-baseline --no-ceiling > gpurun_out/r2_bench_short.json 2> gpurun_out/r2_bench_short.err && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_bench_launches.csv \
      python bench.py --steps 5 --warmup 3 --e2e-steps 1 --no-cpu-baseline --no-ceiling > gpurun_out/ncu_e.log 2>&1
rm -f gpurun_out/r2_hostceil_1gpu.txt
for th in 4 8 16; do tools/exp/hostceil --gpus 1 --threads $th >> gpurun_out/r2_hostceil_1gpu.txt; done
cat gpurun_out/r2_oddts.txt gpurun_out/r2_hostceil_1gpu.txt
ls -la gpurun_out/
