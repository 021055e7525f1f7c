#!/bin/bash
# round-2 evidence run on one B200 (gpurun): ncu --set full captures of the bench kernel (turned into text on the box:
# the reports themselves, with sources imported, exceed what gpurun copies back), odd-ts timings, the bench's launch
# list, host ceilings.  Numbers printed by commands under ncu are never bench values.
set -x
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -k regex:fg_cta -c 1"
cap() {  # name, skip, command...
  local name=$1 skip=$2; shift 2
  $NCU -s $skip -f -o /tmp/$name "$@" > gpurun_out/ncu_$name.log 2>&1
  python tools/ncu_summary.py /tmp/$name.ncu-rep > gpurun_out/$name.txt 2>&1
  python tools/ncu_stalls.py /tmp/$name.ncu-rep > gpurun_out/${name}_stalls.txt 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
  ls -la /tmp/$name.ncu-rep
}
python tools/kbench.py --batch 8192 --steps 20 > gpurun_out/r2_kb.txt 2>&1
cap r2_ncu_S10_B8192 3 python tools/kbench.py --batch 8192 --steps 2
cap r2_ncu_S10_B65536 2 python tools/kbench.py --batch 65536 --steps 2 --warmup 2
cap r2_ncu_G7_B4096 3 python tools/kbench.py --workload G7_skywalker_ts100 --batch 4096 --steps 2
cap r2_ncu_S10_ts199_B8192 3 python tools/kbench.py --ts 199 --batch 8192 --steps 2
# the traffic record bench.py reads, tied to the SASS of the library that was captured
cp /tmp/r2_ncu_S10_B65536.ncu-rep /tmp/r2_S10_B65536.ncu-rep
python tools/traffic_record.py S10_tempest_ts200_B65536 /tmp/r2_ncu_S10_B65536.ncu-rep 65536 > gpurun_out/traffic_S10.txt 2>&1
python tools/traffic_record.py G7_skywalker_ts100_B4096 /tmp/r2_ncu_G7_B4096.ncu-rep 4096 > gpurun_out/traffic_G7.txt 2>&1
cp profiles/roofline_traffic.json gpurun_out/roofline_traffic.json
rm -f gpurun_out/r2_oddts.txt
for t in 200 199 45 44 33 32; do for ov in 0 2; do python tools/kbench.py --ts $t --batch 65536 --steps 20 --overlap $ov >> gpurun_out/r2_oddts.txt 2>&1; done; done
for t in 100 99; do python tools/kbench.py --workload G7_skywalker_ts100 --ts $t --batch 65536 --steps 20 >> gpurun_out/r2_oddts.txt 2>&1; done
python bench.py --steps 5 --warmup 3 --e2e-steps 1 --no-cpu-baseline --no-ceiling > gpurun_out/r2_bench_short.json 2> gpurun_out/r2_bench_short.err && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_bench_launches.csv \
      python bench.py --steps 5 --warmup 3 --e2e-steps 1 --no-cpu-baseline --no-ceiling > gpurun_out/ncu_e.log 2>&1
rm -f gpurun_out/r2_hostceil_1gpu.txt
for th in 4 8 16; do tools/exp/hostceil --gpus 1 --threads $th >> gpurun_out/r2_hostceil_1gpu.txt; done
du -sh gpurun_out
