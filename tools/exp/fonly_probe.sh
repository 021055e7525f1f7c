#!/bin/bash
# F-only launches (needG = 0): the dedicated 64-register flavour against the plain flavour (variant built with
# tools/exp/build_variant.sh 8 1 8 0 -DTOLCUDA_NO_FONLY), plus the parity tests that exercise F-only calls
V=build/var/libtolcuda_npp8_nbuf1_unit8_pad0.so
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "section_8d or snopta or config2 or single_trajectory or batch_device or degenerate or long_traj" 2>&1 | tail -3
for lib in $V tol_b200/libtolcuda.so; do
  echo "== $lib"
  for args in "--batch 65536 --steps 30" "--batch 8192 --steps 100" "--workload G7_skywalker_ts100 --batch 65536 --steps 30" "--workload G7_skywalker_ts100 --batch 4096 --steps 200" "--ts 45 --batch 65536 --steps 30"; do
    for ov in 0 2; do TOLCUDA_LIB=$lib python tools/kbench.py $args --need F --overlap $ov 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except: continue
    print('$args ov=$ov  ms %.4f  %.0f GB/s  frac %.4f  %.3e node-evals/s'%(d['ms'],d['GBps'],d['frac_of_6544'],d['node_evals_per_s']))"; done; done; done
