V=build/var/libtolcuda_npp8_nbuf1_unit8_pad0.so
TOLCUDA_LIB=$V python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "single_trajectory or section_8d or batch_device or launch_shapes or compact_rows_path or summary" 2>&1 | tail -3
for lib in tol_b200/libtolcuda.so $V; do
  echo "== $lib"
  for args in "--workload G7_skywalker_ts100 --batch 4096 --steps 200" "--workload G7_skywalker_ts100 --batch 65536 --steps 20" "--workload S10_tempest_ts100 --batch 65536 --steps 20" "--ts 33 --batch 65536 --steps 20" "--ts 45 --batch 65536 --steps 20" "--ts 200 --batch 65536 --steps 20" "--ts 200 --batch 8192 --steps 50" "--ts 150 --batch 65536 --steps 20"; do
    for ov in 0 2; do TOLCUDA_LIB=$lib python tools/kbench.py $args --overlap $ov 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except: continue
    print('$args ov=$ov  ms %.4f frac %.4f'%(d['ms'],d['frac_of_6544']))"; done; done; done
