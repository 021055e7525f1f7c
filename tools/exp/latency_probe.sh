#!/bin/bash
# the single-trajectory callback from C: completion word polled by the host (default) against stream synchronisation
R=oracle/_ref/params/
for ts in 100 200; do
  echo "== ts=$ts, completion word"; tools/exp/latency $R tempest S10 $ts
  echo "== ts=$ts, stream synchronisation"; TOLCUDA_LATENCY_POLL=0 tools/exp/latency $R tempest S10 $ts
done
echo "== G7 ts=100"; tools/exp/latency $R skywalker G7 100
