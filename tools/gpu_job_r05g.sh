#!/bin/bash
# heads + loss kernel with the next env's rows prefetched
mkdir -p gpurun_out
for v in default ul_nopf ul_pf_mb4 default ul_nopf; do
  unset MARL_MAZE_LIB
  if [ $v != default ]; then export MARL_MAZE_LIB=$PWD/variants/$v.so; fi
  timeout 300 python tools/k5_bench.py --skip-update 2>&1 | grep '"rows"' | head -1 | python -c "
import json,sys,os
d=json.loads(sys.stdin.read()); print(json.dumps({'lib': os.path.basename(os.environ.get('MARL_MAZE_LIB','default')), 'rows': d.get('rows'), 'heads_loss_ms': d.get('heads_loss_ms')}))" | tee -a gpurun_out/r05g_heads_loss.jsonl
done
