#!/bin/bash
# Where the launch time goes, by executed action type: bench.py with action types rewritten to the no-op
# (CYG_BENCH_EXCLUDE=mode:type,...; diagnostics).  usage (GPU box): bash profiles/exclude_sweep.sh <out.log>
out=${1:-gpurun_out/exclude_sweep.log}
: > $out
run() {
  CYG_BENCH_EXCLUDE=$2 python bench.py --steps 240 --warmup 6 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read())
print('%-34s fused %.2f us/step   single %.2f us/launch' % ('$1', l['ms_per_step']*1e3, l['single_step_launch']['launch_us']))" >> $out
}
run "all types" ""
run "no block/unblock" "0:6,0:9"
run "no clean/revert/upgrade" "0:1,0:3,0:4"
run "no attack" "1:1"
run "no flips, deposits, attack" "0:6,0:9,0:1,0:3,0:4,1:1"
run "all no-op" "0:0,0:1,0:2,0:3,0:4,0:5,0:6,0:7,0:9,0:11,0:12,0:13,1:0,1:1,1:2,1:4"
cat $out
