#!/bin/bash
# frame time of config 2 on historical commits (worktrees under _bisect/, built in the container) and on the current variants
mkdir -p gpurun_out
for d in _bisect/*/; do
  c=$(basename $d)
  echo "== commit $c"
  (cd $d && timeout 300 python tools/sweep.py --persistent 0 --frames 20 --out $GRAFT_REPO_ROOT/gpurun_out/bisect_$c.json 2>&1 | cut -c1-230)
done
for lib in libyv_b200.so libyv_b200_imad.so; do
  echo "== $lib"
  YV_B200_LIB=$lib timeout 300 python tools/sweep.py --persistent 0 --frames 20 --out gpurun_out/sweep_${lib%.so}_primary.json 2>&1 | cut -c1-230
done
