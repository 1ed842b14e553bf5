#!/bin/bash
# usage: tools/sweep_env.sh "B C H W R dtype" "ENV1=a ENV2=b" "ENV1=c" ...   -- times one shape under several env settings
shape="$1"; shift
for cfg in "$@"; do
  echo "== $cfg"
  env $cfg python tools/time_shapes.py $shape 2>&1 | tail -1
  env $cfg NFPB200_BENCH_NO_HINT=1 python tools/time_shapes.py $shape 2>&1 | tail -1 | sed 's/^/   no-hint: /'
done
