#!/usr/bin/env bash
# usage: tools/bench_wl.sh <workload> [ENV=val ...] -> frames/s of one workload (3 timed runs)
wl=$1; shift
env "$@" timeout 900 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline --no-extra 2>/dev/null | python -c '
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(sys.argv[1:], round(d["value"]), round(d["roofline"]["frac"],4), round(d["roofline"]["whole_step_frac"],4))' $wl "$@"
