#!/usr/bin/env bash
# usage: tools/bench_quick.sh [ENV=val ...]  -> one line: frames/s, dominant-kernel frac, whole-step frac
env "$@" timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra 2>/dev/null | python -c '
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(sys.argv[1:], round(d["value"]), round(d["roofline"]["frac"],4), round(d["roofline"]["whole_step_frac"],4))' "$@"
