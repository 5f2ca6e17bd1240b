#!/usr/bin/env bash
# Developer aid (GPU box): bench.py under a few scheduling variants; prints value / ms per run / e2e.
cd "$(dirname "$0")/.."
IFS=";" read -ra VS <<< "${VARIANTS:-MCM_DUAL=0;MCM_DUAL=1;MCM_FUSED=0}"
for v in "${VS[@]}"; do
  env $v python bench.py --steps 3 --warmup 3 > /tmp/bv.json 2> /tmp/bv.err || { echo "$v FAILED"; tail -3 /tmp/bv.err; continue; }
  python - "$v" <<'PY'
import json, sys
d = json.loads(open('/tmp/bv.json').read().strip().splitlines()[-1])
r = d["roofline"]
print(sys.argv[1], "| frames/s", round(d["value"]), "| ms/run", round(d["ms_per_step"], 1), "| e2e", round(d["e2e"]["value"]),
      "| tcgen05 ms", round(r["all_tcgen05_kernels_ms_per_run"], 1), "row ms", round(r["row_kernel_ms_per_run"], 1), "| frac", round(r["frac"], 3),
      "| clocks", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
PY
done
