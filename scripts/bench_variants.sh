#!/bin/bash
# headline-only bench.py runs under different environment switches; one summary line per variant
#   bash scripts/bench_variants.sh "LOVR_LN_FOLD=0" "LOVR_LN_FOLD=1" ...
for v in "$@"; do
  env $v timeout 300 python bench.py --steps 6 --warmup 3 --no-scores --no-cpu-baseline 2>/dev/null | tail -1 > /tmp/bv.json
  python - "$v" <<'PY'
import json, sys
d = json.load(open("/tmp/bv.json"))
print(sys.argv[1], round(d["value"], 1), round(d["e2e"]["value"], 1), d["clocks"]["sm_mhz"],
      [(k["kernel"][:10], round(k["share_of_step"], 3), round(k["achieved"])) for k in d["roofline"]["kernels"]], flush=True)
PY
done
