#!/bin/bash
# A/B of a tracker change on configurations 3, 5a, 5b (and the headline): bench lines without the CPU legs.
# usage: tools/cfg_ab.sh <tag> [ENV=VALUE ...]
tag=$1; shift
for c in ${CFGS:-5a 5b 3 2}; do
  env "$@" python bench.py --config $c --steps 100 --warmup 10 --e2e-steps 8 --no-cpu-baseline --no-extra-configs > gpurun_out/ab_${tag}_$c.json 2> gpurun_out/ab_${tag}_$c.err
  python - <<P
import json
for l in open("gpurun_out/ab_${tag}_$c.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print("$tag", "$c", "value", round(d["value"]), "us_per_batch", round(d["us_per_batch"], 1), "kernels", [(k["kernel"], round(k["avg_launch_us"], 1)) for k in d.get("roofline_kernels", [])], d.get("tracker_stage_us"))
P
done
