#!/bin/bash
# A/B the bench under option sets: tools/ab_bench.sh tag "opt1=v opt2=v" "opt3=v" ...   (one bench run per quoted set)
tag=$1; shift
i=0
for set in "$@"; do
  args=""
  for kv in $set; do args="$args --opt $kv"; done
  python bench.py --no-cpu --steps ${STEPS:-20} --warmup 5 --batch ${BATCH:-512} $args > gpurun_out/ab_${tag}_$i.json 2> gpurun_out/ab_${tag}_$i.err
  python - "$set" gpurun_out/ab_${tag}_$i.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print("%-48s value %8.0f  e2e %8.0f  ms %.3f  self_check %s" % (sys.argv[1] or "(defaults)", d["value"], d["e2e"]["value"], d["ms_per_step"], d.get("self_check_max_deg_vs_simt_path", d.get("config", {}).get("self_check_max_deg_vs_simt_path"))))
except Exception as e:
    print("%-48s FAILED %s" % (sys.argv[1], e))
PY
  i=$((i+1))
done
