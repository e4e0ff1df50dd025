#!/bin/bash
# Bench sweep over library variants (experiments only).  usage: variant_sweep.sh <tag> <lib suffixes...>
tag=$1; shift
B="python bench.py --steps 300 --warmup 10 --no-cpu-baseline --no-env"
for v in "$@"; do
  lib=pednstream_b200/lib/libpns_b200${v:+_$v}.so
  [ "$v" = "default" ] && lib=pednstream_b200/lib/libpns_b200.so
  PNS_B200_LIB=$lib $B > gpurun_out/${tag}_$v.json 2> gpurun_out/${tag}_$v.err
  python - "$v" gpurun_out/${tag}_$v.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print(sys.argv[1], round(d["ms_per_step"]*1e3,2), {k:round(v*1e3,2) for k,v in d["roofline"]["kernel_ms"].items()}, round(d["roofline"]["step"]["frac"],4), d["check"])
except Exception as e: print(sys.argv[1], "FAILED", e)
PY
done
