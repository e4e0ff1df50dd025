#!/bin/bash
# run under ONE ncu (3 metrics => single pass, caches left alone): steady-state DRAM bytes per kernel
# args: "<lib suffix>[:ENV=VAL]" ...
for spec in "$@"; do
  v=${spec%%:*}; envs=${spec#*:}; [ "$envs" = "$spec" ] && envs=""
  env PNS_NODE_THREADS=1 $envs PNS_B200_LIB=pednstream_b200/lib/libpns_b200_$v.so python bench.py --steps 12 --warmup 6 --no-cpu-baseline --no-env > /dev/null 2>> gpurun_out/dram_probe.err
done
