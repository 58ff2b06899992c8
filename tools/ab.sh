#!/bin/bash
# A/B: bench headline for the default build and each variant in tools/variants (usage: tools/ab.sh [variant names...])
run() { python bench.py --no-extras --no-prove --no-cpu-baseline --steps 10 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['phases_ms'].items()}, 'e2e', round(d['e2e']['ms_per_step'],3), 'batched', round(d['batched']['ms_per_step'],3))"; }
run default
for v in "$@"; do H2A_LIB=tools/variants/libh2agg_$v.so run $v; done
