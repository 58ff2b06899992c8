#!/bin/bash
# experiment: NTT kernel variants (tools/variants/libh2agg_*.so) at k = 18..24
P='
import sys,json
out=[]
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    if d.get("op") in ("ntt_fr","coset_ntt_fr"): out.append("%s k=%s %.3f %s" % (d["op"], d.get("log_n"), d.get("ms",0), [round(v,3) for v in d.get("phases_ms",{}).values()]))
print(" | ".join(out))
'
for v in "" "$@"; do
  lib=""; [ -n "$v" ] && lib=tools/variants/libh2agg_$v.so
  echo "== ${v:-default}"; H2A_LIB=$lib timeout 200 python tools/sweep.py --msm "" --ntt 18,20,22,24 2>&1 | python -c "$P"
done
