#!/bin/bash
# experiment: small MSMs (2^16, 2^18) over table width x reduction segment length; NTT tile sizes
P='
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print(d.get("op"), d.get("log_n"), d.get("window"), round(d.get("ms",0),3), {k:round(v,3) for k,v in d.get("phases_ms",{}).items()})
'
for t in 10 11; do echo "== ntt tile=$t"; H2A_NTT_LOG_TILE=$t timeout 200 python tools/sweep.py --msm "" --ntt 16,18,20,22,24 2>&1 | python -c "$P"; done
for seg in 2 4 8 16; do for pre in 14 16 18 20; do echo "== seg=$seg pre=$pre"; H2A_MSM_SEG=$seg timeout 100 python tools/sweep.py --msm 16,18 --ntt "" --precompute $pre 2>&1 | python -c "$P" | grep precomputed; done; done
