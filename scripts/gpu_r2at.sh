#!/bin/bash
cd "$(dirname "$0")/.."
for d in 0 1 2 4 7; do echo "== FGC_SF_DBG=$d"; FGC_SF_DBG=$d timeout -k 10 120 python scripts/prof_small.py 2>&1 | head -4; done
for n in 2 3; do echo "== FGC_SWG_NKK=$n"; FGC_SWG_NKK=$n timeout -k 10 120 python scripts/prof_small.py 2>&1 | head -3; done
