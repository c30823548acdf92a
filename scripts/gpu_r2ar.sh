#!/bin/bash
cd "$(dirname "$0")/.."
for d in 0 1 2 4 8 12 15; do echo "== FGC_SWG_DBG=$d"; FGC_SWG_DBG=$d timeout -k 10 120 python scripts/prof_small.py 2>&1 | head -3; done
