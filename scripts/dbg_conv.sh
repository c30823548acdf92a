#!/bin/bash
cd "$(dirname "$0")/.."
for i in 0 1 2 3 4 5 6 7 8 9; do
    s=$(date +%s)
    timeout -k 5 40 python -m pytest -q -m gpu -p no:cacheprovider tests/test_ops_gpu.py -k "test_conv_fwd and ${i}-tcgen05" > /tmp/dbg_$i.log 2>&1
    rc=$?
    e=$(date +%s)
    echo "case $i rc=$rc time=$((e - s)) : $(tail -n 1 /tmp/dbg_$i.log)"
done
