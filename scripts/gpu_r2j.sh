#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2j}
echo "=== ncu full: pair kernel"
FGC_H2_DBG=4 ONLY_FIRST=1 REPS=1 timeout -k 10 300 ncu --set full --clock-control none --import-source on -k "regex:conv_halo2_kernel" -c 2 -f -o gpurun_out/prof_halo2_$T python scripts/prof_conv.py > gpurun_out/ncu_halo2_$T.log 2>&1
echo "=== ncu full: single-CTA kernel"
FGC_HALO2=0 ONLY_FIRST=1 REPS=1 timeout -k 10 300 ncu --set full --clock-control none --import-source on -k "regex:conv_halo_kernel" -c 2 -f -o gpurun_out/prof_halo1_$T python scripts/prof_conv.py > gpurun_out/ncu_halo1_$T.log 2>&1
ls -la gpurun_out/*.ncu-rep
