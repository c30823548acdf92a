cd /root/repo
for w in 4 8 12 16 24 32; do echo "== waves $w"; FGC_ROWRED_WAVES=$w REPS=5 ONLY=chan_stats,cbn_act_fwd,cbn_act_bwd,minmax_fwd,minmax_bwd python scripts/prof_elem.py 2>&1 | tail -n 5; done
