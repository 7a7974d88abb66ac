cd $GRAFT_REPO_ROOT
L=gpurun_out/r02_prof_g.log
for m in 95 94 93 91 87 79 31 0 32; do
  echo "== skip $m" >> $L
  GTTS_DEBUG_SKIP=$m GTTS_PROFILE=1 python tools/profile_run.py --utts 1036 --frames 100 --reps 2 --lib ab/profx.so 2>&1 | tail -4 >> $L
done
cat $L
