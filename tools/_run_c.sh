cd $GRAFT_REPO_ROOT
L=gpurun_out/r02_prof_c.log
for k in v2 v1; do echo "== $k" >> $L; GTTS_KERNEL=$k timeout 300 python tools/profile_run.py --utts 1036 --frames 200 --reps 4 >> $L 2>&1; done
echo "== v2 role profile" >> $L; GTTS_PROFILE=1 timeout 300 python tools/profile_run.py --utts 1036 --frames 200 --reps 2 --lib ab/prof.so >> $L 2>&1
timeout 600 python -m pytest tests -m gpu -x -q -k "golden or ragged or stress or fresh or config2" >> $L 2>&1
cat $L
