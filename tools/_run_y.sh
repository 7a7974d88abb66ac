cd $GRAFT_REPO_ROOT
GTTS_PROFILE=1 python tools/profile_run.py --mixed --utts 4144 --frames 100 --reps 2 --lib ab/prof.so 2>&1 | tail -4 | cut -c1-400
for v in own0 own2 own4 own7; do echo "== $v"; python tools/profile_run.py --mixed --utts 4144 --frames 100 --reps 3 --lib ab/$v.so | tail -2 | head -1; done
