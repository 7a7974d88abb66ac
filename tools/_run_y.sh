cd $GRAFT_REPO_ROOT
GTTS_PROFILE=1 python tools/profile_run.py --mixed --utts 4144 --frames 100 --reps 2 --lib ab/prof.so 2>&1 | tail -4 | cut -c1-420
