cd $GRAFT_REPO_ROOT
for v in base rt1 rt2 rt3 an60 ansc base; do echo "== $v"; python tools/profile_run.py --utts 1036 --frames 200 --reps 4 --lib ab/$v.so | tail -3 | head -2; done
