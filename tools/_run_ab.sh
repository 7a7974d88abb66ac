cd $GRAFT_REPO_ROOT
for v in walk1 walk2 walk1 walk2; do echo "== $v"; python tools/profile_run.py --utts 1036 --frames 200 --reps 4 --lib ab/$v.so | tail -3 | head -2; done
