cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -q -k "model5" 2>&1 | tail -3
for v in m5c1 m5c2 m5c8; do echo "== $v"; python tools/profile_run.py --model5 --utts 1776 --frames 60 --reps 3 --lib ab/$v.so | tail -2 | head -1; python tools/profile_run.py --model5 --utts 1 --frames 332 --reps 3 --lib ab/$v.so | tail -2 | head -1; done
echo "== default (chunk 4)"
python tools/profile_run.py --model5 --utts 1776 --frames 60 --reps 3 | tail -2 | head -1
python tools/profile_run.py --model5 --utts 1 --frames 332 --reps 3 | tail -2 | head -1
