cd $GRAFT_REPO_ROOT
for v in 1 2 4 13; do echo "== src1 unroll $v"; GTTS_LIB_PATH=ab/su$v.so python tools/config3_probe.py --utts 8192 2>&1 | tail -1; done
