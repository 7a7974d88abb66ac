cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q -k "model5 or wide_batch" > gpurun_out/r02_tests_r.log 2>&1; tail -15 gpurun_out/r02_tests_r.log
python - <<'PY'
import time, numpy as np, sys
sys.path.insert(0, '.')
import gama_tts_b200 as g
from gama_tts_b200 import tracks as T
from gama_tts_b200.voices import default_voice5
import torch
s = g.TubeSynthesizer(0)
v = default_voice5("male")
for n_utt, n_frames in ((1, 332), (1184, 500), (2368, 500)):
    tracks = [T.synthetic_track(10 + (u % 64), n_frames) for u in range(n_utt)]
    frames, fo = g.pack_tracks(tracks)
    b = s.prepare5(v, fo)
    d_frames = torch.from_numpy(frames).cuda(); d_out = torch.empty(b.n_out_total, dtype=torch.float32, device="cuda")
    st = torch.cuda.current_stream()
    for _ in range(2): b.run_device(d_frames.data_ptr(), d_out.data_ptr(), st.cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(3): b.run_device(d_frames.data_ptr(), d_out.data_ptr(), st.cuda_stream)
    e1.record(st); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    audio = float(b.n_out.sum()) / 48000.0
    print("model5 utts %d frames %d: %.2f ms, %.0f audio-s/s, %.1f ns per internal sample per utterance" % (n_utt, n_frames, ms, audio / (ms * 1e-3), ms * 1e6 / float(b.n_internal.sum()) ))
PY
