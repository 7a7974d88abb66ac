#!/usr/bin/env python
"""The control-frame leg of bench.py alone (events_kernel: headline shape chained in front of the synthesis, and a batch
that fills the GPU with its HBM roofline), on a GPU box."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
import gama_tts_b200 as g  # noqa: E402

if __name__ == "__main__":
    print(json.dumps(bench.events_leg(g.TubeSynthesizer(0), 0), indent=1))
