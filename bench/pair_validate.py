"""Validation of the pair-row layer-1 forward kernel (HEFL_FWD_PAIR=1) against the default forward: the same
three-epoch training run, several seeds each, per-epoch loss / accuracy side by side."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from test_gpu_train import _setup

mode = os.environ.get("HEFL_FWD_PAIR", "1")
out = []
for seed in (5, 6, 7, 8):
    tr, pack, feed = _setup("tcgen05", "bf16", True, seed=seed)
    h = tr.fit(feed, None, 3, early_stopping=None, reduce_lr_patience=None)
    out.append([(round(s.loss, 4), round(s.accuracy, 3)) for s in h])
print("pair", mode, json.dumps(out))
