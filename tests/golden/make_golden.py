"""Generate the golden fixtures from the oracle (run here, on CPU; commit the .npz files).

    python tests/golden/make_golden.py

There is no reference-side golden vector for this path (demo/tests.py:1-3 is empty, the arithmetic lives in
the absent third-party `vilbert` package), so these pin the ORACLE's outputs, not the reference's:
"parity unpinned" in the sense of the task statement.  Inputs and the 268 M-parameter checkpoint are
regenerated from seeds (oracle.make_inputs / oracle.init_weights), only the small head outputs are stored.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import vilbert_ref as R  # noqa: E402

CASES = [  # (B, Tin, V, seed, pad_regions)
    (1, 30, 36, 1234, 0),     # BASELINE.json configs[0]: single pair, 36 regions x 30 tokens
    (2, 37, 101, 2025, 0),    # the demo's shape: max_length 37 (worker.py:408), 100 boxes + global (worker.py:71, 433)
    (3, 16, 10, 31337, 2),    # odd batch (binary head falls back), short text, masked regions
]
KEEP = {0: "vil_prediction", 1: "vil_prediction_gqa", 2: "vil_logit", 3: "vil_binary_prediction",
        4: "vil_tri_prediction", 6: "vision_logit", 8: "linguisic_logit"}


def main():
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    model = R.build(seed=42)
    here = os.path.dirname(os.path.abspath(__file__))
    for B, Tin, V, seed, pad in CASES:
        inp = R.make_inputs(B, Tin, V, seed=seed, pad_regions=pad)
        out = model(*inp, compute_pretraining_heads=False)
        d = {name: out[i].numpy().astype(np.float32) for i, name in KEEP.items()}
        d.update(B=B, Tin=Tin, V=V, seed=seed, pad=pad, weight_seed=42)
        fn = os.path.join(here, f"full_B{B}_T{Tin}_V{V}.npz")
        np.savez_compressed(fn, **d)
        print(fn, {k: (v.shape, float(v.std())) for k, v in d.items() if hasattr(v, "shape") and v.ndim > 0})


if __name__ == "__main__":
    main()
