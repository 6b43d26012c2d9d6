"""Loader of tests/golden/flower_golden.npz (made by scripts/make_golden_flowers.py from the unmodified reference)."""
import json
import os

import numpy as np

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "flower_golden.npz")


def cases():
    z = np.load(PATH)
    for name in json.loads(bytes(z["names"]).decode()):
        d = json.loads(bytes(z[name + "/flower"]).decode())
        fl = {"n_events": d["n_events"], "seqs": [s.encode() for s in d["seqs"]], "seq_event": d["seq_event"], "end_side": d["end_side"],
              "adj": [tuple(a) for a in d["adj"]]}
        yield name, fl, d["params"], z[name + "/stream"], z[name + "/bar"]
