"""Loader for the golden fixtures written by tests/golden/make_golden.py."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


class Case(dict):
    """info dict + arrays (attribute access for arrays via .arr[name])."""

    def __init__(self, info, arr):
        super().__init__(info)
        self.arr = arr

    def sub(self, prefix):
        """Arrays whose name starts with `prefix`, prefix stripped."""
        return {k[len(prefix):]: v for k, v in self.arr.items() if k.startswith(prefix)}


def load(name):
    z = np.load(os.path.join(HERE, "golden", name + ".npz"), allow_pickle=False)
    meta = json.loads(str(z["meta"]))
    per = [dict() for _ in meta]
    for k in z.files:
        if k == "meta":
            continue
        ci, nm = k.split("__", 1)
        per[int(ci[1:])][nm] = z[k]
    return [Case(info, arr) for info, arr in zip(meta, per)]


def cases(name, **filt):
    out = []
    for c in load(name):
        if all(c.get(k) == v for k, v in filt.items()):
            out.append(c)
    return out
