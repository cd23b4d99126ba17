"""Loader for tests/golden/refbin_*.npz (written by oracle/refbin/make_golden.py).

A snapshot variable is stored under "<label>/<var>" only when it changed since the
previous label; `Golden.get(label, var)` walks back to the last stored value.
Static inputs live under "static/<var>".
"""
from pathlib import Path

import numpy as np

GOLD = Path(__file__).resolve().parent / "golden"


class Golden:
    def __init__(self, name):
        self.d = np.load(GOLD / f"refbin_{name}.npz")
        self.labels = [str(x) for x in self.d["labels"]]
        self.events = self.d["events"]

    def get(self, label, var):
        if f"static/{var}" in self.d:
            return self.d[f"static/{var}"]
        i = self.labels.index(label)
        while i >= 0:
            k = f"{self.labels[i]}/{var}"
            if k in self.d:
                return self.d[k]
            i -= 1
        raise KeyError((label, var))

    def snap(self, label):
        return _Snap(self, label)


class _Snap:
    def __init__(self, g, label):
        self.g, self.label = g, label

    def __contains__(self, var):
        try:
            self.g.get(self.label, var)
            return True
        except KeyError:
            return False

    def __getattr__(self, var):
        return self.g.get(self.label, var)

    def __getitem__(self, var):
        return self.g.get(self.label, var)
