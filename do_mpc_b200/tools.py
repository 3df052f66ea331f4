"""Small helpers of `do_mpc.tools` that the reference's example scripts use around the controller path
(reference do_mpc/tools/_timer.py: Timer.tic / toc / info / hist)."""
import time

import numpy as np


class Timer:
    def __init__(self, name="timer", unit="ms"):
        factors = {"ms": 1000.0, "s": 1.0, "min": 1.0 / 60.0, "h": 1.0 / 3600.0}
        if unit not in factors:
            raise ValueError(f"unit must be one of {list(factors)}")
        self.name, self.unit, self.factor = name, unit, factors[unit]
        self.t_start, self.t_list = None, []

    def tic(self):
        self.t_start = time.time()

    def toc(self):
        if self.t_start is None:
            raise Exception("Timer was not started (tic) before toc")
        self.t_list.append((time.time() - self.t_start) * self.factor)
        self.t_start = None

    def info(self):
        t = np.asarray(self.t_list)
        if t.size == 0:
            print(f"{self.name}: no samples")
            return
        print(f"Mean runtime {self.name}: {t.mean():.2f}{self.unit} +- {t.std():.2f}{self.unit} (n={t.size}), "
              f"max {t.max():.2f}{self.unit}, min {t.min():.2f}{self.unit}")

    def hist(self, *args, **kwargs):
        """Histogram counts of the recorded times (the reference draws them with matplotlib; plotting is out of scope)."""
        return np.histogram(np.asarray(self.t_list), *args, **kwargs)
