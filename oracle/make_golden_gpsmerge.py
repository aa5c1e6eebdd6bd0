"""Test infrastructure (never imported by the product): golden files for the GPSmerge drop-in.

Builds a small synthetic KITTI ``oxts`` folder under tests/golden/oxts/ (12 frames of 30 columns, frame 5 missing, frame 7
with two rows, stamps with nanosecond digits) and runs the UNMODIFIED reference functions on it -- /root/reference/GPSmerge.py
imported with a stub for tkinter (its only missing import; the dialogs are not called) -- writing
tests/golden/oxts/expected_combined.txt and expected_timestamps.txt.  Run in the build container (the reference does not
travel to the GPU box):  python oracle/make_golden_gpsmerge.py
"""
import importlib.util
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden", "oxts")
TIME_OFFSET = 0.0375


def make_inputs():
    os.makedirs(os.path.join(OUT, "data"), exist_ok=True)
    rng = np.random.default_rng(404)
    t, lines = 40.354663127, []
    lat, lon, alt = 49.033603440345, 8.3950031909457, 112.83492279053
    for k in range(12):
        sec = t + 0.1 * k + rng.uniform(-0.004, 0.004)
        lines.append("2011-09-30 11:50:%012.9f" % sec)
        if k == 5:
            continue
        rows = 2 if k == 7 else 1
        with open(os.path.join(OUT, "data", f"{k:010d}.txt"), "w") as f:
            for _ in range(rows):
                lat += rng.normal(0, 2e-6); lon += rng.normal(0, 3e-6); alt += rng.normal(0, 0.02)
                motion = rng.normal(0, 1.0, 20)
                vals = [repr(float(lat)), repr(float(lon)), repr(float(alt))] + [repr(float(v)) for v in motion]
                vals += [repr(float(rng.uniform(0.02, 0.6))), repr(float(rng.uniform(0.01, 0.05))), "4", str(int(rng.integers(6, 12))), "5", "5", "6"]
                f.write(" ".join(vals) + ("\n" if k % 3 else ""))            # some frames end without a newline
    with open(os.path.join(OUT, "timestamps.txt"), "w") as f:
        f.write("\n".join(lines) + "\n")


def reference_module():
    tk = types.ModuleType("tkinter"); tk.filedialog = types.ModuleType("tkinter.filedialog"); tk.Tk = object
    sys.modules.setdefault("tkinter", tk); sys.modules.setdefault("tkinter.filedialog", tk.filedialog)
    spec = importlib.util.spec_from_file_location("ref_gpsmerge", "/root/reference/GPSmerge.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    make_inputs()
    ref = reference_module()
    stamps = ref.load_timestamps(os.path.join(OUT, "timestamps.txt"), TIME_OFFSET)
    with open(os.path.join(OUT, "expected_timestamps.txt"), "w") as f:
        f.write("\n".join(stamps) + "\n")
    ref.create_combined_file(stamps, os.path.join(OUT, "data"), os.path.join(OUT, "expected_combined.txt"))
    print(open(os.path.join(OUT, "expected_combined.txt")).read())
