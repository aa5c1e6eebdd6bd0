# -*- coding: utf-8 -*-
"""Drop-in for the reference's ``GPSmerge.py``: KITTI raw ``oxts`` folder -> the merged GNSS text file that
``EKFGPSSLAM.load_gps_data`` reads (pair B's ``combined_output.txt`` was made this way).

Same function names, arguments and output bytes as /root/reference/GPSmerge.py.  What runs where:

  * the per-frame files ``data/0000000000.txt ...`` (30 numbers each) are read as bytes, concatenated and parsed in ONE
    launch of the device text parser (``gsf_parse_table_dev``: correctly rounded decimal -> double, numpy.loadtxt rules), instead
    of one ``np.loadtxt`` call per frame (:42-50);
  * the timestamp arithmetic (:8-40) stays on the host in Python floats: it is a strictly sequential float recurrence
    (``t[i] = t[i-1] + (o[i] - o[i-1]) + offset`` -- the offset is added at EVERY step, a quirk the merged files carry) of a
    few thousand terms, and the output format is Python's ``f"{t:.18e}"`` / ``str(float)``;
  * there is no CPU fallback for the parsing: without the built library or without a B200 the call raises.

    python GPSmerge.py OXTS_FOLDER TIME_OFFSET [--out combined_output.txt]

replaces the tkinter folder dialog and the ``input()`` prompt (:66-81).
"""
from __future__ import annotations

import os
import sys
from datetime import datetime
from typing import List, Tuple

import numpy as np
import torch

from gps_optimize_slam_b200 import fusion

OXTS_COLUMNS = 30


def load_timestamps(timestamp_path: str, time_offset: float) -> List[str]:
    """GPSmerge.py:8-40.  ``timestamps.txt`` lines ``YYYY-mm-dd HH:MM:SS.fffffffff`` (first 26 characters used) -> strings
    ``"%.18e"`` of: first = time_offset, then previous + (difference of the original stamps) + time_offset."""
    original = []
    with open(timestamp_path, "r") as f:
        for line in f:
            stamp = datetime.strptime(line.strip()[:26], "%Y-%m-%d %H:%M:%S.%f")
            original.append((stamp - datetime(1970, 1, 1)).total_seconds())
    out = [time_offset]
    for i in range(1, len(original)):
        out.append(out[i - 1] + (original[i] - original[i - 1]) + time_offset)
    return [f"{t:.18e}" for t in out]


def _parse_frames(paths: List[str]) -> Tuple[np.ndarray, List[int]]:
    """All frame files in one device parse.  Returns (table [rows, cols], rows per file)."""
    chunks, rows_per_file = [], []
    for p in paths:
        raw = open(p, "rb").read()
        rows_per_file.append(sum(1 for ln in raw.split(b"\n") if ln.split(b"#")[0].strip()))
        chunks.append(raw if raw.endswith(b"\n") else raw + b"\n")
    blob = np.frombuffer(b"".join(chunks), dtype=np.uint8)
    if blob.size == 0 or sum(rows_per_file) == 0:
        return np.empty((0, 0)), rows_per_file
    table, status = fusion.parse_table(torch.from_numpy(blob.copy()).to("cuda"), 0, max_cols=OXTS_COLUMNS + 2)
    if status & 1:
        raise ValueError("could not convert string to float")
    if status & (2 | 16):
        raise ValueError("Wrong number of columns")
    if status & 4:
        raise ValueError("a field has more than 19 significant digits with an ambiguous rounding")
    table = table.cpu().numpy()
    if table.shape[0] != sum(rows_per_file):
        raise ValueError(f"oxts frames: {sum(rows_per_file)} data lines on disk, {table.shape[0]} rows parsed")
    return table, rows_per_file


def load_data_from_file(data_path: str):
    """GPSmerge.py:42-50: one frame file -> (first three columns [rows,3], numsats, velmode)."""
    table, _ = _parse_frames([data_path])
    return table[:, :3], int(table[0, 25]), int(table[0, 27])


def create_combined_file(timestamps: List[str], data_folder: str, output_file: str) -> None:
    """GPSmerge.py:53-64: one output line ``timestamp lat lon alt numsats velmode`` per row of every existing frame file
    ``{idx:010d}.txt``; numsats / velmode come from the file's first row; a missing file is reported and skipped."""
    present, paths = [], []
    for idx in range(len(timestamps)):
        p = os.path.join(data_folder, f"{idx:010d}.txt")
        if os.path.exists(p):
            present.append(idx); paths.append(p)
        else:
            print(f"warning: frame file not found: {p}")
    table, rows_per_file = _parse_frames(paths) if paths else (np.empty((0, 0)), [])
    with open(output_file, "w") as out:
        r = 0
        for idx, nrows in zip(present, rows_per_file):
            if nrows == 0:
                raise IndexError("index 0 is out of bounds for axis 0 with size 0")      # what data[0, 25] raises on an empty frame
            numsats, velmode = int(table[r, 25]), int(table[r, 27])
            for row in table[r:r + nrows, :3]:
                out.write(f"{timestamps[idx]} {' '.join(map(str, row))} {numsats} {velmode}\n")
            r += nrows


def main(argv=None) -> int:
    import argparse
    ap = argparse.ArgumentParser(description="KITTI oxts folder -> merged GNSS file (timestamp lat lon alt numsats velmode)")
    ap.add_argument("oxts_folder", help="folder holding timestamps.txt and data/")
    ap.add_argument("time_offset", type=float, help="time difference between the SLAM and the GPS sequence, seconds")
    ap.add_argument("--out", default="combined_output.txt")
    a = ap.parse_args(argv)
    timestamps_file = os.path.join(a.oxts_folder, "timestamps.txt")
    data_folder = os.path.join(a.oxts_folder, "data")
    if not os.path.exists(timestamps_file):
        print(f"timestamp file not found: {timestamps_file}")
        return 1
    if not os.path.exists(data_folder):
        print(f"data folder not found: {data_folder}")
        return 1
    create_combined_file(load_timestamps(timestamps_file, a.time_offset), data_folder, a.out)
    print(f"merged file written: {a.out}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
