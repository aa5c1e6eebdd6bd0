"""GPSmerge drop-in against golden files made by the unmodified reference (oracle/make_golden_gpsmerge.py)."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OXTS = os.path.join(ROOT, "tests", "golden", "oxts")
TIME_OFFSET = 0.0375


def test_load_timestamps_matches_reference():
    """GPSmerge.py:8-40: 26-character truncation, offset added at every step, "%.18e" strings (host arithmetic)."""
    import GPSmerge
    got = GPSmerge.load_timestamps(os.path.join(OXTS, "timestamps.txt"), TIME_OFFSET)
    want = open(os.path.join(OXTS, "expected_timestamps.txt")).read().split("\n")[:-1]
    assert got == want and len(got) == 12


@pytest.mark.gpu
def test_create_combined_file_matches_reference_bytes(tmp_path):
    """12 frames (frame 5 missing, frame 7 with two rows, some files without a final newline) parsed in one device launch:
    the merged file equals the reference's byte for byte."""
    import GPSmerge
    stamps = GPSmerge.load_timestamps(os.path.join(OXTS, "timestamps.txt"), TIME_OFFSET)
    out = tmp_path / "combined.txt"
    GPSmerge.create_combined_file(stamps, os.path.join(OXTS, "data"), str(out))
    assert out.read_bytes() == open(os.path.join(OXTS, "expected_combined.txt"), "rb").read()
    xyz, numsats, velmode = GPSmerge.load_data_from_file(os.path.join(OXTS, "data", "0000000007.txt"))
    assert xyz.shape == (2, 3) and (numsats, velmode) == (4, 5)
    assert GPSmerge.main([OXTS, str(TIME_OFFSET), "--out", str(tmp_path / "cli.txt")]) == 0
    assert (tmp_path / "cli.txt").read_bytes() == out.read_bytes()


@pytest.mark.gpu
def test_merged_file_feeds_the_gnss_loader(tmp_path):
    """The merged file is what EKFGPSSLAM.load_gps_data reads (pair B's combined_output.txt is such a file)."""
    import EKFGPSSLAM
    import GPSmerge
    stamps = GPSmerge.load_timestamps(os.path.join(OXTS, "timestamps.txt"), TIME_OFFSET)
    out = tmp_path / "combined.txt"
    GPSmerge.create_combined_file(stamps, os.path.join(OXTS, "data"), str(out))
    gps = EKFGPSSLAM.load_gps_data(str(out), data_label="oxts", filter_config_override={"enabled": False})
    assert gps["positions"].shape == (12, 3) and gps["utm_zone"] == "32N"
