"""Formats either side of the hot path: .bin batch reader, sc_*.npy / pedestrians_*.json writers."""
import io
import json
import os

import numpy as np
import pytest

from conftest import SCAN_IDS


def test_npy_writer_is_byte_identical_to_np_save(tmp_path, golden_grids):
    from distropointclouds_b200 import io as dio
    for shape in [(3,), (160, 720), (1, 1, 1), (12345, 67), (0, 720)]:
        b = io.BytesIO()
        np.save(b, np.zeros(shape, np.float32))
        assert b.getvalue().startswith(dio.npy_header(shape)), shape
    g = golden_grids["000840"]
    np.save(tmp_path / "a.npy", g.astype(np.float32))            # kitti_dataset.py:474-475
    dio.save_sc_npy(str(tmp_path / "b.npy"), g.astype(np.float64))
    assert (tmp_path / "a.npy").read_bytes() == (tmp_path / "b.npy").read_bytes()
    paths = dio.save_sc_batch(str(tmp_path / "o"), ["000840", "x"], np.stack([g, g * 2]))
    assert [os.path.basename(p) for p in paths] == ["sc_000840.npy", "sc_x.npy"]
    assert np.array_equal(np.load(paths[1]), g * 2)


def test_bin_batch_reader(tmp_path, golden_scans):
    from distropointclouds_b200 import io as dio
    for sid in SCAN_IDS[:3]:
        golden_scans[sid].tofile(tmp_path / f"{sid}.bin")
    (tmp_path / "empty.bin").write_bytes(b"")
    (tmp_path / "notes.txt").write_text("ignored")
    names = dio.list_bin_files(str(tmp_path))
    assert names == sorted(f"{s}.bin" for s in SCAN_IDS[:3]) + ["empty.bin"]
    pts, offs = dio.read_bin_batch([str(tmp_path / n) for n in names])
    assert pts.dtype == np.float32 and pts.shape[1] == 4 and offs.dtype == np.int64
    assert offs.tolist() == np.concatenate([[0], np.cumsum([len(golden_scans[s]) for s in SCAN_IDS[:3]] + [0])]).tolist()
    for i, sid in enumerate(SCAN_IDS[:3]):
        assert np.array_equal(pts[offs[i]:offs[i + 1]], golden_scans[sid])
    (tmp_path / "bad.bin").write_bytes(b"\0" * 20)               # np.fromfile(...).reshape(-1, 4) raises too
    with pytest.raises(ValueError):
        dio.read_bin_batch([str(tmp_path / "bad.bin")])


def test_pedestrian_json_writer_matches_reference_files(tmp_path, golden_peds):
    from distropointclouds_b200 import io as dio
    dio.save_pedestrians_json(str(tmp_path / "p.json"), golden_peds["002073"])
    assert json.load(open(tmp_path / "p.json")) == golden_peds["002073"]
    golden = os.path.join(os.path.dirname(__file__), "golden", "pedestrians_002073.json")
    assert (tmp_path / "p.json").read_text() == open(golden).read()           # json.dump(..., indent=2)


@pytest.mark.gpu
def test_process_bin_dir_reproduces_golden_npy(tmp_path, golden_scans, golden_grids):
    from distropointclouds_b200 import io as dio
    src = tmp_path / "bins"
    src.mkdir()
    for sid in SCAN_IDS:
        golden_scans[sid].tofile(src / f"{sid}.bin")
    out = dio.process_bin_dir(str(src), str(tmp_path / "out"), batch_bytes=4 << 20)   # forces two batches
    assert [os.path.basename(p) for p in out] == [f"sc_{s}.npy" for s in SCAN_IDS]
    for sid, p in zip(SCAN_IDS, out):
        assert np.array_equal(np.load(p), golden_grids[sid])
    # pinned batch reader
    from distropointclouds_b200.engine import default_engine
    pts, offs = dio.read_bin_batch([str(src / f"{s}.bin") for s in SCAN_IDS[:2]], default_engine(0))
    assert np.array_equal(pts[offs[1]:offs[2]], golden_scans[SCAN_IDS[1]])
