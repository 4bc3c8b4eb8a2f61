"""Recorded-sequence reader: the capture node's on-disk format
(data.csv + data/ folder) against what the reference's own reader parses from it, and the
decode -> pipeline -> panoramas path against the cv2 chain on the decoded frames."""
import json
import os
import shutil

import numpy as np
import pytest

from helpers import synthetic_chain
from multicamera_stitching_b200 import recorded, synthetic
from oracle import stitcher_ref

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_parser_matches_the_reference_reader(tmp_path):
    """tests/golden/recorded_reference.json = MediaPlayer/model.py data_reader on recorded_data.csv."""
    shutil.copy(os.path.join(GOLDEN, "recorded_data.csv"), tmp_path / "data.csv")
    with open(os.path.join(GOLDEN, "recorded_reference.json")) as f:
        ref = json.load(f)
    seq = recorded.RecordedSequence(str(tmp_path))
    assert seq.camera_labels == ref["camera_labels"]
    assert seq.timestamps == ref["timestamps"]
    assert seq.images == ref["images"]
    assert seq.line_count == ref["line_count"]
    assert seq.get_image(2, 1, 0) == ref["images"][0][1][2]
    assert seq.n_captures == 2 and seq.n_frames(0) == 4 and seq.n_frames(1) == 3
    assert seq.labels() == ["C", "LL", "RR"]


def _make_capture(path, n_cams, n_frames, h=90, w=160):
    sets = []
    for f in range(n_frames):
        fr = synthetic.make_frames(n_cams, h, w, 3, frame_index=f, kind="smooth")
        sets.append(fr)
    recorded.write_capture(str(path), sets, capture_id=0)
    return sets


def test_write_then_read_round_trip(tmp_path):
    import cv2
    sets = _make_capture(tmp_path, 3, 4)
    seq = recorded.RecordedSequence(str(tmp_path))
    labels = sorted(sets[0].keys())
    assert seq.labels() == labels
    assert seq.n_frames(0) == 4
    with open(tmp_path / "data.csv") as f:
        assert f.readline().strip() == "capture_id,timestamp,camera_label,image_file"
    fs = seq.load_frame_set(0, 2)
    for l in labels:
        assert np.array_equal(fs[l], cv2.imread(seq.image_path(0, l, 2)))
        assert fs[l].shape == sets[2][l].shape
    batch = seq.read_batch(0, 1, 4, workers=2)
    for l in labels:
        assert tuple(batch[l].shape) == (3,) + sets[0][l].shape
        assert np.array_equal(batch[l][1].numpy(), seq.load_frame_set(0, 2)[l])
    # a second capture appends
    recorded.write_capture(str(tmp_path), sets[:2], capture_id=1, timestamps=[5, 6])
    seq2 = recorded.RecordedSequence(str(tmp_path))
    assert seq2.n_captures == 2 and seq2.n_frames(1) == 2 and seq2.timestamps[1] == ["5", "6"]


def test_malformed_files_raise(tmp_path):
    (tmp_path / "data.csv").write_text("capture_id,timestamp,camera_label,image_file\n2,1,C,a.jpg\n")
    with pytest.raises(ValueError):
        recorded.RecordedSequence(str(tmp_path))
    (tmp_path / "data.csv").write_text("capture_id,timestamp,camera_label,image_file\n0,1,C\n")
    with pytest.raises(ValueError):
        recorded.RecordedSequence(str(tmp_path))
    (tmp_path / "data.csv").write_text("capture_id,timestamp,camera_label,image_file\n0,1,C,missing.jpg\n")
    with pytest.raises(IOError):
        recorded.RecordedSequence(str(tmp_path)).load_frame_set(0, 0)


@pytest.mark.gpu
def test_stitch_capture_matches_the_cv2_chain(cuda_device, tmp_path):
    n_cams, n_frames = 3, 7
    _make_capture(tmp_path, n_cams, n_frames)
    seq = recorded.RecordedSequence(str(tmp_path))
    st, states, labels, images = synthetic_chain(n_cams, 90, 160, 3, kind="smooth")
    first, panos = recorded.stitch_capture(st, seq, capture=0, device=cuda_device, chunk=2, depth=2, batch=3)
    assert first == 0 and panos.shape[0] == n_frames
    for f in range(n_frames):
        decoded = seq.load_frame_set(0, f)
        assert np.array_equal(panos[f].numpy(), stitcher_ref.stitch_chain(states, labels, decoded))
    # frame-range sharding: two ranks' shares tile the capture
    parts = [recorded.stitch_capture(st, seq, capture=0, device=cuda_device, chunk=2, rank=r, world_size=2)
             for r in range(2)]
    assert parts[0][0] == 0 and parts[1][0] == parts[0][1].shape[0]
    joined = np.concatenate([p[1].numpy() for p in parts])
    assert np.array_equal(joined, panos.numpy())
