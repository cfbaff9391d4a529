"""R1 oracle pinning (CPU): the restated OpenCV rasterisation (oracle/trajectory.py) against cv2 itself — the
reference's rasteriser IS cv2 (scripts/run_inference_vipseg_json_repro.py:438-449)."""
import json
import random
from pathlib import Path

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
GOLDEN = Path(__file__).parent / "golden"


def test_thick_line_matches_cv2_everywhere():
    from oracle.trajectory import thick_line
    rng = random.Random(0)
    H, W = 48, 80
    for t in range(2500):
        m = rng.choice([0, 0, 6, 40, 60000])   # in-image, near the border, far outside (lost CoTracker tracks)
        p0 = (rng.randint(-m, W - 1 + m), rng.randint(-m, H - 1 + m))
        p1 = (rng.randint(-m, W - 1 + m), rng.randint(-m, H - 1 + m))
        if t % 10 == 0:
            p1 = p0
        if t % 7 == 0:
            p1 = (p0[0] + rng.randint(-3, 3), p0[1] + rng.randint(-3, 3))
        a = np.zeros((H, W, 3), np.uint8)
        b = np.zeros((H, W, 3), np.uint8)
        cv2.line(a, p0, p1, (0, 0, 255), 3)
        thick_line(b, p0, p1, (0, 0, 255), 3)
        assert np.array_equal(a, b), (p0, p1)


def test_circle_fill_matches_cv2():
    from oracle.trajectory import circle_fill
    rng = random.Random(1)
    H, W = 32, 40
    for _ in range(400):
        c = (rng.randint(-6, W + 6), rng.randint(-6, H + 6))
        a = np.zeros((H, W, 3), np.uint8)
        b = np.zeros((H, W, 3), np.uint8)
        cv2.circle(a, c, 3, (0, 255, 0), -1)
        circle_fill(b, c[0], c[1], 3, (0, 255, 0))
        assert np.array_equal(a, b), c


def test_real_tracks_restated_equals_reference_algorithm():
    """The shipped CoTracker file 9_E0zfiF4DCt8.json, rescaled like the inference script (:431), 14 frames."""
    from oracle.trajectory import trajectory_maps_cv2, trajectory_maps_restated
    from posetraj_b200.trajectory import rescale_tracks
    fx = json.loads((GOLDEN / "traj_9_E0zfiF4DCt8.json").read_text())
    tracks = rescale_tracks(fx["tracks"], [320, 576], fx["assumed_original_size"])
    want = trajectory_maps_cv2(tracks, 14, 320, 576)
    got = trajectory_maps_restated(tracks, 14, 320, 576)
    assert want[:13].any() and not want[13].any()          # 13 drawn maps + the black padding frame
    assert np.array_equal(want, got)
    # colours: red lines, green heads, nothing else
    assert set(map(tuple, want.reshape(-1, 3)[::7])) <= {(0, 0, 0), (255, 0, 0), (0, 255, 0)}


def test_rescale_styles():
    from posetraj_b200.trajectory import rescale_tracks
    js = {"0": [[1234, 807], [1138, 798]], "1": [[-48031.5, 20424.2], [3, 4]]}
    a = rescale_tracks(js, [320, 576], (720, 1280, 3))
    assert a[0][0] == [int(1234 * (576 / 1280)), int(807 * (320 / 720))] == [555, 358]
    assert a[1][0][0] == int(-48031.5 * (576 / 1280))      # int() truncates toward zero, also for lost tracks
    b = rescale_tracks(js, [320, 576], (720, 1280, 3), style="dataset")
    assert b[0][1] == [int(1138 / 1280 * 576), int(798 / 720 * 320)]
    with pytest.raises(ValueError):
        rescale_tracks(js, [320, 576], (720, 1280), style="nope")
