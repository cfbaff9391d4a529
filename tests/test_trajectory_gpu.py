"""R1 parity (GPU): pt_rasterize_tracks against cv2 — bit-exact, integer work."""
import json
import random
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
cv2 = pytest.importorskip("cv2")
GOLDEN = Path(__file__).parent / "golden"


def test_real_file_bit_exact(cuda_dev):
    from oracle.trajectory import preprocess, trajectory_maps_cv2
    from posetraj_b200.trajectory import rasterize_tracks, rescale_tracks
    fx = json.loads((GOLDEN / "traj_9_E0zfiF4DCt8.json").read_text())
    tracks = rescale_tracks(fx["tracks"], [320, 576], fx["assumed_original_size"])
    for start in (0, 17, 32):
        want = trajectory_maps_cv2(tracks, 14, 320, 576, start=start)
        got = rasterize_tracks(tracks, 14, 320, 576, cuda_dev, output="u8", start=start).cpu().numpy()
        assert np.array_equal(got, want), start
        got_f = rasterize_tracks(tracks, 14, 320, 576, cuda_dev, output="f32", start=start).cpu().numpy()
        assert np.array_equal(got_f, preprocess(want))


@pytest.mark.parametrize("H,W,K,F,spread", [(320, 576, 40, 14, 0), (320, 576, 64, 25, 80), (64, 96, 30, 14, 60000),
                                            (40, 72, 200, 5, 10), (576, 1024, 16, 25, 300), (7, 9, 5, 3, 4)])
def test_random_tracks_bit_exact(cuda_dev, H, W, K, F, spread):
    """Random walks incl. overlapping tracks (painter's order), border crossings and far-outside (lost) points."""
    from oracle.trajectory import trajectory_maps_cv2
    from posetraj_b200.trajectory import rasterize_tracks
    rng = random.Random(H * 1000 + K)
    tracks = []
    for _ in range(K):
        x, y = rng.randint(-spread, W - 1 + spread), rng.randint(-spread, H - 1 + spread)
        pts = []
        for f in range(F):
            pts.append([x, y])
            step = rng.choice([0, 3, 12, 60])
            x += rng.randint(-step, step)
            y += rng.randint(-step, step)
            if spread > 1000 and rng.random() < 0.2:
                x, y = rng.randint(-spread, spread), rng.randint(-spread, spread)
        tracks.append(pts)
    want = trajectory_maps_cv2(tracks, F, H, W)
    got = rasterize_tracks(tracks, F, H, W, cuda_dev, output="u8").cpu().numpy()
    assert np.array_equal(got, want)


def test_edge_cases(cuda_dev):
    from posetraj_b200.trajectory import rasterize_tracks
    out = rasterize_tracks([], 14, 32, 48, cuda_dev, output="f32")          # no tracks: all black -> -1
    assert out.shape == (14, 3, 32, 48) and bool((out == -1).all())
    out = rasterize_tracks([[[5, 5]]], 1, 16, 16, cuda_dev, output="u8")    # a single frame: only the padding image
    assert out.shape == (1, 16, 16, 3) and not bool(out.any())
    with pytest.raises(ValueError):
        rasterize_tracks([[[1, 2]] * 3], 14, 32, 48, cuda_dev)               # too few points
    with pytest.raises(RuntimeError):
        rasterize_tracks([[[1, 2]] * 14], 14, 32, 48, "cpu")                 # no CPU fallback


def test_idempotent_and_deterministic(cuda_dev):
    from posetraj_b200.trajectory import rasterize_tracks
    rng = random.Random(3)
    tracks = [[[rng.randint(0, 95), rng.randint(0, 63)] for _ in range(14)] for _ in range(50)]
    a = rasterize_tracks(tracks, 14, 64, 96, cuda_dev, output="u8")
    b = rasterize_tracks(tracks, 14, 64, 96, cuda_dev, output="u8")
    assert torch.equal(a, b)


@pytest.mark.parametrize("K", [1, 2, 7])
def test_dataset_variant_channel_swaps(cuda_dev, K):
    """utils/dataset.py:762: cvtColor inside the track loop — colours depend on the number of tracks drawn afterwards."""
    from oracle.trajectory import trajectory_maps_cv2_dataset
    from posetraj_b200.trajectory import rasterize_tracks
    rng = random.Random(K)
    tracks = [[[rng.randint(0, 95), rng.randint(0, 63)] for _ in range(6)] for _ in range(K)]
    want = trajectory_maps_cv2_dataset(tracks, 6, 64, 96)
    got = rasterize_tracks(tracks, 6, 64, 96, cuda_dev, output="u8", style="dataset").cpu().numpy()
    assert np.array_equal(got, want)
