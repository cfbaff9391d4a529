"""Generates tests/golden/traj_9_E0zfiF4DCt8.json from the reference's shipped CoTracker annotation
(/root/reference/dataset/VIPSeg/output_cotracker_all/9_E0zfiF4DCt8.json: 12 tracks x 46 frames, original-resolution
pixels).  The reference ships no frames, so the original frame size is an assumption recorded in the fixture.
Run here only (needs /root/reference):  python tests/golden/gen_traj_fixture.py"""
import json
import os

SRC = "/root/reference/dataset/VIPSeg/output_cotracker_all/9_E0zfiF4DCt8.json"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "traj_9_E0zfiF4DCt8.json")

with open(SRC) as f:
    tracks = json.load(f)
with open(OUT, "w") as f:
    json.dump({"source": "dataset/VIPSeg/output_cotracker_all/9_E0zfiF4DCt8.json of the reference (12 CoTracker tracks x 46 "
                         "frames, original-resolution pixels)", "assumed_original_size": [720, 1280], "tracks": tracks}, f)
print("wrote", OUT)
