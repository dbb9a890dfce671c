"""Convert the reference's track DATA (waypoint pickles, lane CSVs, max-error tables) into
.npz files under mpc_racing_b200/data/tracks/.  Data only - no reference code is copied.

    python tools/convert_tracks.py            (build container; reads /root/reference)

Keys per track: waypoints [n,2] (track id and z dropped, NO closing point added),
left / right [m,2] (lanes/<track>_left.csv, _right.csv as named on disk),
max_error [k,3] columns (ss, right, left) (lanes/<track>_max_error.csv).
"""
import os
import pickle

import numpy as np
import pandas as pd

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "mpc_racing_b200", "data", "tracks")
TRACKS = ["shanghai_intl_circuit", "t1_triple", "t2_triple", "t3", "t4"]

if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    for t in TRACKS:
        with open(os.path.join(REF, "waypoints", t), "rb") as f:
            w = pickle.load(f)
        wp = np.array([[p[0], p[1]] for p in w[1:]], dtype=np.float64)     # w[0] is the track id
        left = pd.read_csv(os.path.join(REF, "lanes", f"{t}_left.csv"))[["x", "y"]].to_numpy(np.float64)
        right = pd.read_csv(os.path.join(REF, "lanes", f"{t}_right.csv"))[["x", "y"]].to_numpy(np.float64)
        err = pd.read_csv(os.path.join(REF, "lanes", f"{t}_max_error.csv"))[["ss", "right", "left"]].to_numpy(np.float64)
        np.savez_compressed(os.path.join(OUT, t + ".npz"), waypoints=wp, left=left, right=right, max_error=err)
        print(t, wp.shape, left.shape, right.shape, err.shape)
