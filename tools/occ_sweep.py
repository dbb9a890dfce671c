"""Throughput of the warp kernel vs resident warps per SM and L1/shared carve-out."""
import os, sys, subprocess
B = sys.argv[1] if len(sys.argv) > 1 else "16384"
for n, carve in ((7, 100), (6, 100), (6, 84), (5, 100), (5, 70), (4, 100), (4, 56), (3, 42)):
    env = dict(os.environ, MPCR_BLOCKS_PER_SM=str(n), MPCR_CARVEOUT=str(carve))
    out = subprocess.run([sys.executable, "tools/prof_one.py", B, "warp", "3"], env=env, capture_output=True, text=True).stdout
    print(n, carve, out.strip(), flush=True)
