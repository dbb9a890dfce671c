"""Throughput of the warp kernel vs resident warps per SM (MPCR_BLOCKS_PER_SM)."""
import os, sys, subprocess
B = sys.argv[1] if len(sys.argv) > 1 else "16384"
for n in (3, 4, 5, 6, 7):
    env = dict(os.environ, MPCR_BLOCKS_PER_SM=str(n))
    out = subprocess.run([sys.executable, "tools/prof_one.py", B, "warp", "3"], env=env, capture_output=True, text=True).stdout
    print(n, out.strip(), flush=True)
