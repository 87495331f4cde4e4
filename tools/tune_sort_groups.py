"""GPU: sweep the sort-bucket grouping (ISX_SORT_GY / ISX_SORT_GX, read once per process) for one workload."""
import os, subprocess, sys
n = sys.argv[1] if len(sys.argv) > 1 else "1024"
field = sys.argv[2] if len(sys.argv) > 2 else "torus"
for gy, gx in ((1, 1), (2, 1), (2, 2), (4, 2), (4, 4), (8, 4), (8, 8), (16, 8), (16, 16), (32, 16)):
    env = dict(os.environ, ISX_SORT_GY=str(gy), ISX_SORT_GX=str(gx))
    out = subprocess.run([sys.executable, "tools/detail_timing.py", n, field], env=env, capture_output=True, text=True).stdout
    keep = [l for l in out.splitlines() if "per call" in l or "k_seg_sort" in l or "k_scan_entries" in l or "radix_pass<false" in l]
    print(f"gy={gy} gx={gx}: " + " | ".join(" ".join(l.split()) for l in keep), flush=True)
