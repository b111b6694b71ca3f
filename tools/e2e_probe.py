"""where does the host-buffer off-target call spend its time? (one GPU)"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import bench
from guido_b200 import _native as N
lens, seed, nf, _ = bench.genome_lens(sys.argv[1] if len(sys.argv) > 1 else "human")
n = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
img = N.Image.synth(lens, seed=seed, n_frac=nf).upload(device=0)
g = bench.sample_guides(img, n, seed + 1)
for rep in range(4):
    st = N.Stats()
    t0 = time.perf_counter()
    h, _ = img.offtarget_search(g, want_raw_flags=False, stats=st, pinned=True)
    t1 = time.perf_counter()
    print(f"rep {rep}: wall {1e3*(t1-t0):.1f} ms  device total {st.total_ms:.1f} ms  kernels {st.kernel_ms:.1f} ms  hits {len(h)} launches {st.n_launches}")
