"""Wall-clock of the public API on one GPU next to the CPU oracle (BASELINE configs[0] / [4] shapes)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import guido_b200 as guido
from guido_b200 import _native as N
from guido_b200.mmej import generate_mmej_patterns_batch
from oracle import oracle as O

img = N.Image.synth([2_500_000] * 4, seed=1001, n_frac=0.005).upload(device=0)
seq = img.fetch(1, 1_000_000, 100_000)
loc = guido.Locus(sequence=seq, name="chr2", start=1_000_001)
loc.find_guides()
t0 = time.perf_counter(); guides = loc.find_guides(); t1 = time.perf_counter()
print(f"find_guides 100 kb: {len(guides)} guides in {1e3*(t1-t0):.1f} ms (GPU scan + Python records)")
t0 = time.perf_counter(); w = O.find_guides_records(seq, 1_000_001, None, chromosome="chr2"); t1 = time.perf_counter()
print(f"oracle (Python restatement of the reference loop, 1 core): {len(w)} guides in {1e3*(t1-t0):.1f} ms")
st = N.Stats()
fw, rv = N.pam_scan_seq(seq, "NGG", stats=st)
print(f"  of which the scan call: kernel {st.kernel_ms:.3f} ms, device total {st.total_ms:.3f} ms")

wins = [g.locus_seq for g in guides if "N" not in g.locus_seq]
cuts = [g.relative_cut_pos for g in guides if "N" not in g.locus_seq]
st = N.Stats()
N.mmej_batch(wins[:64], cuts[:64])
t0 = time.perf_counter(); tup, offs, nc = N.mmej_batch(wins, cuts, stats=st); t1 = time.perf_counter()
print(f"mmej_batch: {len(wins)} windows, {len(tup)} kept tuples, kernel {st.kernel_ms:.2f} ms, call {1e3*(t1-t0):.1f} ms "
      f"-> {len(wins)/(st.kernel_ms*1e-3):.0f} windows/s kernel, {int(nc.sum())} candidates")
t0 = time.perf_counter(); loc.simulate_end_joining(); t1 = time.perf_counter()
print(f"Locus.simulate_end_joining (kernel + dict building in Python): {1e3*(t1-t0):.1f} ms for {len(guides)} guides")
t0 = time.perf_counter()
for wv, c in zip(wins[:200], cuts[:200]):
    O.mmej_tuples(wv, c)
t1 = time.perf_counter()
print(f"oracle C mmej: {1e3*(t1-t0)/200:.3f} ms per window (1 core)")
