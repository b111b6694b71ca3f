"""ncu_lines.py <report.ncu-rep> [min_pct] -- warp instructions and stall samples per CUDA source line
(the cuda,sass correlated view of one capture, summed per line; needs -lineinfo and --import-source on)"""
import csv, io, subprocess, sys
rep = sys.argv[1]
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.7
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
fname, hdr, lines = None, None, []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr) and r[2] == "-":   # a source line (its SASS rows follow with an address)
        lines.append((fname, r))
ix_inst = hdr.index("Instructions Executed")
ix_samp = hdr.index("# Samples")
ix_thr = hdr.index("Avg. Threads Executed")
stall_ix = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[ix_inst]) for _, r in lines)
tot_s = sum(int(r[ix_samp]) for _, r in lines)
print(f"total warp instr {tot}  samples {tot_s}")
agg = {}
for h_i, h in stall_ix:
    agg[h] = sum(int(r[h_i] or 0) for _, r in lines)
print("stalls:", ", ".join(f"{k[6:]} {100 * v / max(tot_s, 1):.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > 0.01 * tot_s))
for f, r in lines:
    n, s = int(r[ix_inst]), int(r[ix_samp])
    if 100 * n / max(tot, 1) >= thr or 100 * s / max(tot_s, 1) >= thr:
        top = sorted(((int(r[i] or 0), h[6:]) for i, h in stall_ix), reverse=True)[:2]
        print(f"{f}:{r[0]:>4} inst {100 * n / tot:5.1f}% smp {100 * s / max(tot_s, 1):5.1f}% thr {r[ix_thr]:>5} "
              f"[{top[0][1]} {top[0][0]}, {top[1][1]} {top[1][0]}]  {r[1].strip()[:100]}")
