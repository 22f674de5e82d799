"""Aggregate an ncu source page (cuda,sass view) per CUDA source line: instructions executed, thread-instructions, avg active threads, stall samples."""
import csv, sys, subprocess
rep, kern = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 45
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern,
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
lines = []
fname = ""
hdr = None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if len(r) > 8 and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr): continue
    if r[0] == "" : continue
    try:
        ln = int(r[0])
    except ValueError:
        continue
    g = lambda n: r[hdr.index(n)]
    try:
        lines.append((fname, ln, r[1].strip()[:110], int(g("Instructions Executed")), int(g("Thread Instructions Executed")), int(g("# Samples"))))
    except ValueError:
        pass
ti = sum(l[3] for l in lines); tt = sum(l[4] for l in lines); ts = sum(l[5] for l in lines)
print("kernel", rows[0][:2] if rows else None)
print("total warp-instr %d  thread-instr %d  avg active %.2f  samples %d" % (ti, tt, tt / max(ti, 1), ts))
lines.sort(key=lambda l: -l[3])
print("%-18s %5s %7s %7s %6s %7s  %s" % ("file", "line", "inst%", "thr%", "act", "smpl%", "source"))
for f, ln, src, ie, te, sm in lines[:topn]:
    print("%-18s %5d %6.2f%% %6.2f%% %6.1f %6.2f%%  %s" % (f, ln, 100.0 * ie / ti, 100.0 * te / tt, te / max(ie, 1), 100.0 * sm / max(ts, 1), src))
