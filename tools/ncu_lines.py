"""Aggregate an ncu source page (cuda,sass) by CUDA source line: executed instructions and stall samples."""
import csv, subprocess, sys
rep, pat = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "k_align")
top = int(sys.argv[3]) if len(sys.argv) > 3 else 45
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = cur_fn = hdr = None
agg = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split('/')[-1]; continue
    if r[0] == "Function Name": cur_fn = r[1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or pat not in (cur_fn or ''): continue
    if r[0] != "":
        try: n = int(r[hdr.index("Instructions Executed")]); s = int(r[hdr.index("# Samples")])
        except ValueError: continue
        a = agg.setdefault((cur_file, int(r[0]), r[1].strip()[:100]), [0, 0]); a[0] += n; a[1] += s
tot = sum(v[0] for v in agg.values()) or 1; tots = sum(v[1] for v in agg.values()) or 1
print("total warp-inst %d, samples %d" % (tot, tots))
for k, v in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
    print("%5.1f%% inst %5.1f%% smp  %s:%d  %s" % (100 * v[0] / tot, 100 * v[1] / tots, k[0], k[1], k[2]))
