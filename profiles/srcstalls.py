"""Summarise an `ncu --page source --csv --print-source cuda,sass` dump: stall samples and executed
instructions per CUDA source line.  usage: python profiles/srcstalls.py file.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
cur_file = None
agg = {}
hdr = None
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur_file = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': continue
    if r[0] == 'Line No': hdr = r; continue
    if hdr is None or len(r) < 8: continue
    if r[0].strip().isdigit():            # a source line summary row
        i_s, i_i = hdr.index('# Samples'), hdr.index('Instructions Executed')
        try:
            s = int(r[i_s]); n = int(r[i_i])
        except ValueError:
            continue
        key = (cur_file, int(r[0]), r[1].strip()[:100])
        a = agg.setdefault(key, [0, 0]); a[0] += s; a[1] += n
tot_s = sum(a[0] for a in agg.values()); tot_i = sum(a[1] for a in agg.values())
print(f"total samples {tot_s}, warp instructions {tot_i}")
for (f, ln, src), (s, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*s/max(tot_s,1):5.1f}% smp {100*n/max(tot_i,1):5.1f}% ins  {f}:{ln:<4d} {src}")
