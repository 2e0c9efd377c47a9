"""Per-source-line totals of an `ncu --page source --csv --print-source cuda,sass` export:
python profiles/ncu_lines.py file.csv [top]  ->  file:line, instructions executed, stall samples."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur = None
tot = collections.OrderedDict()
hdr = None
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': continue
    if r[0] == 'Line No': hdr = r; continue
    if r[0] != '' and hdr:
        try:
            ie = int(r[hdr.index('Instructions Executed')]); ss = int(r[4])
        except ValueError:
            continue
        key = (cur, int(r[0]))
        a = tot.setdefault(key, [0, 0, r[1].strip()[:110]])
        a[0] += ie; a[1] += ss
T = sum(v[0] for v in tot.values()); S = sum(v[1] for v in tot.values())
print("total inst %d  samples %d" % (T, S))
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%-14s %4d  inst %5.1f%%  samp %5.1f%%  %s" % (k[0], k[1], 100.0 * v[0] / T, 100.0 * v[1] / S, v[2]))
