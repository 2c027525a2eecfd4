"""per-source-line instruction / stall-sample shares from `ncu --page source --csv --print-source sass,cuda`"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
fn = fp = hdr = None
data = {}
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        fp = r[1]
        continue
    if r[0] == 'Function Name':
        fn = r[1][:70]
        continue
    if r[0] == 'Line No':
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        if r[2] != '-':  # SASS row; line rows carry the per-line aggregate
            continue
        try:
            ie = int(d['Instructions Executed'])
        except ValueError:
            continue
        if ie == 0:
            continue
        data.setdefault(fn, []).append((fp.split('/')[-1], int(d['Line No']), ie, int(d['# Samples']), r[1][:100]))
for fn, l in data.items():
    tot = sum(x[2] for x in l)
    ts = sum(x[3] for x in l)
    print('=====', fn, 'total inst', tot, 'samples', ts)
    for x in sorted(l, key=lambda x: -x[3])[:top]:
        print(f"{x[0]:16s} {x[1]:4d} inst={100 * x[2] / tot:5.1f}% samp={100 * x[3] / ts:5.1f}%  {x[4]}")
