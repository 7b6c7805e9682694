"""Per-source-line hot spots from an .ncu-rep (needs -lineinfo): stall samples + instructions.
python profiles/ncu_source_hot.py <rep> [top_n]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file, hdr, lines = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        cur_file = r[1].split('/')[-1]
    elif r[0] == 'Line No':
        hdr = r
    elif hdr and r[0].isdigit():
        d = dict(zip(hdr, r))
        try:
            lines.append((int(d['# Samples'] or 0), int(d['Instructions Executed'] or 0),
                          int(d['Thread Instructions Executed'] or 0), cur_file, int(r[0]), r[1].strip()[:110]))
        except Exception:
            pass
tot_s = sum(l[0] for l in lines) or 1
tot_i = sum(l[1] for l in lines) or 1
print('total samples %d, warp instructions %d' % (tot_s, tot_i))
for s, i, t, f, n, src in sorted(lines, reverse=True)[:top]:
    print('%5.1f%% smp %5.1f%% inst  lanes %4.1f  %s:%d  %s' % (100.0 * s / tot_s, 100.0 * i / tot_i, t / max(i, 1), f, n, src))
