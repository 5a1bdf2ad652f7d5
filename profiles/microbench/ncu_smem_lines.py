"""Rank source lines of one kernel by shared-memory wavefronts (ncu source page).
usage: ncu_smem_lines.py report.ncu-rep kernel_regex [top]"""
import csv, sys, subprocess
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass', '-k', 'regex:' + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = None; hdr = None; agg = []
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur_file = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': continue
    if r[0] == 'Line No': hdr = r; continue
    if hdr is None or r[0] == '': continue
    try: ln = int(r[0])
    except ValueError: continue
    d = dict(zip(hdr[4:], r[4:]))
    def toi(v):
        try: return int(v)
        except Exception: return 0
    agg.append((toi(d.get('L1 Wavefronts Shared', '0')), toi(d.get('L1 Wavefronts Shared Ideal', '0')),
                toi(d.get('Instructions Executed', '0')), cur_file, ln, r[1].strip()[:100]))
tw = sum(a[0] for a in agg); ti = sum(a[2] for a in agg)
print('total shared wavefronts', tw, 'ideal', sum(a[1] for a in agg), 'instructions', ti)
for a in sorted(agg, key=lambda a: -a[0])[:top]:
    print('%5.1f%% wf (x%.2f of ideal) %5.1f%% inst  %s:%d  %s' % (100 * a[0] / max(tw, 1), a[0] / max(a[1], 1), 100 * a[2] / max(ti, 1), a[3], a[4], a[5]))
