import csv, sys, subprocess
rep, kern = sys.argv[1], sys.argv[2]
kid = sys.argv[4] if len(sys.argv) > 4 else None
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','cuda,sass',] + (['--kernel-id', kid] if kid else ['-k','regex:'+kern]),capture_output=True,text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = None; hdr = None; agg = []
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur_file = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': continue
    if r[0] == 'Line No': hdr = r; continue
    if hdr is None or r[0] == '': continue
    try:
        ln = int(r[0])
    except ValueError:
        continue
    d = dict(zip(hdr[4:], r[4:]))
    def toi(v):
        try: return int(v)
        except Exception: return 0
    inst = toi(d.get('Instructions Executed','0'))
    samp = toi(d.get('# Samples','0'))
    agg.append((inst, samp, cur_file, ln, r[1].strip()[:90], d))
ti = sum(a[0] for a in agg); ts = sum(a[1] for a in agg)
print('total inst', ti, 'samples', ts)
for a in sorted(agg, key=lambda a: -a[0])[:top]:
    d = a[5]
    st = {k: int(v) for k, v in d.items() if k.startswith('stall_') and 'Not Issued' not in k and v not in ('', '-') and int(v) > 0}
    tops = ','.join('%s:%d' % (k[6:], v) for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print('%5.1f%% inst %5.1f%% samp  %s:%d  %s   [%s]' % (100*a[0]/ti, 100*a[1]/max(ts,1), a[2], a[3], a[4], tops))
