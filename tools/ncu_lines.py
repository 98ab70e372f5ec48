"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump by CUDA source line.
usage: python tools/ncu_lines.py dump.csv <records-per-launch> [top]"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
nrec = float(sys.argv[2]); top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
fp = None; src = {}; hdr = None; cur = None
inst = collections.Counter(); ops = collections.defaultdict(collections.Counter)
stalls = collections.defaultdict(collections.Counter)
names = ('stall_long_sb', 'stall_wait', 'stall_short_sb', 'stall_no_inst', 'stall_dispatch', 'stall_not_selected',
         'stall_selected', 'stall_barrier', 'stall_mio', 'stall_branch_resolving', 'stall_math', 'stall_lg', '# Samples')
for r in rows:
    if not r: continue
    if r[0] == 'File Path': fp = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': continue
    if r[0] == 'Line No': hdr = r; continue
    if r[0] != '': cur = (fp, int(r[0])); src[cur] = r[1]; continue
    if r[2] in ('-', '...'): continue
    try: n = int(r[hdr.index('Instructions Executed')])
    except Exception: continue
    inst[cur] += n
    t = r[3].split(); o = t[1] if t[0].startswith('@') else t[0]
    ops[cur][o.split('.')[0]] += n
    for nm in names:
        try: stalls[nm][cur] += int(r[hdr.index(nm)])
        except Exception: pass
warps = nrec / 32
tot = sum(inst.values())
print(f"total warp-instructions {tot}  per record-warp {tot / warps:.1f} (inlined lines are listed at every level: halve if doubled)")
print({k: sum(v.values()) for k, v in stalls.items()})
for k, v in inst.most_common(top):
    print(f"{v / warps:7.1f} {k[0]}:{k[1]:4d} {src[k][:95]}  {dict(ops[k].most_common(3))}")
for nm in ('stall_long_sb', 'stall_short_sb', 'stall_wait', 'stall_no_inst', 'stall_mio'):
    print('==', nm)
    for k, v in stalls[nm].most_common(6):
        print(f"  {v:6d} {k[0]}:{k[1]} {src[k][:100]}")
