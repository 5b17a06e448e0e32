import csv, re, sys, io, subprocess
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 14
out = subprocess.run(['ncu','-i',rep,'--page','source','--csv'], capture_output=True, text=True).stdout
lines = out.splitlines()
starts = [i for i,l in enumerate(lines) if l.startswith('"Kernel Name"')]
for si, st in enumerate(starts):
    en = starts[si+1] if si+1 < len(starts) else len(lines)
    rows = list(csv.reader(io.StringIO('\n'.join(lines[st:en]))))
    name = rows[0][1]; hdr = rows[1]
    iS = hdr.index('Warp Stall Sampling (All Samples)'); iI = hdr.index('Instructions Executed'); isrc = hdr.index('Source')
    stall_cols = [(i,h) for i,h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
    data = [r for r in rows[2:] if len(r) > iS and r[iS].isdigit()]
    tot = sum(int(r[iS]) for r in data) or 1
    print(f'=== {name[:70]}  samples={tot} inst={sum(int(r[iI]) for r in data)}')
    agg = {h: sum(int(r[i]) for r in data if r[i].isdigit()) for i,h in stall_cols}
    print('   stalls:', ', '.join(f'{h[6:]}={100*v/tot:.0f}%' for h,v in sorted(agg.items(), key=lambda kv:-kv[1])[:6]))
    for r in sorted(data, key=lambda r: -int(r[iS]))[:topn]:
        why = max(stall_cols, key=lambda ih: int(r[ih[0]]) if r[ih[0]].isdigit() else 0)[1][6:]
        print(f'{100*int(r[iS])/tot:5.1f}%  n={r[iI]:>7s} {why:10s} {r[isrc][:100]}')
