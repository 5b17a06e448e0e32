import csv, subprocess, io, re, collections, json, os, sys
OUT = 'profiles'
os.makedirs(OUT, exist_ok=True)
def raw(rep):
    out = subprocess.run(['ncu','-i',rep,'--page','raw','--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]
KEYS = ['gpu__time_duration.sum','launch__grid_size','launch__block_size','launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic','dram__bytes_read.sum','dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','lts__t_bytes.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed' ,
        'sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum']
lines = ['# Round 1 ncu summaries', '',
         'Captured with `ncu --set full --clock-control none --import-source on` under gpurun (one B200); the `.ncu-rep`',
         'files stay in gpurun_out/ (scratch).  Durations under ncu are cold-cache and serialised: use SHARES, not absolutes.', '']
for title, rep in [('k_focf_fused_step: the cooperative FOCF step (forward | barrier | per-CTA batch statistics + gradients | barrier | Adam) at the ML-1M shape (scratch/fused_prof.py)', 'gpurun_out/r01_fused_step.ncu-rep'),
                   ('k_apply<kAdamFused> at the scale-out shape (2M users x 262k items, d=128; scratch/apply_prof.py)', 'gpurun_out/r01_apply_scaleout.ncu-rep'),
                   ('forward / loss / gradient kernels at the scale-out shape (B = 2^18)', 'gpurun_out/r01_step_scaleout.ncu-rep'),
                   ('k_fullsort_tc (tcgen05 3xTF32) 37,888 users x 262,144 items, d=128 (scratch/tc_prof.py)', 'gpurun_out/r01_fullsort_tc.ncu-rep'),
                   ('layer kernels of the PFCN / FairGo MLPs: k_gemm<true> (forward), k_wgrad, k_gemm<false> (backward-data) at M=9748, K=64, N=128 (scratch/wgrad_one.py)', 'gpurun_out/r01_layer_gemms.ncu-rep'),
                   ('k_linear_tc: MLP layer forward on tcgen05 (TMA raw K-blocks, in-smem TF32 split, 3xTF32 UMMA into TMEM) at M=2048,K=128,N=256 and M=9748,K=64,N=128 (scratch/linear_tc_prof.py)', 'gpurun_out/r01_linear_tc.ncu-rep'),
                   ('k_fullsort_exact at the ML-1M shape (first version, before the graph/TC work)', 'gpurun_out/prof_fullsort_r1.ncu-rep')]:
    if not os.path.exists(rep): continue
    hdr, units, rows = raw(rep)
    lines += [f'## {title}', '', f'source: `{rep}`', '']
    for r in rows:
        name = r[hdr.index('Kernel Name')]
        lines.append(f'**{name[:110]}**')
        lines.append('')
        lines.append('| metric | value | unit |')
        lines.append('|---|---|---|')
        for k in KEYS:
            if k in hdr:
                lines.append(f'| {k} | {r[hdr.index(k)]} | {units[hdr.index(k)]} |')
        lines.append('')
open(os.path.join(OUT, 'r01_ncu_summary.md'), 'w').write('\n'.join(lines))

# launch list of the default bench command
rows=[r for r in csv.reader(open('gpurun_out/r01_launches.csv')) if len(r)>5]
hdr=rows[0]; iK=hdr.index('Kernel Name'); iV=hdr.index('Metric Value')
agg=collections.defaultdict(list)
for r in rows[1:]:
    agg[re.sub(r'\(.*','',r[iK]).replace('void ','').replace('fr::','')].append(float(r[iV])/1e3)
tot=sum(sum(v) for v in agg.values())
with open(os.path.join(OUT,'r01_launches_summary.md'),'w') as f:
    f.write('# Round 1 launch list -- `ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 python bench.py --steps 24 --warmup 3 --no-cpu-baseline --no-probe`\n\n')
    f.write('First 8000 launches of the default bench command (ML-1M shape): warm-up steps, the timed FOCF steps (CUDA-graph kernel nodes are\nprofiled individually: gather + prepare on the side stream, the cooperative fused step on the main one), the end-to-end steps,\nthe evaluation passes and the PFCN / FairGo legs.  Times are cold-cache and serialised by ncu: compare SHARES.\n\n')
    f.write('| kernel | launches | total us | share | avg us |\n|---|---|---|---|---|\n')
    for k,v in sorted(agg.items(), key=lambda kv:-sum(kv[1])):
        f.write(f'| {k} | {len(v)} | {sum(v):.1f} | {100*sum(v)/tot:.1f}% | {sum(v)/len(v):.2f} |\n')
print(open(os.path.join(OUT,'r01_launches_summary.md')).read()[:2500])
