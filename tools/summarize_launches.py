"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total
device time and share of the step.  usage: summarize_launches.py launches.csv > summary.md"""
import csv, sys, collections, re
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>5]
hdr=rows[0]
ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value'); iu=hdr.index('Metric Unit')
agg=collections.OrderedDict()
tot=0.0
for r in rows[1:]:
    try: v=float(r[iv].replace(',',''))
    except ValueError: continue
    u=r[iu]
    us = v/1e3 if u in ('nsecond','ns') else (v if u in ('usecond','us') else v*1e3 if u in ('msecond','ms') else v)
    name=re.sub(r'\(.*$','',r[ik]).strip()[:110]
    a=agg.setdefault(name,[0,0.0]); a[0]+=1; a[1]+=us; tot+=us
ours=sum(v[1] for k,v in agg.items() if 'mflow' in k or k.startswith(('rqs_','chanmix','squeeze','permute','logtanh','gauss','jtj','affine','conv_')))
print(f'# launch list summary: {sum(v[0] for v in agg.values())} launches, {tot/1e3:.2f} ms total device time (cold-cache, serialised: compare shares)\n')
print(f'libmflow_b200 kernels: {ours/1e3:.2f} ms = {100*ours/tot:.1f}% of the device time\n')
print('| kernel | launches | total us | share |\n|---|---:|---:|---:|')
for k,(c,t) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:60]:
    print(f'| `{k}` | {c} | {t:.0f} | {100*t/tot:.1f}% |')
