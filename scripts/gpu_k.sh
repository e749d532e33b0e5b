mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_tc.sum --clock-control none -k regex:"attention_kernel|linear_kernel" -c 200 --csv --log-file gpurun_out/attn_kernels.csv python benchmarks/micro_attn.py > gpurun_out/ncu_attn2.log 2>&1; echo "rc=$?"
python - <<'PY'
import csv, collections
rows=list(csv.reader(open('gpurun_out/attn_kernels.csv')))
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
h=rows[hdr]; ki=h.index('Kernel Name'); mi=h.index('Metric Name'); vi=h.index('Metric Value'); gi=h.index('Grid Size')
per=collections.OrderedDict()
for r in rows[hdr+2:]:
    if len(r)<=vi: continue
    per.setdefault(r[0],{'k':r[ki][:30],'g':r[gi]})[r[mi]]=r[vi]
tot=0
for i,(id_,d) in enumerate(per.items()):
    t=float(d['gpu__time_duration.sum'].replace(',',''))/1000
    if i<159: tot+=t
    if i<60: print(id_, d['k'], d['g'], f"{t:.1f}us", d.get('sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed'))
print('sum first 159 launches (one forward) us:', tot)
PY
