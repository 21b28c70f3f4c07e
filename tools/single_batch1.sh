timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 60 -c 90 --csv --log-file gpurun_out/single_b1.csv python tools/single_batch1.py > gpurun_out/single_b1.log 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/single_b1.csv')) if len(r)>10]
h=rows[0]; k=h.index('Kernel Name'); v=h.index('Metric Value'); g=h.index('Grid Size')
d=collections.defaultdict(list)
for r in rows[1:]:
    try: d[(r[k][:70],r[g])].append(float(r[v].replace(',','')))
    except: pass
for (n,gr),x in d.items(): print("%-72s %-14s n=%3d mean %.2f us min %.2f"%(n,gr,len(x),sum(x)/len(x)/1000,min(x)/1000))
PY
