ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_tmp.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/launches_tmp.csv')))
for i,r in enumerate(rows):
    if 'Kernel Name' in r: hdr=r; start=i; break
idx={h:i for i,h in enumerate(hdr)}
out=[]
for r in rows[start+1:]:
    if len(r)<len(hdr) or r[idx['Metric Name']]!='gpu__time_duration.sum': continue
    v=float(r[idx['Metric Value']].replace(',','')); u=r[idx['Metric Unit']]
    if u.startswith('n'): v/=1000
    elif u.startswith('m'): v*=1000
    out.append((r[idx['Kernel Name']].split('(')[0][-14:], round(v,1)))
print(out[-15:])
PY
