python tools/rank_sim.py 8 10
python tools/rank_sim.py 4 10
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_tmp.csv python tools/rank_sim.py 8 1 > /dev/null 2>&1
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
    out.append((r[idx['Kernel Name']].split('(')[0][-10:], round(v,1)))
print(out[-20:])
PY
