# launch list restricted to the Gauss-Newton kernels (durations by kernel and grid size)
set -u
python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base mangled -k regex:'gn_eval|rgb_step_gn|gn_step|gn_init' -s 20 -c 80 --csv --log-file gpurun_out/gn.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_gn.log 2>&1
python - <<'PY'
import csv, collections
lines=[l for l in open('gpurun_out/gn.csv') if not l.startswith('==')]
d=collections.defaultdict(list)
for r in csv.DictReader(lines):
    d[r['Kernel Name'][:40]+' '+r.get('Grid Size','')].append(round(float(r['Metric Value'].replace(',',''))/1000,1))
for k,v in d.items(): print(k, len(v), sorted(v))
PY
