set -u
python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base mangled -k regex:gn_step -s 20 -c 60 --csv --log-file gpurun_out/step.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_step.log 2>&1
python - <<'PY'
import csv
lines=[l for l in open('gpurun_out/step.csv') if not l.startswith('==')]
d=[float(r['Metric Value'].replace(',','')) for r in csv.DictReader(lines)]
print(len(d), sorted(round(x/1000,1) for x in d))
PY
