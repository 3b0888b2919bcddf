#!/bin/bash
mkdir -p gpurun_out
python tools/post_time.py 1; python tools/post_time.py 64
for b in 1 64; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"select_infer|sort_keys|nms_kernel" -c 60 --csv --log-file gpurun_out/post_b$b.csv python tools/post_time.py $b 3 > /dev/null 2>&1
  python - <<PY
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/post_b$b.csv')) if len(r)>10 and r[0].isdigit()]
d=collections.defaultdict(list)
for r in rows: d[r[4].split('(')[0]].append(float(r[-1]))
for k,v in d.items(): print('B=$b', k, 'n', len(v), 'first-half mean us', round(sum(v[:len(v)//2])/max(1,len(v)//2),1), 'second-half mean us', round(sum(v[len(v)//2:])/max(1,len(v)-len(v)//2),1))
PY
done
