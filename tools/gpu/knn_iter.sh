# correctness of dc_knn_cells against dc_knn + kernel durations and instruction counts (developer loop on the GPU box)
timeout 600 python tools/check_knn_cells.py 16 0.3,0.2,0.15 > gpurun_out/knncells.log 2>&1; echo rc=$? >> gpurun_out/knncells.log
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:knn_ --csv --log-file gpurun_out/pk_launches.csv python tools/prof_knn_cells.py 16 ${1:-0.2} > gpurun_out/pk_ncu1.log 2>&1
cat gpurun_out/knncells.log
python - <<'PY'
import csv
rows = list(csv.reader(open('gpurun_out/pk_launches.csv')))
hdr = None
for r in rows:
    if r and r[0] == 'ID':
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        print(d['ID'], d['Kernel Name'].split('(')[0][:40], d['Metric Name'], d['Metric Value'])
PY
