#!/bin/bash
# ncu --set full of the recorded kNN kernel on the corridor map ($1 scans, default 64), summarised on the box (the report
# stays there if large)
cd "$GRAFT_REPO_ROOT"
export DC_KNN=record
N=${1:-64}
timeout 300 python tools/prof_knn_one.py $N || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:knn_record -s 1 -c 1 -o gpurun_out/kr -f python tools/prof_knn_one.py $N > gpurun_out/kr_ncu.log 2>&1
tail -2 gpurun_out/kr_ncu.log
python tools/ncu_summary.py gpurun_out/kr.ncu-rep > gpurun_out/kr_summary.md 2>&1
python tools/ncu_lines.py gpurun_out/kr.ncu-rep knn_record 70 > gpurun_out/kr_lines.txt 2>&1
ls -la gpurun_out/kr.ncu-rep
[ $(stat -c %s gpurun_out/kr.ncu-rep) -gt 30000000 ] && rm gpurun_out/kr.ncu-rep
head -80 gpurun_out/kr_lines.txt | cut -c1-200
