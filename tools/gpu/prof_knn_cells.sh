set -x
python tools/prof_knn_cells.py 16 0.2 > gpurun_out/pk_plain.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:knn --csv --log-file gpurun_out/pk_launches.csv python tools/prof_knn_cells.py 16 0.2 > gpurun_out/pk_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:knn_cell_kernel -s 1 -c 1 -o gpurun_out/pk_cell -f python tools/prof_knn_cells.py 16 0.2 > gpurun_out/pk_ncu2.log 2>&1
cat gpurun_out/pk_plain.log; cat gpurun_out/pk_launches.csv | tail -8
