ncu --set full --clock-control none --import-source on -k regex:knn_cell_kernel -s 1 -c 1 -o gpurun_out/pk_cell -f python tools/prof_knn_cells.py 16 ${1:-0.2} > gpurun_out/pk_ncu2.log 2>&1
tail -3 gpurun_out/pk_ncu2.log
