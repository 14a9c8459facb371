# final measurements of the round (1 GPU): bench line, launch list of the timed region, ncu --set full of the repo's
# kernels in the timed region (summarised on the box: the report itself is too large to bring back)
set -x
python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-strong-anchor > gpurun_out/r2_ncu_launch.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:'knn_record_kernel|knn_thread_list_kernel|step_forward_kernel|step_chain_kernel|step_points_kernel|pack_records_batched_kernel|gather_points_kernel|cell_table_kernel|cell_keys_kernel|world_points_batched_kernel' \
    -o gpurun_out/r2_timed -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-strong-anchor > gpurun_out/r2_ncu_full.log 2>&1
tail -2 gpurun_out/r2_ncu_full.log
python tools/ncu_summary.py gpurun_out/r2_timed.ncu-rep gpurun_out/r2_timed_region_kernels.md "Round 2 final: kernels of the timed region of bench.py (64 scans, 8.37 M points, k=32)" gpurun_out/r2_traffic.json 8366086
python tools/ncu_lines.py gpurun_out/r2_timed.ncu-rep knn_record_kernel 50 > gpurun_out/r2_knn_record_lines.txt 2>&1
python tools/ncu_lines.py gpurun_out/r2_timed.ncu-rep step_forward_kernelILi0ELb1E 30 > gpurun_out/r2_step_forward_lines.txt 2>&1
ls -la gpurun_out/
sz=$(stat -c %s gpurun_out/r2_timed.ncu-rep); if [ "$sz" -gt 40000000 ]; then rm gpurun_out/r2_timed.ncu-rep; fi
