#!/bin/bash
# round-2 GPU call F: full GPU suite (with durations), ncu launch list of the bench command, ncu --set full captures of
# machine_kernel (a bulk slice) and drain_kernel (first dense pass), summarised on the box
OUT=gpurun_out
mkdir -p $OUT
echo "== F1 full GPU suite"
timeout 1500 python -m pytest tests -m gpu -x -q --durations=6 > $OUT/r2f_pytest_gpu.log 2>&1; echo "exit $?"; tail -12 $OUT/r2f_pytest_gpu.log
echo "== F2 launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $OUT/r02_launches_bench.csv \
  python bench.py --steps 1 --warmup 1 --nodes 300000 --trees 0 --volume-trees 0 --cpu-sample 1000 > $OUT/r2f_launch_bench.json 2> $OUT/r2f_launch_err.log
echo "exit $?"; wc -l $OUT/r02_launches_bench.csv
echo "== F3 machine_kernel --set full (second slice of the 10^6-node pass)"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:machine_kernel -s 1 -c 1 -f -o $OUT/r02_machine \
  python bench.py --steps 1 --warmup 0 --nodes 1000000 --trees 0 --volume-trees 0 --cpu-sample 1000 > /dev/null 2> $OUT/r2f_ncu_machine_err.log
echo "exit $?"
python scripts/ncu_summary.py $OUT/r02_machine.ncu-rep > $OUT/r02_machine_kernel_bulk_slice.txt 2>&1
python scripts/ncu_stalls.py $OUT/r02_machine.ncu-rep >> $OUT/r02_machine_kernel_bulk_slice.txt 2>&1
tail -30 $OUT/r02_machine_kernel_bulk_slice.txt
echo "== F4 drain_kernel --set full (express + first dense launch)"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:drain_kernel -s 0 -c 2 -f -o $OUT/r02_drain \
  python bench.py --steps 1 --warmup 0 --nodes 1000000 --trees 0 --volume-trees 0 --cpu-sample 1000 > /dev/null 2> $OUT/r2f_ncu_drain_err.log
echo "exit $?"
ncu -i $OUT/r02_drain.ncu-rep --page raw --csv > $OUT/r02_drain_raw.csv 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open("gpurun_out/r02_drain_raw.csv")))
hdr=rows[0]
want=['Kernel Name','launch__grid_size','launch__block_size','launch__registers_per_thread','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__thread_inst_executed_per_inst_executed.ratio','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum']
with open("gpurun_out/r02_drain_kernel_first_pass.txt","w") as f:
    for r in rows[2:]:
        f.write("---- launch\n")
        for h,v,u in zip(hdr,r,rows[1]):
            if h in want: f.write("%-80s %-12s %s\n"%(h,u,v))
print(open("gpurun_out/r02_drain_kernel_first_pass.txt").read())
PY
rm -f $OUT/r02_drain.ncu-rep
ls -la $OUT | tail -15
