# round-2 (r): launch list of one CubeNET training step on the final library, default and deterministic mode.
# Each ncu command follows a plain run of the same command line that exited 0.
set -x
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"
S=gpurun_out
timeout 60 python tools/prof_step.py > $S/plain_step_r2r.log 2>&1 &&
timeout 90 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file $S/launches_r2r_step.csv python tools/prof_step.py > $S/ncu_r2r_1.log 2>&1
HPRI_DETERMINISTIC=1 timeout 60 python tools/prof_step.py > $S/plain_step_det_r2r.log 2>&1 &&
HPRI_DETERMINISTIC=1 timeout 90 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file $S/launches_r2r_step_det.csv python tools/prof_step.py > $S/ncu_r2r_2.log 2>&1
wc -l $S/launches_r2r_step.csv $S/launches_r2r_step_det.csv
