# per-kernel duration, warp instructions and active lanes of the velocity solvers, with and without the long-solve split (1M envs)
# (the code under test was measured and NOT kept: DESIGN.md §8, "Measured dead ends"; this script is the record of the A/B)
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active
for L in 0 1; do
QB_ENVS=1048576 MRP_SPARES=0 MRP_LONG=$L ncu --metrics $M --clock-control none -k 'regex:^k_solve_(vel|vel_long|pos)$' -s $((62*(2+L))) -c $((2*(2+L))) --csv --log-file gpurun_out/r2_exp49_long$L.csv python profiles/profile_step.py > gpurun_out/r2_exp49_$L.log 2>&1
done
