# Round-2 evidence run (one B200): tests, bench at the driver's flags and at the defaults, then the ncu captures.
set -x
mkdir -p gpurun_out/r02
python -m pytest tests -q -m gpu 2>&1 | tail -3 > gpurun_out/r02/pytest_gpu.log
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02/bench_reference_k20.json 2> gpurun_out/r02/bench_reference_k20.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r02/bench_k20.json 2> gpurun_out/r02/bench_k20.err
python bench.py --no-cpu-baseline > gpurun_out/r02/bench_default.json 2> gpurun_out/r02/bench_default.err
python profiles/chain_bench.py > gpurun_out/r02/chain_p2.log 2>&1
python profiles/chain_bench.py --players 9 --steps 100 > gpurun_out/r02/chain_p9.log 2>&1
python profiles/chain_bench.py --players 4 --steps 100 > gpurun_out/r02/chain_p4.log 2>&1
python profiles/ramp_probe.py > gpurun_out/r02/ramp_probe.log 2>&1
python profiles/eval_bench.py > gpurun_out/r02/eval_bench.log 2>&1
PKB_EVAL_IMPL=hash3 python profiles/eval_bench.py >> gpurun_out/r02/eval_bench.log 2>&1
python profiles/exhaustive_bench.py > gpurun_out/r02/exhaustive_bench.log 2>&1
python profiles/ql_breakdown.py > gpurun_out/r02/ql_breakdown.log 2>&1
PKB_ENV_PHASE_ROUNDS=0 python profiles/ql_breakdown.py > gpurun_out/r02/ql_breakdown_one_phase.log 2>&1
python profiles/nine_mix_probe.py > gpurun_out/r02/nine_mix_probe.log 2>&1
python profiles/repeat_probe.py > gpurun_out/r02/repeat_probe.log 2>&1
# launch list of the bench command (serialised, cold cache: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02/launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02/ncu_launches.log 2>&1
# steady-state DRAM traffic of the pipelined launches (range replay: the launches overlap, no cache flush)
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,sm__warps_active.avg.pct_of_peak_sustained_active
PKB_RANGE_LAUNCHES=24 ncu --replay-mode range --cache-control none --clock-control none --metrics $M --csv --log-file gpurun_out/r02/range_p2.csv python profiles/range_probe.py > gpurun_out/r02/ncu_range_p2.log 2>&1
PKB_RANGE_LAUNCHES=12 PKB_RANGE_PLAYERS=9 ncu --replay-mode range --cache-control none --clock-control none --metrics $M --csv --log-file gpurun_out/r02/range_p9.csv python profiles/range_probe.py > gpurun_out/r02/ncu_range_p9.log 2>&1
# full-set captures of one launch each (L2 flushed before the step by prof_step.py)
ncu --set full --clock-control none --import-source on -k regex:step_warp -s 36 -c 1 -f -o gpurun_out/r02/step_p2 python profiles/prof_step.py --steps 40 > gpurun_out/r02/ncu_full_p2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_warp -s 36 -c 1 -f -o gpurun_out/r02/step_p9 python profiles/prof_step.py --players 9 --steps 40 > gpurun_out/r02/ncu_full_p9.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:eval_masks_fast -s 2 -c 1 -f -o gpurun_out/r02/eval_fast python profiles/eval_prof.py > gpurun_out/r02/ncu_full_eval.log 2>&1
ls -la gpurun_out/r02
