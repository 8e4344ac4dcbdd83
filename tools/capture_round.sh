#!/bin/bash
# One GPU call that collects what profiles/ keeps for a round: tests, the driver-like bench line, the reference arm,
# the block trace of a step, the ncu launch list of the bench command and one `ncu --set full` capture of the step's kernels.
#   gpurun -- 'bash tools/capture_round.sh r3a'
out=gpurun_out/${1:-cap}; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q > $out/gpu_tests.log 2>&1; echo "pytest rc=$?" >> $out/gpu_tests.log; tail -3 $out/gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $out/smoke.log 2>&1; tail -1 $out/smoke.log
timeout 400 python bench.py --steps 20 --warmup 5 > $out/bench20.json 2> $out/bench20.err; echo "bench rc=$?"
timeout 400 python bench.py --impl reference --steps 20 --warmup 5 > $out/ref.json 2> $out/ref.err; echo "ref rc=$?"
timeout 300 python bench.py --steps 200 --warmup 10 --no-extra --no-cpu > $out/bench200.json 2> $out/bench200.err
timeout 200 python tools/trace_step.py 32 > $out/trace_step.log 2>&1
timeout 200 python tools/skip_probe.py > $out/skip_probe.log 2>&1
timeout 300 python tools/run_steps.py 12 > $out/plain_steps.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $out/launches.csv python bench.py --steps 20 --warmup 5 --no-cpu --no-extra --e2e-steps 3 > $out/ncu_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"ff_forward2|pair_gemm|wide_td_dz2|wide_update|umma_gemm|wide_gather" -s 40 -c 9 -o $out/step_full -f python tools/run_steps.py 12 > $out/ncu_full.log 2>&1
ls -la $out
