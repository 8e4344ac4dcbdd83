#!/bin/bash
# multi-GPU evidence for a round: DP parity tests on N GPUs, then bench.py as the driver launches it (strong = default) and weak.
#   gpurun --gpus N -- 'bash tools/dp_bench.sh N rXX'
N=$1; out=gpurun_out/${2:-dp}; mkdir -p $out
nvidia-smi -L | wc -l > $out/ngpus_$N.txt
if [ "$3" != "notest" ]; then timeout 600 python -m pytest tests/test_gpu_dp.py -m gpu -q > $out/dp_tests_w$N.log 2>&1; echo "dp tests rc=$?"; tail -2 $out/dp_tests_w$N.log; fi
for mode in strong weak; do
  extra=""; [ $mode = weak ] && extra="--scaling weak"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 --no-extra $extra > $out/n${N}_$mode.json 2> $out/n${N}_$mode.err; echo "bench $mode rc=$?"
  python tools/show_bench.py $out/n${N}_$mode.json 2>/dev/null | head -1 | cut -c1-120
done
