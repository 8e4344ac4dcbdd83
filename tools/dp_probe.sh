mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 10 --no-cpu --e2e-steps 5 > gpurun_out/r1o_$tag.json 2> gpurun_out/r1o_$tag.err; python - <<PY
import json
l=[x for x in open('gpurun_out/r1o_$tag.json') if x.startswith('{')]
if l:
    j=json.loads(l[0]); print('$tag', round(j['value']/1e6,1),'M/s', round(j['ms_per_step']*1e3,1),'us/step')
else:
    print('$tag FAILED'); print(open('gpurun_out/r1o_$tag.err').read()[-800:])
PY
}
run graph A=1
run eager GOLDQ_NO_GRAPH=1
run graph_ll NCCL_PROTO=LL
run graph_ll128 NCCL_PROTO=LL128
run graph_tree NCCL_ALGO=Tree
