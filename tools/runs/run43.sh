python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/n2_err.log | tail -1 > gpurun_out/bench_n2.json
python -c "
import json; d=json.load(open('gpurun_out/bench_n2.json')); print('N=2', d['value'], d['ms_per_step'], d['e2e']['value'], d['n_gpus'], d['config'])"
tail -3 gpurun_out/n2_err.log | cut -c1-300
