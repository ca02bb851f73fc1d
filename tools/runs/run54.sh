python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/n8_err.log | tail -1 > gpurun_out/bench_n8.json
python -c "
import json; d=json.load(open('gpurun_out/bench_n8.json')); print('N=8', d['value'], d['ms_per_step'], d['e2e']['value'], d['n_gpus'], d['clocks'])"
grep -v "^\*\|OMP_NUM" gpurun_out/n8_err.log | tail -3 | cut -c1-300
