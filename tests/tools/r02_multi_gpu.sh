#!/bin/bash
# N GPUs (gpurun --gpus N -- bash tests/tools/r02_multi_gpu.sh N): at N = 2 the NCCL train-step test; the iteration at 1024 games per GPU; the bench; at N = 8 also BASELINE config 4 (2048 games per GPU, merged replay)
cd "$(dirname "$0")/../.."
O=gpurun_out
N=${1:-2}
P=29511
if [ $N = 2 ]; then ( timeout 600 python -m pytest tests/test_train_gpu.py -m gpu -q 2>&1 | tail -8 ) > $O/r02_pytest_2gpu.log; tail -4 $O/r02_pytest_2gpu.log; fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P tools/iteration.py --games-per-gpu 1024 --plies 60 --steps 20 --gather > $O/r02_iteration_${N}gpu.json 2> $O/r02_iteration_${N}gpu.err; tail -2 $O/r02_iteration_${N}gpu.err; cat $O/r02_iteration_${N}gpu.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((P+1)) bench.py --gpus $N --steps 5 --warmup 3 > $O/r02_bench_${N}gpu.json 2> $O/r02_bench_${N}gpu.err; tail -2 $O/r02_bench_${N}gpu.err
python -c "
import json; d=json.load(open('$O/r02_bench_${N}gpu.json')); print('bench', d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])"
if [ $N = 8 ]; then
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((P+2)) bench.py --gpus $N --games 2048 --steps 5 --warmup 3 --gather --no-cpu-baseline > $O/r02_bench_${N}gpu_2048games.json 2> $O/r02_bench_${N}gpu_2048games.err; tail -2 $O/r02_bench_${N}gpu_2048games.err
python -c "
import json; d=json.load(open('$O/r02_bench_${N}gpu_2048games.json')); print('bench2048', d['n_gpus'], d['value'], d['ms_per_step'], d['replay_merge'], d['transition_stream'])"
fi
