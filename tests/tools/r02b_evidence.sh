#!/bin/bash
# round 2, second evidence run (one GPU) after the tower's row stencil / split conv2: parity tests, bench (both arms),
# per-kernel splits, the tensor-core error study, phase stamps, ncu captures of the network kernels
cd "$(dirname "$0")/../.."
O=gpurun_out/ev2
mkdir -p $O
( timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 ) > $O/pytest_gpu.log; tail -3 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2
python bench.py --steps 5 --warmup 3 > $O/bench_1gpu.json 2> $O/bench_1gpu.err; tail -1 $O/bench_1gpu.err
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_1gpu_20steps.json 2>/dev/null
python bench.py --games 2048 --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_1gpu_2048games.json 2>/dev/null
python bench.py --impl reference --steps 5 --warmup 3 > $O/bench_reference_arm.json 2>/dev/null
split() { python tools/profile_step.py --plies 2 --warm 2 "$@" 2>/dev/null; }
split --games 1024 > $O/split_1024_1lane.json
split --games 1024 --lanes 1 > $O/split_1024_2lanes.json
for g in 256 512 768; do split --games $g --lanes 1 > $O/split_${g}.json; done
python tools/single_game.py 40 > $O/single_game_gpu.json 2>/dev/null
timeout 900 python tools/net_error_study.py --positions 20000 --paths tc --label default --out $O/err_default.json > $O/err_default.log 2>&1; tail -8 $O/err_default.log
timeout 300 python tests/tools/check_f16.py 600 > $O/check_f16.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 800 --csv --log-file $O/ncu_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_launches.log 2>&1
python tools/profile_step.py --games 1024 --plies 1 --warm 1 > $O/plain_split.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_tower16|k_fc16' -s 24 -c 4 -o $O/prof_net python tools/profile_step.py --games 1024 --plies 1 --warm 1 > $O/ncu_net.log 2>&1
ls -la $O
