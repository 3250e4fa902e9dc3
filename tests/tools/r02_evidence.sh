#!/bin/bash
# round 2 evidence run (one GPU): parity tests, bench (both arms), per-kernel splits, sweeps, error study, ncu captures
cd "$(dirname "$0")/../.."
O=gpurun_out/ev
mkdir -p $O
( timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 ) > $O/pytest_gpu.log; tail -3 $O/pytest_gpu.log
python bench.py --steps 5 --warmup 3 > $O/bench_1gpu.json 2> $O/bench_1gpu.err; tail -1 $O/bench_1gpu.err
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_1gpu_20steps.json 2>/dev/null
python bench.py --games 2048 --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_1gpu_2048games.json 2>/dev/null
python bench.py --impl reference --steps 5 --warmup 3 > $O/bench_reference_arm.json 2>/dev/null
split() { python tools/profile_step.py --plies 2 --warm 2 "$@" 2>/dev/null; }
split --games 1024 > $O/split_1024_1lane.json
split --games 1024 --lanes 1 > $O/split_1024_2lanes.json
for g in 256 512 768; do split --games $g --lanes 1 > $O/split_${g}.json; done
python tools/profile_step.py --tree-sweep > $O/tree_sweep.json 2>/dev/null
python tools/profile_step.py --env > $O/env_sweep.json 2>/dev/null
python tools/single_game.py 40 > $O/single_game_gpu.json 2>/dev/null
python tests/tools/single_game_cpu.py 4 > $O/single_game_cpu.json 2>/dev/null
timeout 900 python tools/net_error_study.py --positions 20000 --paths tc simt cpu32 --label default --out $O/err_default.json > $O/err_default.log 2>&1; tail -13 $O/err_default.log
timeout 600 python tools/net_error_study.py --positions 20000 --paths tc --fc0-chunk 3 --label fc0_chunk3 --out $O/err_chunk3.json > $O/err_chunk3.log 2>&1; tail -5 $O/err_chunk3.log
timeout 300 python tests/tools/check_f16.py 600 > $O/check_f16.log 2>&1
# ncu: launch list of the bench command, then full captures (each after the plain run of the same command above)
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 800 --csv --log-file $O/ncu_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_launches.log 2>&1
python tools/profile_step.py --games 1024 --plies 1 --warm 1 > $O/plain_split.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_tower16|k_fc16' -s 24 -c 4 -o $O/prof_net python tools/profile_step.py --games 1024 --plies 1 --warm 1 > $O/ncu_net.log 2>&1
python tools/tree_pool_run.py 1024 2048 160 > $O/plain_tree.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_select_expand|k_apply' -s 8 -c 4 -o $O/prof_tree python tools/tree_pool_run.py 1024 2048 160 > $O/ncu_tree.log 2>&1
python tools/profile_step.py --env > $O/plain_env.log 2>&1 && \
ncu --set full --clock-control none -k regex:k_env_step -s 27 -c 2 -o $O/prof_env python tools/profile_step.py --env > $O/ncu_env.log 2>&1
ls -la $O
