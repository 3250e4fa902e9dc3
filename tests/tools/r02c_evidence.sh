#!/bin/bash
# round 2, second session: the evidence runs behind profiles/r02_train_step.md, r02_tree_pool_sweep.json,
# r02_ncu_tree_after_runs_*.csv, r02_small_pools.json, r02_iteration_reference_shaped_1gpu.json (one GPU; `gpurun -- bash tests/tools/r02c_evidence.sh`)
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
# trainer step: parity tests, wall clock at three minibatch sizes, launch list of three steps
( timeout 600 python -m pytest tests/test_train_gpu.py tests/test_trainer_gpu.py -q 2>&1 | tail -3 ) > $O/c_pytest_train.log; cat $O/c_pytest_train.log
for n in 128 64 16; do python tools/train_step_time.py --n $n; done > $O/c_train_step_time.jsonl 2>/dev/null; cat $O/c_train_step_time.jsonl
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/c_train_launches.csv python tools/train_step_time.py --steps 2 --warm 1 > $O/c_train_ncu.log 2>&1
python tools/ncu_launch_shares.py $O/c_train_launches.csv | head -20
# an iteration of the reference's own shape (src/config.rs:89-100): 50-game pool at 600 simulations per move, 600 updates
python tools/iteration.py --games-per-gpu 50 --plies 80 --count 600 --steps 600 > $O/c_iteration_reference_shaped.json 2>/dev/null; cat $O/c_iteration_reference_shaped.json
# tree kernels: parity (incl. the late-game positions), pool-size sweep, small pools, ncu rows at 2 048 and 65 536 trees
( timeout 600 python -m pytest tests/test_tree_gpu.py tests/test_golden_gpu.py tests/test_vloss_gpu.py tests/test_arena_gpu.py tests/test_selfplay_gpu.py -q 2>&1 | tail -3 ) > $O/c_pytest_tree.log; cat $O/c_pytest_tree.log
python tools/profile_step.py --tree-sweep > $O/c_tree_sweep.json 2>/dev/null; cat $O/c_tree_sweep.json
for g in 256 512; do python tools/profile_step.py --games $g --plies 2 --warm 2 > $O/c_pool_$g.json 2>/dev/null; done
python tools/single_game.py > $O/c_single_game.json 2>/dev/null; tail -1 $O/c_single_game.json
python tools/tree_pool_run.py 1024 2048 160 > $O/c_plain_tree.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_select_expand|k_apply' -s 8 -c 2 -o $O/c_prof_tree python tools/tree_pool_run.py 1024 2048 160 > $O/c_ncu_tree.log 2>&1
python tools/tree_pool_run.py 32768 768 160 > $O/c_plain_tree65k.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_select_expand|k_apply' -s 8 -c 2 -o $O/c_prof_tree65k python tools/tree_pool_run.py 32768 768 160 > $O/c_ncu_tree65k.log 2>&1
python tools/ncu_summary.py $O/c_tree $O/c_prof_tree.ncu-rep > /dev/null 2>&1; python tools/ncu_summary.py $O/c_tree65k $O/c_prof_tree65k.ncu-rep > /dev/null 2>&1
ls -la $O | tail -20
