#!/bin/bash
cd "$(dirname "$0")/../.."
O=gpurun_out
( timeout 1500 python -m pytest tests/test_train_gpu.py tests/test_trainer_gpu.py tests/test_cpp_mirror.py tests/test_selfplay_gpu.py -m gpu -x -q 2>&1 | tail -12 ) > $O/r02c10_pytest.log; tail -5 $O/r02c10_pytest.log
python tools/iteration.py --games-per-gpu 256 --plies 4 --steps 20 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('train_step_ms', d['train_step_ms'], 'losses', d['first_loss'], d['last_loss'])"
