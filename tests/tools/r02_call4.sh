#!/bin/bash
cd "$(dirname "$0")/../.."
O=gpurun_out
( timeout 1200 python -m pytest tests/test_train_gpu.py tests/test_trainer_gpu.py -m gpu -q 2>&1 | tail -25 ) > $O/r02c4_pytest.log; tail -25 $O/r02c4_pytest.log
cp omok-ai_b200/libomok_b200.so /tmp/keep.so
for v in f1h1 c3f1h1; do
  cp omok-ai_b200/_build/variants/$v.so omok-ai_b200/libomok_b200.so
  timeout 1500 python tools/net_error_study.py --positions 20000 --paths tc --label $v --out $O/r02c4_err_$v.json 2>&1 | tail -6
  cp /tmp/keep.so omok-ai_b200/libomok_b200.so
done
