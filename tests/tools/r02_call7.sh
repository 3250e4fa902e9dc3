#!/bin/bash
cd "$(dirname "$0")/../.."
O=gpurun_out
( timeout 1500 python -m pytest tests/test_net_gpu.py tests/test_golden_gpu.py tests/test_trainer_gpu.py -m gpu -x -q 2>&1 | tail -25 ) > $O/r02c7_pytest.log; tail -6 $O/r02c7_pytest.log
split() { python tools/profile_step.py --games 1024 --plies 2 --warm 2 "$@" 2>/dev/null; }
cp omok-ai_b200/libomok_b200.so /tmp/keep.so
for v in current A; do
  [ $v != current ] && cp omok-ai_b200/_build/variants/$v.so omok-ai_b200/libomok_b200.so
  for rep in 1 2; do split | python -c "
import json,sys; d=json.load(sys.stdin); k=d['kinds']
print('$v', 'sims/s %.3fM'%(d['sims_per_s']/1e6), ' '.join('%s %.4f'%(n, k[n]['ms']/max(1,k[n]['launches'])) for n in ('tower','fc0','fc1','heads','select_expand','apply')))"; done
  split --lanes 1 | python -c "
import json,sys; d=json.load(sys.stdin); print('$v 2 lanes', 'sims/s %.3fM'%(d['sims_per_s']/1e6))"
  cp /tmp/keep.so omok-ai_b200/libomok_b200.so
done
timeout 600 python tests/tools/check_f16.py 600 2>&1 | tail -34
