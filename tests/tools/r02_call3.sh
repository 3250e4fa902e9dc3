#!/bin/bash
cd "$(dirname "$0")/../.."
O=gpurun_out
( timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > $O/r02c3_pytest.log; tail -25 $O/r02c3_pytest.log
python tools/iteration.py --games-per-gpu 256 --plies 4 --steps 10 > $O/r02c3_iter_1gpu.json 2> $O/r02c3_iter_1gpu.err; cat $O/r02c3_iter_1gpu.json; tail -3 $O/r02c3_iter_1gpu.err
cp omok-ai_b200/libomok_b200.so /tmp/keep.so
for v in current c3f1h1; do
  [ $v != current ] && cp omok-ai_b200/_build/variants/$v.so omok-ai_b200/libomok_b200.so
  for g in 1 8 32 64 128 256; do
    for sk in 2048 0; do
      OMK_FC0_SPLITK_MAX=$sk python tools/profile_step.py --games $g --plies 1 --warm 1 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); k=d['kinds']
print('$v games $g splitk_max $sk', 'sims/s %.3fM'%(d['sims_per_s']/1e6), ' '.join('%s %.4f'%(n, k[n]['ms']/max(1,k[n]['launches'])) for n in ('tower','fc0','fc1','heads','select_expand','apply')))"
    done
  done
  cp /tmp/keep.so omok-ai_b200/libomok_b200.so
done
