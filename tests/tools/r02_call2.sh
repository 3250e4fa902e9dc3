#!/bin/bash
cd "$(dirname "$0")/../.."
O=gpurun_out
split() { python tools/profile_step.py --games 1024 --plies 2 --warm 2 "$@" 2>/dev/null; }
timeout 600 python tests/tools/check_f16.py 600 > $O/r02c2_check_f16.log 2>&1; tail -26 $O/r02c2_check_f16.log
cp omok-ai_b200/libomok_b200.so /tmp/keep.so
for v in f1h1 f2h1 c3f1h1 c6f2h1; do
  cp omok-ai_b200/_build/variants/$v.so omok-ai_b200/libomok_b200.so
  timeout 600 python tools/net_error_study.py --positions 6000 --seeds 0 --paths tc --label $v --out $O/r02c2_err_$v.json 2>&1 | tail -3
  split > $O/r02c2_split_$v.json
  python -c "
import json; d=json.load(open('$O/r02c2_split_$v.json')); k=d['kinds']
print('$v', 'sims/s %.3fM'%(d['sims_per_s']/1e6), ' '.join('%s %.4f'%(n, k[n]['ms']/max(1,k[n]['launches'])) for n in ('tower','fc0','fc1','heads')))"
  cp /tmp/keep.so omok-ai_b200/libomok_b200.so
done
timeout 900 python tools/net_error_study.py --positions 6000 --seeds 0 --paths cpu32 tc simt --label yardstick --out $O/r02c2_err_yardstick.json 2>&1 | tail -8
# source-level profile of the tower (one launch, after a plain run of the same command)
python tools/profile_step.py --games 1024 --plies 1 --warm 1 > $O/r02c2_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_tower16 -s 20 -c 1 -o $O/r02c2_tower python tools/profile_step.py --games 1024 --plies 1 --warm 1 > $O/r02c2_ncu.log 2>&1
ls -la $O/r02c2_tower.ncu-rep
