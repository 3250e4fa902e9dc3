#!/bin/bash
cd "$(dirname "$0")/../.."
O=gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > $O/r02c5_pytest.log; tail -8 $O/r02c5_pytest.log
for g in 1 8 32 64 128 256 1024; do
  for sk in 2048 0; do
    [ $g = 1024 ] && [ $sk = 0 ] && continue
    OMK_FC0_SPLITK_MAX=$sk python tools/profile_step.py --games $g --plies 1 --warm 1 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); k=d['kinds']
print('c3f1h1+groups games $g splitk_max $sk', 'sims/s %.3fM'%(d['sims_per_s']/1e6), ' '.join('%s %.4f'%(n, k[n]['ms']/max(1,k[n]['launches'])) for n in ('tower','fc0','fc1','heads','select_expand','apply')))"
  done
done
python bench.py --steps 5 --warmup 3 > $O/r02c5_bench.json 2> $O/r02c5_bench.err; tail -2 $O/r02c5_bench.err; python -c "
import json; d=json.load(open('$O/r02c5_bench.json'))
print({k: d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e'], d['transition_stream'], d['roofline']['avg_launch_ms'], d['roofline_second']['avg_launch_ms'], d['clocks'], d.get('cpu_baseline',{}).get('value'))"
