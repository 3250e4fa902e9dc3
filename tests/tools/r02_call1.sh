#!/bin/bash
# round 2, GPU call 1: parity tests with the new default build, tower / tree A/B timings, phase stamps, network error study
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/r02c1_smi.txt 2>&1
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > $O/r02c1_pytest.log
echo "pytest done: $(tail -1 $O/r02c1_pytest.log)"
split() { python tools/profile_step.py --games 1024 --plies 2 --warm 2 "$@" 2>/dev/null; }
cp omok-ai_b200/libomok_b200.so /tmp/keep.so
for v in current base noreuse; do
  [ $v != current ] && cp omok-ai_b200/_build/variants/$v.so omok-ai_b200/libomok_b200.so
  for rep in 1 2; do split > $O/r02c1_split_${v}_$rep.json; done
  split --lanes 1 > $O/r02c1_split2l_${v}.json
  python tools/profile_step.py --games 256 --plies 2 --warm 2 --lanes 1 > $O/r02c1_split_256g_${v}.json 2>/dev/null
  python tools/profile_step.py --games 512 --plies 2 --warm 2 --lanes 1 > $O/r02c1_split_512g_${v}.json 2>/dev/null
  cp /tmp/keep.so omok-ai_b200/libomok_b200.so
done
python tools/profile_step.py --tree-sweep > $O/r02c1_tree_sweep.json 2>/dev/null
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02c1_split*.json')):
    try:
        d=json.load(open(f)); k=d['kinds']
        print(f.split('/')[-1], 'sims/s %.3fM'%(d['sims_per_s']/1e6), ' '.join('%s %.4f'%(n, k[n]['ms']/max(1,k[n]['launches'])) for n in ('tower','fc0','fc1','heads','select_expand','apply')))
    except Exception as e: print(f, 'ERR', e)
PY
timeout 600 python tests/tools/check_f16.py 300 > $O/r02c1_check_f16.log 2>&1; tail -32 $O/r02c1_check_f16.log
# network error study: full sample on the default build, reduced sample on the chunk variants
timeout 1500 python tools/net_error_study.py --positions 20000 --label default_c9_f8_h8 --out $O/r02c1_err_default.json 2>&1 | tail -12
for v in c6 c3 h2 c3h1; do
  cp omok-ai_b200/_build/variants/$v.so omok-ai_b200/libomok_b200.so
  timeout 600 python tools/net_error_study.py --positions 6000 --seeds 0 --paths tc --label $v --out $O/r02c1_err_$v.json 2>&1 | tail -3
  split > $O/r02c1_split_$v.json
  python -c "
import json; d=json.load(open('$O/r02c1_split_$v.json')); k=d['kinds']
print('$v', 'sims/s %.3fM'%(d['sims_per_s']/1e6), ' '.join('%s %.4f'%(n, k[n]['ms']/max(1,k[n]['launches'])) for n in ('tower','fc0','fc1','heads')))"
  cp /tmp/keep.so omok-ai_b200/libomok_b200.so
done
