#!/bin/bash
# On the GPU box: for each variant built by ab_build.sh, copy it over libomok_b200.so, run the network parity tests and
# the one-lane kernel time split; prints one line per variant.  Usage: tests/tools/ab_run.sh NAME...
cd "$(dirname "$0")/../.."
cp omok-ai_b200/libomok_b200.so /tmp/libomok_b200.keep
for v in "$@"; do
  cp omok-ai_b200/_build/variants/$v.so omok-ai_b200/libomok_b200.so
  t=$(python -m pytest tests/test_net_gpu.py -x -q 2>&1 | tail -1)
  for rep in 1 2; do
    python tools/profile_step.py --games 1024 --plies 1 --warm 1 2>/dev/null | python -c "
import json,sys
d=json.load(sys.stdin); k=d['kinds']
print('$v', 'tests:', '$t', '| tower %.4f fc0 %.4f fc1 %.4f heads %.4f ms/launch | sims/s %.3fM' % (k['tower']['ms']/k['tower']['launches'], k['fc0']['ms']/k['fc0']['launches'], k['fc1']['ms']/k['fc1']['launches'], k['heads']['ms']/k['heads']['launches'], d['sims_per_s']/1e6))"
  done
done
cp /tmp/libomok_b200.keep omok-ai_b200/libomok_b200.so
