#!/bin/bash
# A/B of fc0's tail balancing on one box: parity tests, then the bench and the one-lane split with it on and off
cd "$(dirname "$0")/../.."
timeout 900 python -m pytest tests/test_net_gpu.py -x -q 2>&1 | tail -3
for i in 1 2; do
for v in 0 1; do OMK_FC0_BALANCE=$v python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('balance=$v', d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], 'fc0 one-lane', d['roofline_second']['avg_launch_ms'])"; done; done
for v in 0 1; do OMK_FC0_BALANCE=$v python tools/profile_step.py --games 512 --plies 2 --warm 2 --lanes 1 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); k=d['kinds']; print('512 games balance=$v sims/s %.2fM fc0 %.4f'%(d['sims_per_s']/1e6, k['fc0']['ms']/k['fc0']['launches']))"; done
for v in 0 1; do OMK_FC0_BALANCE=$v python tools/profile_step.py --games 256 --plies 2 --warm 2 --lanes 1 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); k=d['kinds']; print('256 games balance=$v sims/s %.2fM fc0 %.4f'%(d['sims_per_s']/1e6, k['fc0']['ms']/k['fc0']['launches']))"; done
