#!/bin/bash
# what the driver runs at round end, on one GPU: the GPU test suite, smoke(), both bench arms
cd "$(dirname "$0")/../.."
O=gpurun_out
( timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -5 ) > $O/final_pytest.log; cat $O/final_pytest.log
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2
python bench.py --steps 5 --warmup 3 > $O/final_bench.json 2> $O/final_bench.err; python -c "
import json; d=json.load(open('$O/final_bench.json')); print('bench', d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'], d['roofline']['avg_launch_ms'], d['roofline_second']['avg_launch_ms'])"
python bench.py --impl reference --steps 5 --warmup 3 > $O/final_bench_ref.json 2>/dev/null; python -c "
import json; d=json.load(open('$O/final_bench_ref.json')); print('ref', d['value'], d['steps'], d['warmup'])"
