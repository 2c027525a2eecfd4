#!/bin/bash
# usage: tools/gpu_cycle.sh TAG [bench args...]   -- GPU tests + one bench run on a B200 via gpurun
set -u
TAG=${1:-cyc}; shift
cd /root/repo
make -C /root/repo/jaxsplat_b200/csrc -j8 2>&1 | grep -i "error" && exit 1
/usr/local/graft/bin/gpurun --timeout 1200 -- "python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${TAG}_pytest.log; python bench.py --steps 100 --warmup 10 --no-cpu-baseline $* > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -4 gpurun_out/${TAG}_pytest.log; tail -3 gpurun_out/${TAG}_bench.err" 2>&1 | tail -7
python - <<EOF
import json
try:
    d=json.load(open('/root/repo/gpurun_out/${TAG}_bench.json'))
    print('ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'I', d['config']['num_intersects'])
    print(' | '.join(f"{k} {v['ms_per_launch']*1000:.1f}" for k,v in d['stages'].items()))
except Exception as e: print('no bench json', e)
EOF
