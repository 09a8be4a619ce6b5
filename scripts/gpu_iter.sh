# one GPU iteration of the rows-kernel work: parity tests, in-kernel stage clocks, short bench.  usage: gpu_iter.sh TAG [dbgflags...]
cd /root/repo
TAG=$1; shift
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_gpu_$TAG.log; cat gpurun_out/pytest_gpu_$TAG.log
timeout 120 python scripts/stage_clock.py 18944 > gpurun_out/stage_clock_$TAG.txt 2>&1; head -21 gpurun_out/stage_clock_$TAG.txt
for f in "$@"; do echo "== RB_DBG=$f"; RB_DBG=$f timeout 120 python scripts/stage_clock.py 18944 2>&1 | head -21 | tee gpurun_out/stage_clock_${TAG}_dbg$f.txt; done
timeout 300 python bench.py --no-extras --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; python -c "
import json
for l in open('gpurun_out/bench_$TAG.json'):
    if l.startswith('{'):
        d=json.loads(l); print('ms_per_step',d['ms_per_step'],'kernel_ms',d['roofline']['kernel_ms'],'frac',d['roofline']['frac'],'e2e',d['e2e']['ms_per_step'],d['clocks'])
"
