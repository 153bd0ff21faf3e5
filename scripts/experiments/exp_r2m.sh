python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --no-cpu --stack-pairs 64 --volume-pairs 8 > gpurun_out/bench_r2m.json 2> gpurun_out/bench_r2m.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/bench_r2m.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel_ms_per_step'], d['stack']['value'], d['volume']['value'], d['clocks'])
"
