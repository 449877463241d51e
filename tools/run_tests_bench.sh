python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py > gpurun_out/bench_cur.json 2> gpurun_out/bench_cur.err
python -c "
import json; d=json.load(open('gpurun_out/bench_cur.json')); print(d['value'], d['ms_per_step'], d['e2e']['value']); [print(k) for k in d['kernels']]"
