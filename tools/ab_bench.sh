# usage: bash tools/ab_bench.sh "ENV=VAL ... [-- bench args]" ...: bench.py kernel table once per setting (no CPU baseline)
for spec in "$@"; do
  echo "== $spec"
  envs="${spec%%--*}"; args=""
  case "$spec" in *--*) args="--${spec#*--}";; esac
  env $envs python bench.py --no-cpu $args > gpurun_out/ab.json 2> gpurun_out/ab.err || tail -5 gpurun_out/ab.err
  python -c "
import json; d=json.load(open('gpurun_out/ab.json')); print(round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value'])); [print(' ', k['kernel'], k['launches'], k['ms']) for k in d['kernels'][:4]]"
done
