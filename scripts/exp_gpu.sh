cd "${GRAFT_REPO_ROOT:-.}"
python __graft_entry__.py > gpurun_out/build_exp.log 2>&1 || { tail gpurun_out/build_exp.log; exit 1; }
run() { tag=$1; shift; env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-pipeline > gpurun_out/exp_$tag.json 2> gpurun_out/exp_$tag.err; python - $tag <<'PY'
import json,sys
tag=sys.argv[1]
try:
    l=json.loads(open(f'gpurun_out/exp_{tag}.json').readline()); r=l['roofline']
    print(tag, 'step_ms %.3f'%l['ms_per_step'], 'pileup_ms %.3f'%r['kernel_ms_per_step'], 'all_kernels_ms %.3f'%r['both_kernels_ms_per_step'], 'frac %.3f'%r['frac'], 'open', l['config'].get('open_columns'), 'e2e %.3e'%l['e2e']['value'])
except Exception as e:
    print(tag,'FAILED',e); print(open(f'gpurun_out/exp_{tag}.err').read()[-800:])
PY
}
for cfg in "$@"; do
  set -- $cfg
  tag=$1; shift
  run $tag "$@"
done
