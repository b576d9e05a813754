for ACC in 0 4; do
  VBT_PW_PERSIST=1 VBT_PW_ACC=$ACC timeout 300 python bench.py --steps 60 --warmup 6 --no-e2e --no-cpu-baseline --op-dump gpurun_out/r71_acc${ACC}_ops.tsv > gpurun_out/r71_acc$ACC.json 2> gpurun_out/r71_acc$ACC.err
  python - <<PY
import csv, json
d=json.load(open('gpurun_out/r71_acc$ACC.json'))
a=list(csv.DictReader(open('gpurun_out/r71_acc${ACC}_ops.tsv'),delimiter='\t'))
print('ACC=$ACC value',round(d['value']),'pw',round(d['kernels']['pw']['ms_per_step'],3), [(x['name'],x['us_per_call']) for x in a[:16] if x['kernel']=='pw'])
PY
done
