mkdir -p gpurun_out; O=gpurun_out
for k in 1 0; do
  rm -f /tmp/tr_k$k*
  STARWAY_PULL_KEEP=$k ABI_MIN=268435456 STARWAY_TRACE=/tmp/tr_k$k timeout -k 10 60 tests/gpu_probe/abi_bench 2>&1 | grep "c-abi" | cut -c18-300
  python tests/tools/trace_tail.py /tmp/tr_k$k* -n 70 > $O/r02_trace_256m_keep$k.txt; 
done
for k in 1 0; do
  STARWAY_PULL_KEEP=$k timeout -k 10 120 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline > $O/bk$k.json 2>$O/err.txt; python -c "
import json;d=json.load(open('gpurun_out/bk$k.json'));print('keep=$k',d['value'],d['ms_per_step']);print([(p['bytes'],p['gbs_per_gpu']) for p in d['sweep']['points']])"
done
