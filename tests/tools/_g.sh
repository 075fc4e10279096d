mkdir -p gpurun_out; O=gpurun_out
run() { env STARWAY_OPTS="$1" timeout -k 10 120 python bench.py --steps 20 --warmup 5 --no-sweep --no-cpu-baseline > $O/be.json 2>$O/err.txt; python -c "
import json;d=json.load(open('gpurun_out/be.json'));e=d['e2e'];print('$1', '->', e['value'], e['rank0_breakdown']['ms_per_step'], e['rank0_breakdown']['copy_kernel_ms_per_step'], e['rank0_breakdown']['bulk_launches_per_step'], 'pageable', d['e2e_pageable']['value'])" || tail -3 $O/err.txt; }
run "pinned_send_direct=0,hostdst_ce=1"
run "pinned_send_direct=0,hostdst_ce=1,stage_batch_bytes=8388608"
run "pinned_send_direct=0,hostdst_ce=1,stage_batch_bytes=2097152"
run "pinned_send_direct=0,hostdst_ce=1,coalesce_us=10"
STARWAY_OPTS="pinned_send_direct=0,hostdst_ce=1" timeout -k 10 300 python -m pytest tests -m gpu -x -q --timeout 280 -k "host or pageable or pinned or mixed or numpy or e2e or flush" 2>&1 | tail -2
