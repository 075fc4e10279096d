#!/bin/bash
# Round-end evidence on ONE GPU: GPU test suite, smoke, final bench lines (both arms), C-ABI and pull probes, ping-pong,
# ncu launch lists of the bench, one full ncu capture of the pull kernel (a launch per batch shape).
mkdir -p gpurun_out; O=gpurun_out
timeout -k 10 500 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -3 | tee $O/r02_pytest_gpu_final.log
timeout -k 10 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout -k 10 400 python bench.py > $O/r02_bench_n1_final.json 2> $O/ev_err1.txt
timeout -k 10 300 python bench.py --impl reference --steps 20 --warmup 3 > $O/r02_bench_ref_n1_final.json 2> $O/ev_err2.txt
python - <<'P'
import json
d=json.load(open("gpurun_out/r02_bench_n1_final.json")); print(d["value"],d["ms_per_step"],d["e2e"]["value"],d["e2e_pageable"]["value"],d["roofline"]["frac"],d["gpu_launches"],d["cpu_baseline"]["value"])
print([(p["bytes"],p["gbs_per_gpu"],p["mmsg_per_s_per_gpu"],p["frac_of_roofline"],p["bit_exact"]) for p in d["sweep"]["points"]])
try:
    r=json.load(open("gpurun_out/r02_bench_ref_n1_final.json")); print("ref",r["value"],r["ms_per_step"])
except Exception as e: print("ref ERR",e)
P
timeout -k 10 60 tests/gpu_probe/abi_bench > $O/r02_abi_bench_final.txt 2>&1; grep -c "c-abi" $O/r02_abi_bench_final.txt
timeout -k 10 60 python bench_scenarios.py pingpong --out $O/r02_pingpong_n1_final.jsonl > $O/pp.txt 2>&1; cut -c1-200 $O/pp.txt | head -5
timeout -k 10 60 tests/gpu_probe/sw_probe pull - > $O/r02_probe_pull_final.txt 2>&1; tail -3 $O/r02_probe_pull_final.txt
STARWAY_RESIDENT=0 timeout -k 10 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/r02_ncu_launches_resident0.csv python bench.py --steps 3 --warmup 3 --no-sweep --no-e2e --no-cpu-baseline > $O/ncu_b0.txt 2>&1; echo "ncu resident0 rc=$?"
timeout -k 10 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/r02_ncu_launches_resident1.csv python bench.py --steps 3 --warmup 3 --no-sweep --no-e2e --no-cpu-baseline > $O/ncu_b1.txt 2>&1; echo "ncu resident1 rc=$?"
SW_PROBE_PULL_LINGER_US=1 timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:sw_pull_kernel -c 9 -f -o $O/r02_ncu_pull tests/gpu_probe/sw_probe pull - 1 > $O/ncu_pull.txt 2>&1; echo "ncu pull rc=$?"
