mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
nproc > gpurun_out/r02_n8_host.txt; cat /sys/fs/cgroup/cpu.max >> gpurun_out/r02_n8_host.txt 2>&1; lscpu | grep -i "numa\|^CPU(s)\|Model name" >> gpurun_out/r02_n8_host.txt; nvidia-smi topo -m >> gpurun_out/r02_n8_host.txt 2>&1
timeout 200 $TR --master-port 29521 bench.py --gpus 8 --steps 60 --warmup 10 > gpurun_out/r02_bench_n8_v3.json 2> gpurun_out/err8a.txt
STARWAY_BENCH_NO_BIND=1 timeout 100 $TR --master-port 29522 bench.py --gpus 8 --steps 60 --warmup 10 --no-sweep --no-e2e --no-cpu-baseline > gpurun_out/r02_bench_n8_v3_nobind.json 2> gpurun_out/err8b.txt
STARWAY_RESIDENT=0 timeout 100 $TR --master-port 29523 bench.py --gpus 8 --steps 60 --warmup 10 --no-sweep --no-e2e --no-cpu-baseline > gpurun_out/r02_bench_n8_v3_resident0.json 2> gpurun_out/err8c.txt
python - <<'P'
import json
for f in ["r02_bench_n8_v3","r02_bench_n8_v3_nobind","r02_bench_n8_v3_resident0"]:
    try:
        d=json.load(open("gpurun_out/"+f+".json")); print(f,d["value"],d["ms_per_step"],d["e2e"]["value"],d["config"].get("per_rank_ms_per_step_and_cpus"),d["roofline"].get("batch_phases_us"))
    except Exception as e: print(f,"ERR",e)
P
timeout 150 python -m pytest tests/test_gpu_multi.py -x -q --timeout 140 2>&1 | tail -4 > gpurun_out/r02_pytest_gpu_multi_8_v3.log; cat gpurun_out/r02_pytest_gpu_multi_8_v3.log
timeout 60 $TR --master-port 29524 bench_scenarios.py allpairs --rounds 20 --out gpurun_out/r02_allpairs_n8_v3.json > gpurun_out/ap8.txt 2>&1; tail -3 gpurun_out/ap8.txt
