mkdir -p gpurun_out; O=gpurun_out
SW_PROBE_PULL_LINGER_US=1 timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:sw_pull_kernel -c 9 -f -o $O/r02_ncu_pull tests/gpu_probe/sw_probe pull - 1 > $O/ncu_pull.txt 2>&1; echo "ncu pull rc=$?"; grep "^pull" $O/ncu_pull.txt
STARWAY_TRACE=/tmp/tr64 timeout -k 10 100 python tests/tools/latency_trace.py --bytes 64 > $O/r02_latency_trace_64.txt 2>&1; cat $O/r02_latency_trace_64.txt
STARWAY_TRACE=/tmp/tr1m timeout -k 10 100 python tests/tools/latency_trace.py --bytes 1048576 > $O/r02_latency_trace_1m.txt 2>&1; cat $O/r02_latency_trace_1m.txt
