#!/bin/bash
# Multi-GPU validation on one box:  bash tools/multi_gpu_validation.sh <N> [full]
# bench (all blocks), bench with the other dense path, correctness matrix, per-path roofline;
# "full" adds the all-reduce sweep and the example families.  Logs go to gpurun_out/.
N=$1
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
S='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(sys.argv[1], d["ms_per_step"], d["loss"], d.get("comm"))'
timeout -k 5 420 $T --master-port 29511 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/b$N.log 2> gpurun_out/b$N.err; echo bench_rc=$?; tail -c 1500 gpurun_out/b$N.log | head -c 600; echo
python -c "$S" N$N < gpurun_out/b$N.log
timeout -k 5 200 $T --master-port 29516 bench.py --gpus $N --steps 50 --warmup 5 --no-extras --no-e2e --comm --dense-nvls off 2>/dev/null | python -c "$S" N${N}_p2p_dense
timeout -k 5 420 $T --master-port 29512 tests/mp_nvlink_worker.py --quick > gpurun_out/matrix$N.log 2>&1; echo matrix_rc=$?; grep -c " OK" gpurun_out/matrix$N.log; grep -E "FAIL|ALL OK|SOME|Error" gpurun_out/matrix$N.log | head -10
timeout -k 5 200 $T --master-port 29513 tools/fabric_roofline.py --out gpurun_out/fabric_roofline_$N.json > gpurun_out/roof$N.log 2>&1; echo roof_rc=$?
if [ "$2" == "full" ]; then
timeout -k 5 240 $T --master-port 29514 tools/allreduce_sweep.py --blocks 128 --iters 10 --out gpurun_out/allreduce_sweep_${N}gpu.json > gpurun_out/sweep$N.log 2>&1; echo sweep_rc=$?; tail -4 gpurun_out/sweep$N.log
timeout -k 5 300 $T --master-port 29515 tests/mp_examples_worker.py > gpurun_out/ex$N.log 2>&1; echo ex_rc=$?; grep -E "OK|FAIL" gpurun_out/ex$N.log | tail -6
fi
