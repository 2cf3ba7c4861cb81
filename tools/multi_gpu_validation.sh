N=$1
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout -k 5 420 $T --master-port 29511 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/b$N.log 2> gpurun_out/b$N.err; echo bench_rc=$?; tail -c 2500 gpurun_out/b$N.log
timeout -k 5 420 $T --master-port 29512 tests/mp_nvlink_worker.py --quick > gpurun_out/matrix$N.log 2>&1; echo matrix_rc=$?; grep -c " OK" gpurun_out/matrix$N.log; grep -E "FAIL|ALL OK|SOME|Error|error" gpurun_out/matrix$N.log | head -10
timeout -k 5 200 $T --master-port 29513 tools/fabric_roofline.py --out gpurun_out/fabric_roofline_$N.json > gpurun_out/roof$N.log 2>&1; echo roof_rc=$?
timeout -k 5 240 $T --master-port 29514 tools/allreduce_sweep.py --blocks 128 --iters 10 --out gpurun_out/allreduce_sweep_${N}gpu.json > gpurun_out/sweep$N.log 2>&1; echo sweep_rc=$?; tail -4 gpurun_out/sweep$N.log
timeout -k 5 300 $T --master-port 29515 tests/mp_examples_worker.py > gpurun_out/ex$N.log 2>&1; echo ex_rc=$?; grep -E "OK|FAIL" gpurun_out/ex$N.log | tail -6
