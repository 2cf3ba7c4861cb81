#!/bin/bash
# One-call hardware validation of everything that so far has only CPU / simulated coverage.
# Every section is time-boxed and writes its own log under gpurun_out/ (merged back by gpurun).
#
#   gpurun --timeout 900          -- bash tools/gpu_validation.sh single
#   gpurun --gpus 4 --timeout 600 -- bash tools/gpu_validation.sh multi 4
#   gpurun --gpus 8 --timeout 900 -- bash tools/gpu_validation.sh multi 8
#
# single : pytest -m gpu, smoke, LM1B / ResNet bench at N=1, launch list, one ncu capture
# multi N: IPC-fabric worker test, LM1B bench (P2P and, for N>=4, forced NVLS), all-reduce sweep,
#          library (NCCL) fabric, the NMT example across N ranks
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH=$PWD
MODE=${1:-single}
N=${2:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run() { # name, seconds, command...
  local name=$1 secs=$2; shift 2
  echo "=== $name"
  timeout "$secs" "$@" > "gpurun_out/val_$name.log" 2>&1
  echo "rc=$? ($name)"; tail -3 "gpurun_out/val_$name.log"
}
python -m parallax_b200.ops.build > gpurun_out/val_build.log 2>&1 || { echo "build failed"; exit 1; }
if [ "$MODE" = single ]; then
  run pytest_gpu 420 python -m pytest tests -x -q -m gpu
  run smoke 120 python -c "import __graft_entry__ as g; g.smoke()"
  run bench_lm1b_1 180 python bench.py --gpus 1 --steps 30 --warmup 5
  run bench_resnet_1 240 python bench.py --gpus 1 --steps 20 --warmup 5 --model resnet50
  run launches 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
      --log-file gpurun_out/val_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e
  run ncu_top 300 ncu --set full --clock-control none --import-source on -k regex:px_ -c 30 \
      -o gpurun_out/val_prof python tools/ncu_targets.py
else
  run mp_worker_$N 240 $TR --master-port 29541 tests/mp_nvlink_worker.py
  run bench_lm1b_$N 240 $TR --master-port 29542 bench.py --gpus "$N" --steps 30 --warmup 5
  if [ "$N" -ge 4 ]; then
    run bench_lm1b_nvls_$N 240 $TR --master-port 29546 bench.py --gpus "$N" --steps 30 --warmup 5 \
        --dense-nvls on
  fi
  run bench_resnet_$N 300 $TR --master-port 29543 bench.py --gpus "$N" --steps 20 --warmup 5 \
      --model resnet50
  run sweep_$N 300 $TR --master-port 29544 tools/allreduce_sweep.py
  PARALLAX_FABRIC=library run library_fabric_$N 240 $TR --master-port 29545 bench.py --gpus "$N" \
      --steps 10 --warmup 3 --small --no-graph
  run nmt_$N 300 python examples/nmt/nmt_distributed_driver.py --synthetic \
      --out_dir gpurun_out/val_nmt_$N --num_train_steps 60 --steps_per_eval 1000 \
      --compute_dtype bf16 --hparams num_units=256,num_layers=2,batch_size=64,steps_per_stats=20 \
      --resource_info_file "localhost:$(seq -s, 0 $((N-1)))"
fi
echo "=== done; logs: gpurun_out/val_*.log"
