# convenience targets (everything is also reachable through plain python)
PY ?= python

build:
	$(PY) -m parallax_b200.ops.build

test-cpu: build
	$(PY) -m pytest tests -x -q -m "not gpu"

test-gpu: build
	$(PY) -m pytest tests -x -q -m gpu

test-multigpu: build
	$(PY) -m torch.distributed.run --nnodes=1 --nproc-per-node $${N:-2} --master-addr 127.0.0.1 \
	    --master-port 29533 tests/mp_nvlink_worker.py

bench: build
	$(PY) bench.py --gpus 1 --steps 20 --warmup 5

style:
	$(PY) tools/style_check.py

.PHONY: build test-cpu test-gpu test-multigpu bench style
