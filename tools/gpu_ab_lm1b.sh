#!/bin/bash
# One-GPU check of the LM1B hot path: fused-op numerics, the bench step with and without the
# second side stream for the bias gradient, and a per-stream kernel timeline of one replayed step.
# Run on a B200 box: /usr/local/graft/bin/gpurun --timeout 420 -- bash tools/gpu_ab_lm1b.sh
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_fused.py -x -q > gpurun_out/t_fused.log 2>&1; echo "fused_rc=$?"
tail -5 gpurun_out/t_fused.log
timeout 120 python bench.py --no-extras --no-e2e --steps 100 --warmup 10 > gpurun_out/bn_new.json 2> gpurun_out/bn_new.err; echo "new_rc=$?"
PARALLAX_LSTM_DBIAS_STREAM=0 timeout 120 python bench.py --no-extras --no-e2e --steps 100 --warmup 10 > gpurun_out/bn_nodb.json 2> gpurun_out/bn_nodb.err; echo "nodb_rc=$?"
timeout 100 python tools/profile_step.py --graph --trace gpurun_out/trace_lm1b_r2b.txt --out gpurun_out/profile_step_r2b.txt > gpurun_out/prof.log 2>&1; echo "prof_rc=$?"
python - <<'PY'
import json
for f in ("bn_new","bn_nodb"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["value"], d.get("gpu_launches"))
    except Exception as e:
        print(f, "ERR", e)
PY
tail -3 gpurun_out/bn_new.err
