#!/usr/bin/env bash
set -u
OUT=gpurun_out/r2_call10
mkdir -p "$OUT"
run() { local name=$1; shift; echo "=== $name: $*"; ( timeout "${T:-300}" "$@" ) > "$OUT/$name.log" 2>&1; echo "exit $? ($name)"; tail -n "${TAILN:-4}" "$OUT/$name.log"; }
TAILN=8 T=300 run imgatt_ab python tools/imgatt_ab.py
TAILN=6 T=900 run tests python -m pytest tests/test_track_gpu.py tests/test_ckpt_gpu.py tests/test_graph_gpu.py tests/test_e2e_gpu.py tests/test_geometry_gpu.py -q
TAILN=60 T=600 run step_breakdown env L4P_SERIAL_HEADS=1 python tools/profile_step.py
TAILN=3 T=900 run bench_graph python bench.py --steps 20 --warmup 5 --no-cpu-baseline
