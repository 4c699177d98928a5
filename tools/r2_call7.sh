#!/usr/bin/env bash
set -u
OUT=gpurun_out/r2_call7
mkdir -p "$OUT"
run() { local name=$1; shift; echo "=== $name: $*"; ( timeout "${T:-300}" "$@" ) > "$OUT/$name.log" 2>&1; echo "exit $? ($name)"; tail -n "${TAILN:-4}" "$OUT/$name.log"; }
TAILN=12 T=300 run imgatt_ab python tools/imgatt_ab.py
T=600 run track_tests python -m pytest tests/test_track_gpu.py tests/test_windowed_gpu.py tests/test_ckpt_gpu.py -q
TAILN=8 T=600 run e2e_diag python tools/e2e_diag.py
TAILN=3 T=900 run bench_n1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
