#!/usr/bin/env bash
set -u
OUT=gpurun_out/r2_call8
mkdir -p "$OUT"
run() { local name=$1; shift; echo "=== $name: $*"; ( timeout "${T:-300}" "$@" ) > "$OUT/$name.log" 2>&1; echo "exit $? ($name)"; tail -n "${TAILN:-4}" "$OUT/$name.log"; }
TAILN=15 T=600 run graph_test python -m pytest tests/test_graph_gpu.py tests/test_geometry_gpu.py tests/test_ckpt_gpu.py -q -x
TAILN=3 T=900 run bench_graph python bench.py --steps 10 --warmup 3 --no-cpu-baseline
TAILN=3 T=900 run bench_nograph python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-graph --skip-configs
T=400 run ncu_imgatt ncu --set full --clock-control none --import-source on -k regex:image_attention -c 2 -o "$OUT/imgatt_stream" -f python tools/imgatt_ab.py
