#!/usr/bin/env bash
set -u
OUT=gpurun_out/r2_call9
mkdir -p "$OUT"
run() { local name=$1; shift; echo "=== $name: $*"; ( timeout "${T:-300}" "$@" ) > "$OUT/$name.log" 2>&1; echo "exit $? ($name)"; tail -n "${TAILN:-4}" "$OUT/$name.log"; }
TAILN=14 T=300 run imgatt_ab python tools/imgatt_ab.py
TAILN=15 T=1500 run pytest_gpu python -m pytest tests -m gpu -q
TAILN=3 T=900 run bench_graph python bench.py --steps 10 --warmup 3 --no-cpu-baseline --skip-configs
T=400 run ncu_imgatt env L4P_IMGATT_STREAM=2 ncu --set full --clock-control none --import-source on -k regex:image_attention -c 2 -o "$OUT/imgatt_mma" -f python tools/imgatt_ab.py 2
