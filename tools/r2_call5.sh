#!/usr/bin/env bash
# Round-2 GPU call 5: full GPU suite (new parity tests), bench.
set -u
OUT=gpurun_out/r2_call5
mkdir -p "$OUT"
run() { local name=$1; shift; echo "=== $name: $*"; ( timeout "${T:-300}" "$@" ) > "$OUT/$name.log" 2>&1; echo "exit $? ($name)"; tail -n "${TAILN:-4}" "$OUT/$name.log"; }
TAILN=40 T=1500 run pytest_gpu python -m pytest tests -m gpu -q -s -rA
TAILN=3 T=900 run bench_n1 python bench.py --steps 10 --warmup 3
