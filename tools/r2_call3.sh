#!/usr/bin/env bash
# Round-2 GPU call 3: split-softmax attention kernel (parity, A/B, POLY sweep, timeline) + the new bench.py legs.
set -u
OUT=gpurun_out/r2_call3
mkdir -p "$OUT"
run() { local name=$1; shift; echo "=== $name: $*"; ( timeout "${T:-300}" "$@" ) > "$OUT/$name.log" 2>&1; echo "exit $? ($name)"; tail -n "${TAILN:-4}" "$OUT/$name.log"; }
T=300 run att_parity python -m pytest tests/test_gemm_gpu.py -q -k attention
TAILN=16 T=300 run att_split_ab python tools/att_ab.py L4P_ATT_SPLIT 0 1
TAILN=30 T=400 run att_poly_ab python tools/att_ab.py L4P_ATT_POLY 0 1 2 3
TAILN=60 T=300 run att_timeline python tools/att_prof.py
T=900 run pytest_gpu python -m pytest tests -m gpu -x -q
TAILN=3 T=900 run bench_ref python bench.py --impl reference --steps 3 --warmup 1
TAILN=3 T=900 run bench_n1 python bench.py --steps 10 --warmup 3
