#!/usr/bin/env bash
# Round-2 GPU call 1: regression + bench + A/B of the attention variants prepared at the end of round 1.
# Variant libraries are prebuilt on CPU (L4P_BUILD_TAG=<tag> python -m l4p_b200.build) and selected with L4P_LIB.
set -u
OUT=gpurun_out/r2_call1
mkdir -p "$OUT"
run() { local name=$1; shift; echo "=== $name: $*"; ( timeout "${T:-300}" "$@" ) > "$OUT/$name.log" 2>&1; echo "exit $? ($name)"; tail -n "${TAILN:-4}" "$OUT/$name.log"; }
P=l4p_b200
T=900 run pytest_gpu python -m pytest tests -m gpu -x -q
T=400 run bench_n1 python bench.py --steps 10 --warmup 3
TAILN=16 T=300 run att_pair_ab python tools/att_pair_ab.py
TAILN=16 T=300 run att_palias_ab env L4P_LIB=$P/libl4p_b200_palias.so python tools/att_pair_ab.py
TAILN=16 T=300 run att_sfirst_ab env L4P_LIB=$P/libl4p_b200_sfirst.so python tools/att_pair_ab.py
TAILN=16 T=300 run att_lsum_ab env L4P_LIB=$P/libl4p_b200_lsum.so L4P_ATT_LSUM_MMA=1 python tools/att_pair_ab.py
T=300 run att_palias_parity env L4P_LIB=$P/libl4p_b200_palias.so python -m pytest tests/test_gemm_gpu.py -q -k attention
T=300 run att_timeline python tools/att_prof.py
T=300 run att_timeline_palias env L4P_LIB=$P/libl4p_b200_palias.so python tools/att_prof.py
T=400 run ncu_att_default ncu --set full --clock-control none --import-source on --profile-from-start off -o "$OUT/att_default" -f python tools/att_ncu.py
T=400 run ncu_att_palias env L4P_LIB=$P/libl4p_b200_palias.so ncu --set full --clock-control none --import-source on --profile-from-start off -o "$OUT/att_palias" -f python tools/att_ncu.py
T=400 run ncu_att_pair env L4P_ATT_PAIR=1 ncu --set full --clock-control none --import-source on --profile-from-start off -o "$OUT/att_pair" -f python tools/att_ncu.py
T=400 run ncu_att_palias_pair env L4P_LIB=$P/libl4p_b200_palias.so L4P_ATT_PAIR=1 ncu --set full --clock-control none --import-source on --profile-from-start off -o "$OUT/att_palias_pair" -f python tools/att_ncu.py
T=400 run track_res16_tests env L4P_TRACK_RES16=1 python -m pytest tests/test_track_gpu.py tests/test_windowed_gpu.py -x -q
T=300 run bench_res16 env L4P_TRACK_RES16=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
ls /root/reference > "$OUT/ref_present.log" 2>&1; nproc >> "$OUT/ref_present.log"; free -g >> "$OUT/ref_present.log"
