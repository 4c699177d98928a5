#!/usr/bin/env bash
# First GPU call of round 2 (one `gpurun --timeout 2400 -- bash tools/r2_first_call.sh`): regression + the experiments that
# were prepared on CPU at the end of round 1. Everything lands in gpurun_out/r2_first/. Each step has its own timeout; all
# device-side waits are bounded (mbar_wait traps), so a protocol bug in an experimental kernel ends as a CUDA error.
set -u
OUT=gpurun_out/r2_first
mkdir -p "$OUT"
run() { local name=$1; shift; echo "=== $name: $*"; ( timeout "${T:-600}" "$@" ) > "$OUT/$name.log" 2>&1; echo "exit $? ($name)"; tail -n 5 "$OUT/$name.log"; }

T=900 run pytest_gpu python -m pytest tests -m gpu -x -q
T=600 run bench_n1 python bench.py --steps 10 --warmup 3
# 1. CTA-pair attention kernel (csrc/attention_pair.cu): parity + back-to-back timing of both arms
T=300 run att_pair_ab python tools/att_pair_ab.py
# 2. if the pair kernel is parity-green: whole-step effect
T=600 run bench_pair env L4P_ATT_PAIR=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
# 2a. ncu --set full of both attention arms (source-level stalls: --import-source on; the build has -lineinfo)
T=600 run ncu_att_default ncu --set full --clock-control none --import-source on --profile-from-start off -o "$OUT/att_default" -f python tools/att_ncu.py
T=600 run ncu_att_pair env L4P_ATT_PAIR=1 ncu --set full --clock-control none --import-source on --profile-from-start off -o "$OUT/att_pair" -f python tools/att_ncu.py
# 2b. track head with a 16-bit per-query token stream (numerically equivalent on CPU vs the live reference: DESIGN.md §7.4)
T=600 run track_res16_tests env L4P_TRACK_RES16=1 python -m pytest tests/test_track_gpu.py tests/test_windowed_gpu.py -x -q
T=600 run bench_res16 env L4P_TRACK_RES16=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
# 3. timeline of the production kernel (per key block: wait_s / ldtm / max / exp / wait_pv / store, MMA thread waits)
T=300 run att_timeline python tools/att_prof.py
# 4. S-first issue order (build-time variant), then restore the default build
T=600 run build_sfirst env L4P_NVCC_EXTRA=-DL4P_ATT_S_FIRST=1 python -m l4p_b200.build
T=300 run att_sfirst env L4P_NVCC_EXTRA=-DL4P_ATT_S_FIRST=1 python tools/att_prof.py
# 5. FA4-style P-aliases-S variant (both tiles TS-mode, no P staging in smem): parity of the attention tests, then timing
T=600 run build_palias env L4P_NVCC_EXTRA=-DL4P_ATT_P_ALIAS=1 python -m l4p_b200.build
T=300 run att_palias_parity env L4P_NVCC_EXTRA=-DL4P_ATT_P_ALIAS=1 python -m pytest tests/test_gemm_gpu.py -q -k attention
T=300 run att_palias env L4P_NVCC_EXTRA=-DL4P_ATT_P_ALIAS=1 python tools/att_prof.py
T=300 run att_palias_pair_ab env L4P_NVCC_EXTRA=-DL4P_ATT_P_ALIAS=1 python tools/att_pair_ab.py   # both arms built with P alias
# 6. softmax denominator from the PV UMMA (ones row in the V^T pad): needs the build macro AND the run-time env
T=600 run build_lsum env L4P_NVCC_EXTRA=-DL4P_ATT_LSUM_MMA=1 python -m l4p_b200.build
T=300 run att_lsum_parity env L4P_NVCC_EXTRA=-DL4P_ATT_LSUM_MMA=1 L4P_ATT_LSUM_MMA=1 python -m pytest tests/test_gemm_gpu.py -q -k attention
T=300 run att_lsum env L4P_NVCC_EXTRA=-DL4P_ATT_LSUM_MMA=1 L4P_ATT_LSUM_MMA=1 python tools/att_prof.py
T=600 run build_default python -m l4p_b200.build
