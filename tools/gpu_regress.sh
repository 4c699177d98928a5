#!/usr/bin/env bash
# One gpurun call that regenerates the evidence under profiles/: full GPU suite, both bench arms, the ncu launch list of one
# step, a fresh ncu --set full capture of the attention kernel, smoke().   gpurun --timeout 3000 -- bash tools/gpu_regress.sh
set -u
OUT=gpurun_out/regress
mkdir -p "$OUT"
run() { local name=$1; shift; echo "=== $name: $*"; ( timeout "${T:-300}" "$@" ) > "$OUT/$name.log" 2>&1; echo "exit $? ($name)"; tail -n "${TAILN:-4}" "$OUT/$name.log"; }
TAILN=6 T=1500 run pytest_gpu python -m pytest tests -m gpu -q
TAILN=2 T=900 run bench_ref python bench.py --impl reference --steps 20 --warmup 5
TAILN=2 T=900 run bench_n1 python bench.py --steps 20 --warmup 5
T=900 run ncu_launches ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file "$OUT/launches_r2.csv" python bench.py --steps 1 --warmup 3 --no-cpu-baseline --ncu-range --no-graph --skip-configs
T=400 run ncu_att ncu --set full --clock-control none --import-source on --profile-from-start off -o "$OUT/att_r2" -f python tools/att_ncu.py
T=400 run ncu_new ncu --set full --clock-control none --import-source on --profile-from-start off -o "$OUT/newk_r2" -f python tools/newkernels_ncu.py
T=300 run ablate python tools/ablate_step.py
python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.log" 2>&1; echo "smoke exit $?"; tail -2 "$OUT/smoke.log"
