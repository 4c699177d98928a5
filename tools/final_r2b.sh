#!/usr/bin/env bash
# Evidence refresh at the end of round 2 (one GPU): A/B logs of the round's kernels, ncu of the new kernels, both bench arms.
set -u
OUT=gpurun_out/final_r2b
mkdir -p "$OUT"
timeout 300 python tools/conv_ab.py > "$OUT/conv_halo_ab.txt" 2>&1; tail -8 "$OUT/conv_halo_ab.txt"
timeout 200 python tools/outk48_prof.py > "$OUT/outk48.txt" 2>&1; tail -4 "$OUT/outk48.txt"
timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -o "$OUT/newk" -f python tools/newkernels_ncu.py > "$OUT/ncu_new.log" 2>&1; tail -2 "$OUT/ncu_new.log"
timeout 400 python tools/encoder_sweep.py 8 > "$OUT/encoder_sweep_b8.txt" 2>&1; timeout 300 python tools/encoder_sweep.py 1 > "$OUT/encoder_sweep_b1.txt" 2>&1; tail -3 "$OUT/encoder_sweep_b1.txt"
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > "$OUT/bench_ref.log" 2>&1; tail -c 400 "$OUT/bench_ref.log"
timeout 900 python bench.py --steps 20 --warmup 5 > "$OUT/bench_n1.log" 2>&1; tail -c 300 "$OUT/bench_n1.log"
