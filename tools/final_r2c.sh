#!/bin/bash
# Evidence of the last session of round 2 (one GPU, ~6 min): the GEMM producer-thread limit and the branch-free epilogue.
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/final_r2c.sh'
# Writes gpurun_out/r2c/*; the summaries under profiles/*r2c* / producer_*_r2b.txt were copied from there.
OUT=gpurun_out/r2c
mkdir -p "$OUT"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/tma_issue_bench tools/ubench/tma_issue_bench.cu -lcuda -L/usr/local/cuda/lib64/stubs 2> "$OUT/nvcc.log"
timeout 100 tools/ubench/tma_issue_bench > "$OUT/tma_issue.txt" 2>&1
timeout 120 python tools/pair_n_sweep.py > "$OUT/pair_n_sweep.txt" 2>&1
L4P_GEMM_K128=0 timeout 120 python tools/pair_n_sweep.py > "$OUT/pair_n_sweep_k64.txt" 2>&1
timeout 400 python tools/blockn_sweep2.py > "$OUT/blockn_sweep2.txt" 2>&1
timeout 100 python tools/gemm_prof.py > "$OUT/gemm_prof.txt" 2>&1
timeout 300 python tools/encoder_sweep.py 1 > "$OUT/encoder_sweep_b1.txt" 2>&1
timeout 300 python tools/encoder_sweep.py 8 > "$OUT/encoder_sweep_b8.txt" 2>&1
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -o "$OUT/encgemm" -f python tools/encgemm_ncu.py > "$OUT/ncu.log" 2>&1
for v in k0 new k0 new; do
  if [ $v = k0 ]; then E="L4P_GEMM_K128=0"; else E="L4P_X=1"; fi
  env $E python bench.py --steps 20 --warmup 5 --no-cpu-baseline --skip-configs 2>/dev/null | grep "^{" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v', round(d['ms_per_step'],3), round(d['value'],1), d['clocks']['sm_mhz'])"
done > "$OUT/ab_k128.txt" 2>&1
python bench.py --steps 20 --warmup 5 > "$OUT/bench_n1.log" 2>&1
tail -c 300 "$OUT/bench_n1.log"
