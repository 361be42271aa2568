#!/bin/bash
# One GPU-box session: parity tests, smoke, bench (both arms), event breakdown, ncu launch list + full capture.
# Usage (under gpurun): bash tools/gpu_round.sh <tag> [stages]     stages default: test,smoke,bench,breakdown,ncu
TAG=${1:-r01}
STAGES=${2:-test,smoke,bench,breakdown,ncu}
OUT=gpurun_out/$TAG
mkdir -p $OUT
has() { [[ ",$STAGES," == *",$1,"* ]]; }
nvidia-smi --query-gpu=name,driver_version,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/gpu.csv 2>&1
if has test; then
  timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
  tail -5 $OUT/pytest_gpu.log
fi
if has smoke; then
  timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke.log
fi
if has bench; then
  timeout 900 python bench.py --steps 3 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cat $OUT/bench.json | cut -c1-1500
  timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "bench ref rc=$?"; cat $OUT/bench_ref.json | cut -c1-600
fi
if has breakdown; then
  for ii in 0 3 6 9; do timeout 300 python tools/profile_step.py --ii $ii --steps 3 --breakdown >> $OUT/breakdown.jsonl 2>> $OUT/breakdown.err; done
  cat $OUT/breakdown.jsonl
fi
if has ncu; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 600 --csv --log-file $OUT/launches_step.csv \
      python tools/profile_step.py --ii 3 --steps 1 --warm 1 > $OUT/ncu_step.log 2>&1; echo "ncu list rc=$?"
  # profile_step runs 1 warm-up + 1 profiled step: per step 36 gemm_persist (QKV, O, fc1 x 12 blocks), 12 gemm_wide
  # (fc2), 12 BERT + 12 CLIP attention launches
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_persist -s 51 -c 3 -f -o $OUT/prof_gemm \
      python tools/profile_step.py --ii 3 --steps 1 --warm 1 > $OUT/ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_wide -s 17 -c 1 -f -o $OUT/prof_gemm_wide \
      python tools/profile_step.py --ii 3 --steps 1 --warm 1 > $OUT/ncu_gemm_wide.log 2>&1; echo "ncu gemm_wide rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_mma -s 41 -c 2 -f -o $OUT/prof_attn \
      python tools/profile_step.py --ii 3 --steps 1 --warm 1 > $OUT/ncu_attn.log 2>&1; echo "ncu attn rc=$?"
  ls -la $OUT
fi
