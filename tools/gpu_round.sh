#!/bin/bash
# One GPU-box session: parity tests, smoke, bench (both arms), event breakdown, ncu launch list + full captures.
# Usage (under gpurun): bash tools/gpu_round.sh <tag> [stages]
#   stages (comma separated): test,smoke,bound,bench,benchx3,configs,breakdown,ncu,ncusmall   default: test,smoke,bench
TAG=${1:-r02}
STAGES=${2:-test,smoke,bench}
OUT=gpurun_out/$TAG
mkdir -p $OUT
has() { [[ ",$STAGES," == *",$1,"* ]]; }
nvidia-smi --query-gpu=name,driver_version,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/gpu.csv 2>&1
nproc > $OUT/nproc.txt
if has smoke; then
  timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $OUT/smoke.log
fi
if has test; then
  timeout 2400 python -m pytest tests -m gpu -q --maxfail=40 --durations=15 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
  tail -30 $OUT/pytest_gpu.log
fi
if has testq; then  # the kernels a GEMM / tower change touches
  timeout 900 python -m pytest tests -m gpu -q --maxfail=10 -k "linear or clip_text or gemm_epilogue or teacher_forced_certified or free_running_certified" > $OUT/pytest_gpu_quick.log 2>&1; echo "pytest quick rc=$?"
  tail -4 $OUT/pytest_gpu_quick.log
fi
if has bound; then
  timeout 600 python tools/cert_bound.py > $OUT/cert_bound.json 2> $OUT/cert_bound.err; echo "bound rc=$?"
  python -c "import json;d=json.load(open('$OUT/cert_bound.json'));d.pop('per_step');print(json.dumps(d,indent=1))"
  timeout 600 python tools/cert_bound.py --order shuffle --sweeps 3 > $OUT/cert_bound_shuffle.json 2>> $OUT/cert_bound.err
  python -c "import json;d=json.load(open('$OUT/cert_bound_shuffle.json'));d.pop('per_step');print(json.dumps(d)[:900])"
fi
if has bench; then
  timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cat $OUT/bench.json | cut -c1-3000
  timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "bench ref rc=$?"; cat $OUT/bench_ref.json | cut -c1-600
fi
if has benchx3; then
  timeout 900 python bench.py --steps 3 --warmup 3 --precision bf16x3 --no-cpu > $OUT/bench_bf16x3.json 2> $OUT/bench_bf16x3.err; echo "bench x3 rc=$?"; cat $OUT/bench_bf16x3.json | cut -c1-1200
  timeout 900 python bench.py --steps 3 --warmup 3 --precision bf16 --no-cpu > $OUT/bench_bf16.json 2> $OUT/bench_bf16.err; echo "bench bf16 rc=$?"; cat $OUT/bench_bf16.json | cut -c1-1200
fi
if has configs; then
  for c in 3 4; do
    timeout 900 python bench.py --config $c --steps 3 --warmup 3 --no-cpu > $OUT/bench_config$c.json 2> $OUT/bench_config$c.err; echo "config $c rc=$?"; cat $OUT/bench_config$c.json | cut -c1-1500
  done
  timeout 1800 python bench.py --config 5 --steps 2 --warmup 3 --no-cpu > $OUT/bench_config5.json 2> $OUT/bench_config5.err; echo "config 5 rc=$?"; cat $OUT/bench_config5.json | cut -c1-600
fi
if has ab; then  # A/B of the switches named in $AB_ENVS (space separated VAR=VALUE), interleaved with the default
  for rep in 1 2; do
    for ev in "X=0" $AB_ENVS; do
      env $ev timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-parity > $OUT/ab_${ev//=/_}_$rep.json 2>> $OUT/ab.err
      python -c "import json,sys;d=json.load(open('$OUT/ab_${ev//=/_}_$rep.json'));print('$ev',d['value'],d['e2e']['value'],d.get('phase_breakdown_ms'))"
    done
  done
fi
if has multi; then  # gpurun --gpus N: the BASELINE configurations that name a GPU count ($MULTI_CONFIGS, e.g. "3 5")
  N=${NGPU:-8}
  nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
  for c in $MULTI_CONFIGS; do
    timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$c \
        bench.py --gpus $N --config $c --steps ${MULTI_STEPS:-3} --warmup 3 --no-cpu > $OUT/bench_config${c}_n$N.json 2> $OUT/bench_config${c}_n$N.err
    echo "config $c x $N GPUs rc=$?"; cut -c1-1200 $OUT/bench_config${c}_n$N.json; tail -3 $OUT/bench_config${c}_n$N.err
  done
fi
if has sanitize; then  # memcheck over the smoke step and the GEMM epilogue forms (TMA boxes, clipped last tiles)
  timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/sanitizer_smoke.log 2>&1; echo "memcheck smoke rc=$?"; tail -4 $OUT/sanitizer_smoke.log
  timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -q -x -k "gemm_epilogue or nontrivial_layernorm" > $OUT/sanitizer_epilogue.log 2>&1; echo "memcheck epilogue tests rc=$?"; tail -4 $OUT/sanitizer_epilogue.log
fi
if has config4; then
  timeout 600 python bench.py --config 4 --steps 2 --warmup 3 --no-cpu > $OUT/bench_config4.json 2> $OUT/bench_config4.err; echo "config 4 rc=$?"
  python -c "import json;d=json.loads(open('$OUT/bench_config4.json').read().strip().splitlines()[-1]);print(d['value'],d['parity'])"
fi
if has vocab; then
  timeout 900 python tools/bench_vocab_paths.py > $OUT/vocab_paths.jsonl 2> $OUT/vocab_paths.err; echo "vocab rc=$?"; cat $OUT/vocab_paths.jsonl
fi
if has breakdown; then
  for ii in 0 3 6 9; do timeout 300 python tools/profile_step.py --ii $ii --steps 3 --breakdown >> $OUT/breakdown.jsonl 2>> $OUT/breakdown.err; done
  cat $OUT/breakdown.jsonl
fi
if has ncustep; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $OUT/launches_step.csv \
      python tools/profile_step.py --ii 3 --steps 1 --warm 1 > $OUT/ncu_step.log 2>&1; echo "ncu list rc=$?"
fi
if has ncu; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $OUT/launches_step.csv \
      python tools/profile_step.py --ii 3 --steps 1 --warm 1 > $OUT/ncu_step.log 2>&1; echo "ncu list rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_persist -s 30 -c 4 -f -o $OUT/prof_gemm \
      python tools/profile_step.py --ii 3 --steps 1 --warm 1 > $OUT/ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_wide -s 26 -c 2 -f -o $OUT/prof_gemm_wide \
      python tools/profile_step.py --ii 3 --steps 1 --warm 1 > $OUT/ncu_gemm_wide.log 2>&1; echo "ncu gemm_wide rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_mma -s 18 -c 1 -f -o $OUT/prof_attn \
      python tools/profile_step.py --ii 3 --steps 1 --warm 1 > $OUT/ncu_attn.log 2>&1; echo "ncu attn rc=$?"
fi
if has ncuwide; then
  env $NCU_ENV timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_wide -s 26 -c 2 -f -o $OUT/prof_gemm_wide \
      python tools/profile_step.py --ii 3 --steps 1 --warm 1 > $OUT/ncu_gemm_wide.log 2>&1; echo "ncu gemm_wide rc=$?"
fi
if has ncusmall; then
  # the once-per-pass kernels of the measured step (the warm-up step launches ~12 of them first), then a few launches of the
  # fp32 attention / LayerNorm kernels from the middle of the step
  timeout 900 ncu --set full --clock-control none --import-source on \
      -k regex:"topk_cluster|clip_logits|cert_round|cert_gather|assemble_kernel|clip_embed|bert_embed|pool_index|image_resize" -s 12 -c 16 -f -o $OUT/prof_small \
      python tools/profile_step.py --ii 3 --steps 1 --warm 1 > $OUT/ncu_small.log 2>&1; echo "ncu small rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"attention_kernel|layernorm_kernel" -s 70 -c 6 -f -o $OUT/prof_small2 \
      python tools/profile_step.py --ii 3 --steps 1 --warm 1 > $OUT/ncu_small2.log 2>&1; echo "ncu small2 rc=$?"
fi
ls -la $OUT
