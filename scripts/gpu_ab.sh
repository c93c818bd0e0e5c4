#!/bin/bash
# A/B of one environment switch on the headline bench, interleaved (A B A B) on ONE box so that box-to-box clock / power
# differences cancel:   gpurun -- 'bash scripts/gpu_ab.sh MMAE_COLRED_FOLD 0 1'
var=$1; a=$2; b=$3; shift 3
mkdir -p gpurun_out
for rep in 1 2; do
  for v in $a $b; do
    env $var=$v timeout 120 python bench.py --steps 30 --warmup 5 --cpu-baseline 0 "$@" > gpurun_out/ab_${var}_${v}_$rep.json 2> gpurun_out/ab_${var}_${v}_$rep.err
    echo "$var=$v rep $rep rc=$? $(python -c "import json,sys; d=json.load(open('gpurun_out/ab_${var}_${v}_$rep.json')); print(d['ms_per_step'], d['value'], d.get('gpu_launches'))" 2>/dev/null)"
  done
done
