#!/bin/bash
# A/B on the GPU box: one short CIM bench line (value, kernel us per step) per library variant.
#   tools/ab.sh "<extra bench args>" variant1 variant2 ...     (variant "base" = the product library)
args="$1"; shift
for v in "$@"; do
  lib=""; [ "$v" != "base" ] && lib="$PWD/maro_b200/libmaro_b200_$v.so"
  MARO_B200_LIB=$lib timeout 300 python bench.py --steps 640 --warmup 20 --skip-extras --skip-e2e --cpu-seconds 0.2 $args 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$v', '%.1f M/s' % (d['value']/1e6), 'kernel_us/step %.3f' % r['kernel_us_per_step'], 'frac %.4f' % r['frac'])"
done
