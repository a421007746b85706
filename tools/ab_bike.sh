#!/bin/bash
# A/B on the GPU box for the citi_bike line: tools/ab_bike.sh "<extra bench args>" variant1 variant2 ...  ("base" = product library)
args="$1"; shift
for v in "$@"; do
  lib=""; [ "$v" != "base" ] && lib="$PWD/maro_b200/libmaro_b200_$v.so"
  MARO_B200_LIB=$lib timeout 300 python bench.py --scenario citi_bike --steps 1280 --warmup 20 --skip-e2e --cpu-seconds 0.2 $args 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$v', '%.1f M/s' % (d['value']/1e6), 'kernel_us/step %.3f' % r['kernel_us'], 'frac %.4f' % r['frac'])"
done
