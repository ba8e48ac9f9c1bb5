#!/bin/bash
# scripts/ab_feature.sh <tag> ...: feature-path kernel times of each library variant (run under gpurun)
for tag in "$@"; do
  lib=salsa_b200/_build/libsalsa_$tag.so
  [ "$tag" = default ] && lib=salsa_b200/libsalsa_b200.so
  SALSA_B200_LIB=$PWD/$lib python bench.py --no-crnn --no-cpu-baseline --no-e2e --no-other-configs --no-fast-mode --no-train --steps 5 --warmup 2 > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err
  python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads(open('gpurun_out/ab_%s.json' % tag).read())
    print(tag, round(d['ms_per_step'], 3), {k[3:]: v for k, v in d['roofline'].items() if k.startswith('ms_')})
except Exception as e:
    print(tag, 'FAILED', e)
PY
done
