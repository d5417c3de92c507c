#!/bin/bash
# A/B timing of library variants built into build/: tools/abtest.sh name1 name2 ...  ("base" = the in-tree library)
for v in "$@"; do
  if [ "$v" = base ]; then lib=""; else lib="build/libwlm_$v.so"; fi
  ms=$(WLM_LIBRARY_PATH=$lib python bench.py --no-e2e --no-cpu-baseline --steps 30 ${ABTEST_ARGS} 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.4f' % d['ms_per_step'])")
  echo "$v: $ms ms"
done
