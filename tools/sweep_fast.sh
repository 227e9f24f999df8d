#!/bin/bash
# usage: tools/sweep_fast.sh "<extra nvcc flags for fast_kernels.cu>"   (GPU box: rebuild, msd bench + full-size drift)
flags="$1"
touch cgmres_cpp_b200/csrc/fast_kernels.cu
make -C cgmres_cpp_b200/csrc FAST_EXTRA="$flags" > /dev/null 2>&1 || { echo "build failed: $flags"; exit 1; }
python bench.py --steps 40 --warmup 4 --no-cpu-baseline --no-other-modes --mode fast 2>&1 | tail -1 > /tmp/sweep_line.json
python tools/drift_full.py --model msd > /tmp/drift.json
python - "$flags" <<'PY'
import json, sys
d = json.loads(open('/tmp/sweep_line.json').read()); r = json.loads(open('/tmp/drift.json').read())
print('%-40s ms/step %.3f value %.3e | drift max %.3e p99 %.3e n>1e-6 %d' % (sys.argv[1] or '(default)', d['ms_per_step'], d['value'], r['max_abs_dx'], r['p99_abs_dx'], r['n_above_1e-6']))
PY
