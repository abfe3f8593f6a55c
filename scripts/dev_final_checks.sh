#!/bin/bash
# GPU-box: the driver's checks in one go (single-process GPU suite, smoke, default bench) + a CG A/B at the 8x8 level
mkdir -p gpurun_out
( time python -m pytest tests -x -q -m gpu ) > gpurun_out/r02_pytest_gpu_full.log 2>&1
tail -4 gpurun_out/r02_pytest_gpu_full.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== dev_gemm_shapes level 3 (8x8), default CG vs AP_GEMM_CG=2"
python scripts/dev_gemm_shapes.py 2>&1 | sed -n '/level 3/,/level 4/p' | head -8
AP_GEMM_CG=2 python scripts/dev_gemm_shapes.py 2>&1 | sed -n '/level 3/,/level 4/p' | head -8
( time python bench.py ) > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err
tail -3 gpurun_out/r02_bench_default.err
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_default.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e'], d['phases_ms'], d['parity_at_bench_shape'], d['c1'], d['cpu_baseline']['value'], d['cpu_baseline']['seconds'], d['cpu_baseline']['c1'], d['roofline_unet_call'], d['roofline_ref_attention']['launch_ms'])"
