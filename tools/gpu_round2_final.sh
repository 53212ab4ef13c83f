#!/bin/bash
# Final GPU pass of round 2: parity suite, bench lines (default K, the driver's K, the other env families, the reference
# arm), the ncu launch list of the bench command, one ncu --set full capture per env family (stationary regime), the
# small configs and the gym path. Outputs under gpurun_out/ (copied into profiles/ afterwards).
T=${1:-r2g}
python -m pytest tests -m gpu -q -s 2>&1 | tail -120 > gpurun_out/${T}_pytest_gpu.log
grep -E "passed|failed|FAILED|parity\]" gpurun_out/${T}_pytest_gpu.log | tail -24
python bench.py > gpurun_out/${T}_bench_hh_1gpu.json 2> gpurun_out/${T}_bench_hh_1gpu.err
python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench_hh_1gpu_k20.json 2> gpurun_out/${T}_bench_hh_1gpu_k20.err
for e in ant ant_tag ant_gather; do
  python bench.py --env $e --no-e2e --no-cpu-baseline --no-per-config > gpurun_out/${T}_bench_${e}_1gpu.json 2> gpurun_out/${T}_bench_${e}_1gpu.err
done
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${T}_bench_reference_arm.json 2> gpurun_out/${T}_bench_reference_arm.err
python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-per-config > gpurun_out/${T}_plain_for_launches.json 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches.csv \
    python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-per-config > gpurun_out/${T}_ncu_launches.log 2>&1
for f in gpurun_out/${T}_bench_*_1gpu*.json; do python - "$f" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1].split('/')[-1], 'ms/step %.4f' % d['ms_per_step'], 'value %.3e' % d['value'], 'frac %.3f' % d['roofline']['frac'],
      'e2e', d['e2e'] and '%.3e' % d['e2e']['value'], 'early %.4f' % d['early_phase']['ms_per_step'])
PY
done
tail -c 400 gpurun_out/${T}_bench_reference_arm.json; wc -l gpurun_out/${T}_launches.csv
python tools/bench_small.py > gpurun_out/${T}_small.log 2>&1; cat gpurun_out/${T}_small.log
(python tools/bench_gym.py; python tools/bench_gym.py 16) > gpurun_out/${T}_gym.log 2>&1; cat gpurun_out/${T}_gym.log
for e in ant_heavenhell ant ant_tag ant_gather; do bash tools/gpu_profile.sh $e $T; done
