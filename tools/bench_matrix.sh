#!/bin/bash
# quick matrix of bench configurations (device-resident only), one line each
for extra in "--content noise" "--content smooth" "--content noise --flags 2" "--content smooth --flags 2"; do
  python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline $extra 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['config']['content'], 'flags', d['config']['flags'], 'kernel', d['config']['kernel_id'], round(d['value']), 'Mpix/s frac', round(d['roofline']['frac'],4), d['clocks']['sm_mhz'], d['clocks']['reasons'])"
done
