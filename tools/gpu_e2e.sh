#!/bin/bash
mkdir -p gpurun_out
{
python tools/e2e_probe.py 2>&1 | head -7
timeout 600 python -m pytest tests -m gpu -q -x -k "pack or spadd or spgemm or sddmm or dropin" 2>&1 | tail -3
for wl in pack spadd spgemm; do
  timeout 600 python bench.py --workload $wl --no-cpu --steps 10 2>/dev/null | tail -1 | python -c "
import sys,json
j=json.loads(sys.stdin.read()); e=j['e2e']
print('$wl', 'value', round(j['value'],2), 'ms', round(j['ms_per_step'],3), 'e2e', round(e['value'],3), 'e2e_ms', round(e['ms_per_step'],2), 'h2d', e['h2d_bytes_per_step'], 'd2h', e['d2h_bytes_per_step'])"
done
} > gpurun_out/e2e_1.txt 2>&1
cat gpurun_out/e2e_1.txt
