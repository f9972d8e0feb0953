#!/bin/bash
# runs the resident bench for each library variant given on the command line (default lib first)
run() { python bench.py --steps 10 --warmup 3 --wave-mib -1 --no-cpu $2 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k=d['roofline']['kernel_ms_per_step']; e=d.get('e2e') or {}
print('$1', 'Mframes/s %.2f'%(d['value']/1e6), 'ms %.2f'%d['ms_per_step'], {a:round(b,2) for a,b in k.items()}, 'e2e Mf/s %.2f ms %.1f'%(e.get('value',0)/1e6, e.get('ms_per_step',0)))"; }
run default "${E2E:---no-e2e}"
for v in "$@"; do MELONIX_B200_LIB=variants/$v.so run $v --no-e2e; done
