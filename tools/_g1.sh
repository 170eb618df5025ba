python -m pytest tests -x -q -m gpu 2>&1 | tail -2
for r in 3 4 5 6 7; do
  PLVI_LSD_BR_ROWS=$r python bench.py --no-cpu --batch 1 --steps 50 --warmup 5 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); km=d['kernel_ms']; print('rows$r b1', d['ms_per_step'], km.get('k_lsd_band_rounds'))
"
done
for r in 4 6; do
  PLVI_LSD_BR_ROWS=$r python bench.py --no-cpu --batch 8 --pipes 1 --steps 20 --warmup 3 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); km=d['kernel_ms']; print('rows$r b8', d['ms_per_step'], km.get('k_lsd_band_rounds'))
"
done
