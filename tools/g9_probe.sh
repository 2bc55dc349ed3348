for d in 2.4 5 8 11; do echo "== G9 L2_DEBIAS=$d"; CIA_L2_TAPS_PER_FLUSH=9 CIA_L2_DEBIAS=$d python tools/heldout_probe.py 1 2>&1 | grep "^set" | grep -v "(0, 0, 0)"; done
echo "== G3 default"; python tools/heldout_probe.py 1 2>&1 | grep "^set" | grep -v "(0, 0, 0)"
CIA_L2_TAPS_PER_FLUSH=9 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('G9 bench', d['value'], {k:round(v['ms_per_step'],2) for k,v in d['cae_layers'].items()})"
