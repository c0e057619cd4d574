for mode in roles fused; do
for B in 8192 16384 32768 65536 131072 262144; do
LTK_SWEEP=$mode python bench.py --steps 5 --warmup 3 --no-cpu-baseline --candidates $B 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$mode', $B, round(d['ms_per_step'],4), d['roofline']['kernel_ms'])"
done; done
