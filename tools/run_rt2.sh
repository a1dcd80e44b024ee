# block-size experiment for k_rollout<HOLE, PROMP, vel, 5>: 128 (default) vs 64 vs 32 threads per block
cp fancy_gym_b200/lib/libfancygym_b200.so /tmp/lib_default.so
for t in default rt64 rt32; do
  if [ $t = default ]; then cp /tmp/lib_default.so fancy_gym_b200/lib/libfancygym_b200.so; else cp build/lib_$t.so fancy_gym_b200/lib/libfancygym_b200.so; fi
  for B in 65536 1048576; do
    python bench.py --no-cpu-baseline --steps 60 --envs-per-gpu $B 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$t', $B, 'value %.4e' % d['value'], 'ms/step %.4f' % d['ms_per_step'], 'kernel_ms %.4f' % d['roofline']['kernel_ms'], 'e2e %.4e' % d['e2e']['value'])"
  done
done
