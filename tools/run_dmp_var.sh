cp fancy_gym_b200/lib/libfancygym_b200.so /tmp/lib_default.so
for v in default 128_8 128_12 128_16 256_20 256_12 256_8; do
  if [ $v = default ]; then cp /tmp/lib_default.so fancy_gym_b200/lib/libfancygym_b200.so; else cp build/lib_dmp_$v.so fancy_gym_b200/lib/libfancygym_b200.so; fi
  echo -n "$v: "; python tools/probe_trajgen_dmp.py 2>&1 | tail -1
done
