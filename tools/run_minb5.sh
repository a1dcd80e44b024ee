cp fancy_gym_b200/lib/libfancygym_b200.so /tmp/lib_orig.so
cp build/lib_minb5.so fancy_gym_b200/lib/libfancygym_b200.so; echo VARIANT=minb5; python tools/bench_configs.py 2>&1 | grep -E "config2|config5"
cp /tmp/lib_orig.so fancy_gym_b200/lib/libfancygym_b200.so
