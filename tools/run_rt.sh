for t in 64 96; do cp build/lib_rt$t.so fancy_gym_b200/lib/libfancygym_b200.so; echo THREADS=$t; python tools/bench_configs.py 2>&1 | grep -E "config2|config5"; done
