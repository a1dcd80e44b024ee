cp fancy_gym_b200/lib/libfancygym_b200.so /tmp/lib_orig.so
for n in np2 np3; do cp build/lib_$n.so fancy_gym_b200/lib/libfancygym_b200.so; echo VARIANT=$n; python tools/probe_trajgen.py 2>&1 | head -1; done
cp /tmp/lib_orig.so fancy_gym_b200/lib/libfancygym_b200.so; echo VARIANT=default; python tools/probe_trajgen.py 2>&1 | head -1
