for m in 2 3 4; do cp build/lib_minb$m.so fancy_gym_b200/lib/libfancygym_b200.so; for e in 32 64; do echo MINB=$m EPB=$e; FG_TRAJ_EPB=$e python tools/probe_trajgen.py 2>&1 | head -1; done; done
cp build/lib_minb2.so fancy_gym_b200/lib/libfancygym_b200.so; python -m pytest tests -m gpu -q -k "trajgen" 2>&1 | tail -2
