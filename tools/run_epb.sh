for e in 8 16 32 64 128 256; do echo EPB=$e; FG_TRAJ_EPB=$e python tools/probe_trajgen.py 2>&1 | head -1; done
