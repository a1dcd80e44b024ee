"""CPU oracle for the movement-primitive black-box rollout path of ALRhub/fancy_gym.

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / `--impl reference` leg may import it, and only as the checker or
as the timed CPU baseline.  fancy_gym_b200/ never imports it and has no CPU fallback.

Contents
  reacher.py     numpy restatement of the classic_control reacher envs (fancy_gym/envs/
                 classic_control/**), batched over environments, same mixed fp32/fp64 arithmetic
                 as the reference.  PINNED: checked against the reference's own files (run
                 unmodified through ref_loader.py) by tests/golden/make_golden.py and against the
                 committed vectors in tests/golden/ by tests/test_oracle_env.py.
  mp.py          restatement of mp_pytorch<=0.1.3 (ProMP / DMP / ProDMP, phase and basis
                 generators).  PARITY UNPINNED: mp_pytorch is a third-party dependency of the
                 reference (pyproject.toml:30) that is neither vendored in /root/reference nor
                 installed here, and the reference holds no numeric golden vector for it
                 (SURVEY.md §8c).  The restatement follows the library's published algorithm
                  and is anchored on the reference's structural tests
                 (param counts, trajectory length, tau/delay plateaus).
  blackbox.py    restatement of BlackBoxWrapper.step / get_trajectory
                 (fancy_gym/black_box/black_box_wrapper.py:96-217) on top of the two above.
  ref_loader.py  runs the reference's files from /root/reference under refstub/ (this
                 container only).
"""
