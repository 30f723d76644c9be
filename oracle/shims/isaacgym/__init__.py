"""TEST INFRASTRUCTURE ONLY -- stand-in for the proprietary `isaacgym` package so that the
unmodified reference files under /root/reference import on a CPU-only box.
Only `isaacgym.torch_utils` carries arithmetic (restated from the public IsaacGymEnvs
`torch_jit_utils.py`, which ships the same code)."""
