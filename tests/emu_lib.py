"""Loader of tests/emu/libmrp_emu.so — the kernel source of gym_puzzles_b200/csrc compiled for the host
(-DMRP_HOST_EMU).  Debug/test tool for the GPU-less build container only; the package never loads it."""
import os
import subprocess

from gym_puzzles_b200 import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
EMU_DIR = os.path.join(_HERE, "emu")
_lib = None


def emu_lib():
    global _lib
    if _lib is None:
        subprocess.check_call(["make", "-C", EMU_DIR, "-s"])
        _lib = abi.MrpLib(os.path.join(EMU_DIR, "libmrp_emu.so"))
        assert _lib.backend.startswith("host-emu")
    return _lib
