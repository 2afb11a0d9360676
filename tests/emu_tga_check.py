"""Developer loop: the emulated packetiser (tests/emu/emu_tga.cpp) against the sequential algorithm on crafted streams."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from rle_model import sequential_packets  # noqa: E402
from test_tga_rle import crafted_stream  # noqa: E402

subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "emu")])
emu = C.CDLL(os.path.join(HERE, "emu", "_build", "libhana_emu_tga.so"))
emu.emu_tga_payload.restype = C.c_size_t


def run(px, threads):
    px = np.ascontiguousarray(px, np.uint32)
    out = np.zeros(4 * len(px) + 64, np.uint8)
    L = emu.emu_tga_payload(px.ctypes.data_as(C.c_void_p), len(px), threads, out.ctypes.data_as(C.c_void_p))
    assert L < 1 << 60, ("disagree at word", (1 << 64) - 1 - L, len(px))
    return out[:L].tobytes()


if __name__ == "__main__":
    rng = np.random.RandomState(int(sys.argv[1]) if len(sys.argv) > 1 else 13)
    bad = 0
    for trial in range(int(sys.argv[2]) if len(sys.argv) > 2 else 1500):
        n = int(rng.choice([1, 2, 3, 5, 31, 32, 33, 63, 64, 65, 127, 128, 129, 255, 256, 257, 1000, 1023, 1024, 1025, 4096, rng.randint(1, 6000)]))
        px = crafted_stream(rng, n, trial % 5)
        ref = sequential_packets(px)
        for th in (1, 5, 16, 1024):
            if run(px, th) != ref:
                bad += 1
                print("BAD", trial, n, th)
                break
    print("bad", bad)
