#!/usr/bin/env python3
"""File exchange with java/src/io/github/dsheirer/gpu/OracleHarness.java (see tools/mint_jvm_goldens.sh).

    jvm_goldens.py export DIR/in     inputs of the committed goldens (tests/golden/*.npz) as little-endian float32 files
    jvm_goldens.py import DIR        DIR/out/* (written by the REAL sdrtrunk classes) -> DIR/jvm.npz

The .npz keeps the names of tests/golden/: tests/test_jvm_goldens.py compares the oracle and the CUDA path with it."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
G = os.path.join(ROOT, "tests", "golden")


def nbfm_input():
    """configs[0]: an NBFM carrier at the 50 kHz polyphase channel rate, 6 assembler buffers (first one below the squelch)"""
    import siggen as sg
    rng = np.random.default_rng(20240103)
    n = 6 * 1024
    z = sg.nbfm(50000.0, n, audio_hz=1000.0, carrier_offset=150.0) * 0.5 + sg.awgn(rng, n, 1e-3)
    z[:1024] *= 1e-4
    return sg.interleave(z)


def export(directory):
    os.makedirs(directory, exist_ok=True)

    def put(name, a):
        np.asarray(a, "<f4").tofile(os.path.join(directory, name))

    put("channelizer_m96_x.f32", np.load(os.path.join(G, "channelizer_m96.npz"))["x"])
    put("filters_x.f32", np.load(os.path.join(G, "filters.npz"))["x"])
    put("fm_x.f32", np.load(os.path.join(G, "fm.npz"))["x"])
    put("nbfm_x.f32", nbfm_input())
    p = np.load(os.path.join(G, "p25_chains.npz"))
    for kind in ("c4fm", "lsm", "hdqpsk", "dmr"):
        put(kind + "_x.f32", p[kind + "_x"])
    print("exported", sorted(os.listdir(directory)))


def import_(directory):
    out = os.path.join(directory, "out")
    data = {}
    for f in sorted(os.listdir(out)):
        name, ext = os.path.splitext(f)
        data[name] = np.fromfile(os.path.join(out, f), "<f4" if ext == ".f32" else np.uint8)
    for f in sorted(os.listdir(os.path.join(directory, "in"))):
        data[os.path.splitext(f)[0]] = np.fromfile(os.path.join(directory, "in", f), "<f4")
    np.savez_compressed(os.path.join(directory, "jvm.npz"), **data)
    print("imported", sorted(data))


if __name__ == "__main__":
    if len(sys.argv) == 3 and sys.argv[1] == "export":
        export(sys.argv[2])
    elif len(sys.argv) == 3 and sys.argv[1] == "import":
        import_(sys.argv[2])
    else:
        sys.exit(__doc__)
