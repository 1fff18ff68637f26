"""The radix passes, padded indexing and digit-reversed read-out of the fused
unpack+FFT kernel (pencilarrays.jl_b200/csrc/fft_core.hpp, shared host/device code)
run on the CPU through tests/fft_host_check.cpp, against numpy.fft."""
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def test_fft_core_matches_numpy(tmp_path):
    exe = os.path.join(str(tmp_path), "fft_host_check")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(HERE, "fft_host_check.cpp")],
                   check=True)
    rng = np.random.default_rng(0)
    for L in (2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048):
        for sign in (-1, 1):
            x = (rng.standard_normal(L) + 1j * rng.standard_normal(L)).astype(np.complex128)
            out = subprocess.run([exe, str(L), str(sign)], input=x.tobytes(), capture_output=True,
                                 check=True).stdout
            y = np.frombuffer(out, dtype=np.complex128)
            ref = np.fft.fft(x) if sign < 0 else np.fft.ifft(x) * L
            assert np.abs(y - ref).max() <= 8 * np.finfo(np.float64).eps * max(1, np.log2(L)) * np.abs(ref).max()
