"""A compiled C host of the C ABI (examples/loco_encode_file.c: C99 + the CUDA runtime, no Python / torch in the process).
CPU: it compiles against include/loco_asr.h as plain C, links against the in-tree library and fails loudly without a GPU.
GPU: fed the same weights and waveforms through files, it returns the bits the ctypes binding returns."""
import os
import shutil
import struct
import subprocess

import numpy as np
import pytest
import torch

from loco_asr_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA_HOME = os.environ.get("CUDA_HOME", "/usr/local/cuda")


def build_c_host(out_dir) -> str:
    if shutil.which("gcc") is None or not os.path.exists(os.path.join(CUDA_HOME, "include", "cuda_runtime_api.h")):
        pytest.skip("gcc or the CUDA runtime headers are not available")
    assert os.path.exists(_lib.LIB_PATH), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    exe = os.path.join(str(out_dir), "loco_encode_file")
    libdir, cudalib = os.path.dirname(_lib.LIB_PATH), os.path.join(CUDA_HOME, "lib64")
    cmd = ["gcc", "-std=c99", "-O2", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I" + os.path.join(ROOT, "include"),
           "-isystem", os.path.join(CUDA_HOME, "include"), os.path.join(ROOT, "examples", "loco_encode_file.c"), "-o", exe,
           "-L" + libdir, "-l:libloco_asr.so", "-L" + cudalib, "-lcudart", "-Wl,-rpath," + libdir, "-Wl,-rpath," + cudalib]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def write_weights(path, sd):
    with open(path, "wb") as fh:
        for k, v in sd.items():
            a = np.ascontiguousarray(v.detach().to(torch.float32).cpu().numpy())
            kb = k.encode()
            fh.write(struct.pack("<I", len(kb)) + kb + struct.pack("<I", a.ndim) + struct.pack("<%dq" % a.ndim, *a.shape))
            fh.write(a.tobytes())


def write_waves(path, waves):
    with open(path, "wb") as fh:
        fh.write(struct.pack("<i", len(waves)))
        fh.write(np.asarray([len(w) for w in waves], dtype=np.int32).tobytes())
        for w in waves:
            fh.write(np.ascontiguousarray(w, dtype=np.float32).tobytes())


def test_c_host_compiles_links_and_fails_loudly_without_a_gpu(tmp_path):
    exe = build_c_host(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 1 and "usage" in r.stderr
    if torch.cuda.is_available():
        return
    for name in ("w.bin", "x.bin"):
        open(tmp_path / name, "wb").close()
    r = subprocess.run([exe, str(tmp_path / "w.bin"), str(tmp_path / "x.bin"), str(tmp_path / "o.bin")], capture_output=True, text=True)
    assert r.returncode == 2, (r.returncode, r.stderr)
    assert "no usable CUDA device" in r.stderr and "no CPU fallback" in r.stderr
    assert not os.path.exists(tmp_path / "o.bin")


@pytest.mark.gpu
def test_c_host_returns_the_bits_of_the_python_binding(tmp_path):
    from loco_asr_b200.encoder import LocoSpeechT5Encoder
    from loco_asr_b200.synth import synth_state_dict, synth_wave

    exe = build_c_host(tmp_path)
    sd = synth_state_dict(seed=0)
    lengths = [400, 6400, 16000, 23456, 48000, 70001]
    waves = [synth_wave(n, 7, i) for i, n in enumerate(lengths)]
    write_weights(tmp_path / "w.bin", sd)
    write_waves(tmp_path / "x.bin", waves)
    r = subprocess.run([exe, str(tmp_path / "w.bin"), str(tmp_path / "x.bin"), str(tmp_path / "o.bin")], capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0, r.stderr
    assert "both paths bit-identical" in r.stderr
    got = torch.from_numpy(np.fromfile(tmp_path / "o.bin", dtype=np.float32).reshape(len(waves), 768))
    enc = LocoSpeechT5Encoder.from_state_dict(sd, device="cuda:0")
    want = enc.encode_packed(torch.from_numpy(np.concatenate(waves)).cuda(), lengths).cpu()
    assert torch.isfinite(got).all()
    assert torch.equal(got, want)
    # a bad file is an error code and a message, not a crash
    with open(tmp_path / "bad.bin", "wb") as fh:
        fh.write(struct.pack("<I", 9) + b"decoder.x" + struct.pack("<I", 1) + struct.pack("<q", 2) + np.zeros(2, np.float32).tobytes())
    r = subprocess.run([exe, str(tmp_path / "bad.bin"), str(tmp_path / "x.bin"), str(tmp_path / "o2.bin")], capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 2 and "unknown tensor key" in r.stderr
