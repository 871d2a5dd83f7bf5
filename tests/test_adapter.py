"""The MFEM-side adapter (adapter/b200_mfem.hpp, compiled against the reference's own headers
and linked with the unmodified reference library): on a GPU it must reproduce MFEM's own CPU
PA path at all three seams; without a GPU it must fail loudly (there is no CPU fallback)."""
import os
import subprocess

import pytest

from conftest import ROOT

EXE = os.path.join(ROOT, "adapter", "_build", "ex1_b200")
needs_exe = pytest.mark.skipif(not os.path.isfile(EXE), reason="adapter/_build/ex1_b200 not built "
                               "(needs /root/reference: make -C adapter)")


@needs_exe
def test_adapter_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([EXE, "-n", "2", "-o", "2"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=120)
    out = r.stdout.decode(errors="replace")
    assert r.returncode != 0
    assert "no CPU path" in out


@needs_exe
@pytest.mark.gpu
@pytest.mark.parametrize("args", [["-n", "6", "-o", "3"], ["-n", "4", "-o", "4"], ["-n", "3", "-o", "8"],
                                  ["-n", "5", "-o", "2", "-k", "mass"], ["-n", "8", "-o", "1"],
                                  ["-n", "4", "-o", "3", "-vd", "3"], ["-n", "3", "-o", "4", "-vd", "3", "-k", "mass"]])
def test_adapter_matches_mfem_at_all_seams(args):
    r = subprocess.run([EXE] + args, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600)
    out = r.stdout.decode(errors="replace")
    assert r.returncode == 0 and "all checks passed" in out, out[-3000:]


EXE_CUDA = os.path.join(ROOT, "adapter", "_build", "ex1_b200_cuda")
needs_exe_cuda = pytest.mark.skipif(not os.path.isfile(EXE_CUDA), reason="adapter/_build/ex1_b200_cuda not built "
                                    "(needs /root/reference: make -C oracle ref_cuda && make -C adapter cuda)")


@needs_exe_cuda
@pytest.mark.gpu
@pytest.mark.parametrize("args", [["-n", "6", "-o", "3"], ["-n", "4", "-o", "4"], ["-n", "5", "-o", "2", "-k", "mass"],
                                  ["-n", "3", "-o", "5"]])
def test_adapter_device_resident_mode(args):
    """The same driver against the reference built with its CUDA backend, `-d cuda`: mfem::Device
    selects the device, every library call takes MFEM's own device pointers (x.Read() / y.Write()),
    and the reference side of each comparison is MFEM's own CUDA PA path."""
    r = subprocess.run([EXE_CUDA, "-d", "cuda"] + args, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600)
    out = r.stdout.decode(errors="replace")
    assert r.returncode == 0 and "all checks passed" in out, out[-3000:]
    assert "device-resident" in out


@needs_exe_cuda
@pytest.mark.gpu
def test_adapter_cuda_build_in_host_mode():
    """An MFEM built with CUDA but configured with Device("cpu"): the adapter falls back to its
    own device copies (host-MFEM mode) -- the mode is a run-time property of mfem::Device."""
    r = subprocess.run([EXE_CUDA, "-d", "cpu", "-n", "4", "-o", "3"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       timeout=600)
    out = r.stdout.decode(errors="replace")
    assert r.returncode == 0 and "all checks passed" in out, out[-3000:]
    assert "host-MFEM" in out
